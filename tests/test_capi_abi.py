"""CPU gate for the drop-in boundary: librdf_b200.so loads without a GPU, exports every symbol include/rdf_b200.h declares,
the ctypes binding covers exactly that set, and argument validation works (returns an error code + message; never aborts,
never computes on the CPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'rdf_b200.h')


def _declared():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'RDF_API\s+[\w\s\*]+?\b(rdf_\w+)\s*\(', text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ['rdf_forest_create', 'rdf_eval_forest', 'rdf_eval_tree', 'rdf_composite', 'rdf_layered_run', 'rdf_mean_shift',
                 'rdf_train_hist', 'rdf_train_pick_best', 'rdf_train_next_active', 'rdf_train_advance_pixels', 'rdf_last_error']:
        assert must in names
    # every entry point cites the reference interface it replaces
    text = open(HEADER).read()
    for ref in ['src/cuda/tree_eval.cu:24-137', 'src/cuda/tree_eval.cu:140-212', 'src/cuda/tree_eval.cu:214-248',
                'src/decision_tree.py:233-264', 'src/cuda/mean_shift.py:19-59', 'src/cuda/tree_train.cu:4-64',
                'src/cuda/tree_train.cu:99-236']:
        assert ref in text, ref


def test_library_exports_every_declared_symbol():
    from rdf_b200 import _capi
    lib = _capi.load()                                   # must work with no GPU present
    for name in _declared():
        assert hasattr(lib, name), f'{name} declared in include/rdf_b200.h but not exported by librdf_b200.so'
    assert sorted(_capi.SIGNATURES) == _declared(), 'ctypes binding and header disagree'
    assert lib.rdf_version() >= 100


def test_no_hidden_exports():
    """Only rdf_* symbols are public (built with -fvisibility=hidden): the library is a C ABI, not a C++ one."""
    import subprocess
    from rdf_b200 import _capi
    out = subprocess.run(['nm', '-D', '--defined-only', _capi.LIB_PATH], capture_output=True, text=True).stdout
    mine = [ln.split()[-1] for ln in out.splitlines() if ' T ' in ln]
    assert mine and all(s.startswith('rdf_') for s in mine), [s for s in mine if not s.startswith('rdf_')][:5]


def test_argument_validation_without_gpu():
    from rdf_b200 import _capi
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.rdf_forest_create(None, 3, 16, 4, None, ctypes.byref(h)) == -1            # RDF_ERR_INVALID
    assert b'canon_dev' in lib.rdf_last_error()
    assert lib.rdf_forest_create(ctypes.c_void_p(16), 3, 99, 4, None, ctypes.byref(h)) == -1
    assert b'max_depth' in lib.rdf_last_error()
    assert lib.rdf_eval_forest(None, None, 1, 8, 8, None, -1, None, None, 1, 1.0, None) == -1
    assert lib.rdf_mean_shift(None, 8, 8, 2, None, 1, None, None, 0, None) == -1
    n = ctypes.c_size_t()
    assert lib.rdf_mean_shift_workspace_bytes(424, 240, 11, ctypes.byref(n)) == 0 and n.value >= 424 * 240 * 4
    with pytest.raises(ValueError):
        _capi.check(-1)
    with pytest.raises(_capi.RdfError):
        _capi.check(-2)


def test_product_has_no_cpu_path():
    """The product package never imports the oracle, and refuses host tensors."""
    import torch
    from rdf_b200 import _capi
    pkg = os.path.join(ROOT, '3d-beats_b200', 'rdf_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn
    with pytest.raises(ValueError):
        _capi.dptr(torch.zeros(4))
    if not torch.cuda.is_available():
        from rdf_b200 import buffers
        with pytest.raises(RuntimeError):
            buffers.GPUArray((4,), dtype='float32')


def _build_c_smoke(tmp_path):
    import subprocess
    from rdf_b200 import _capi
    exe = str(tmp_path / 'abi_smoke')
    libdir = os.path.dirname(_capi.LIB_PATH)
    cmd = ['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), '-I', '/usr/local/cuda/include',
           os.path.join(ROOT, 'tests', 'c', 'abi_smoke.c'), '-o', exe, '-L', libdir, '-lrdf_b200', '-L', '/usr/local/cuda/lib64', '-lcudart',
           '-Wl,-rpath,' + libdir, '-Wl,-rpath,/usr/local/cuda/lib64']
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    return exe


def test_plain_c_program_links_and_gets_the_error_contract(tmp_path):
    """A C99 program compiled with gcc consumes include/rdf_b200.h and librdf_b200.so directly (no C++, no torch)."""
    import subprocess
    exe = _build_c_smoke(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'abi_smoke ok' in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_plain_c_program_device_path(tmp_path):
    import subprocess
    exe = _build_c_smoke(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'device path' in out.stdout, out.stdout + out.stderr


def test_frame_entry_points_validate_arguments_without_gpu():
    """The live-frame entry points refuse bad arguments with a message, on a box without a GPU."""
    from rdf_b200 import _capi
    lib = _capi.load()
    p16 = ctypes.c_void_p(16)
    assert lib.rdf_condition_depth(None, 8, 8, 0., 0., 1., p16, 40., None, 5, 3, p16, None, None) == -1
    assert b'rdf_condition_depth' in lib.rdf_last_error()
    assert lib.rdf_condition_depth(p16, 8, 8, 0., 0., 1., p16, 40., None, 5, 3, p16, None, None) == -1          # in == out
    assert b'alias' in lib.rdf_last_error()
    assert lib.rdf_condition_depth(p16, 8, 8, 0., 0., 1., p16, 40., p16, 4, 3, ctypes.c_void_p(32), None, None) == -1   # even window
    assert b'k_size' in lib.rdf_last_error()
    ids = (ctypes.c_int * 5)(1, 2, 3, 4, 5)
    assert lib.rdf_stencil_hands(p16, 8, 8, p16, 3, 1, 5, ids, ids, ctypes.c_void_p(32), None) == -1               # > 4 hands
    assert b'num_hands' in lib.rdf_last_error()
    assert lib.rdf_fingertip_z(p16, 1, 11, ids, 0, 2, p16, 8, 8, 0., 0., 1., 1., p16, p16, None, None) == -1
    assert b'num_fingertips' in lib.rdf_last_error()
    assert lib.rdf_flip_x(p16, 8, 8, p16, None) == -1 and lib.rdf_grow_groups(p16, 8, 8, p16, None) == -1       # aliased
    assert lib.rdf_mean_shift_batch(None, 2, 8, 8, 2, None, 1, None, None, 0, None) == -1


def test_eval_kernel_reads_its_upper_levels_from_the_constant_bank():
    """SASS of the shipped forest-eval instance (T = 4, complete trees, scale 1): the upper levels are register-indexed constant
    loads from the launch parameters, the levels below are 256-bit global loads, and the kernel neither touches shared memory nor
    meets at a barrier (csrc/rdf_eval.cu: rdf_eval_launch; profiles/r02_eval_const_top.md)."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sym = '_Z22rdf_eval_packed_kernelILi4ELi16ELb1ELi3EEv15rdf_eval_launchIXT_EE'
    from rdf_b200 import _capi
    out = subprocess.run([cuobjdump, '-sass', '-fun', sym, _capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert 'sm_100a' in out and 'Function : ' + sym in out, out[:500]
    assert len(re.findall(r'LDC\.64 R\d+, c\[0x0\]\[R\d+\+', out)) >= 8          # 3 x LDC.64 + 1 x LDC per header, 4 trees
    assert len(re.findall(r'LDG\.E\.ENL2\.256', out)) >= 4                       # one 256-bit header load per tree below
    assert 'FFMA2' in out and 'FADD2.RM' in out                                  # packed fp32 divide + magic-number floor
    assert not re.search(r'\bLDS\b|\bSTS\b|BAR\.SYNC', out)                      # no staging, no barrier
