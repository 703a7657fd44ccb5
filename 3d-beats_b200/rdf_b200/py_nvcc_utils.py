"""Stand-in for the reference's kernel loader (src/cuda/py_nvcc_utils.py).

The reference JIT-compiles ./src/cuda/<n>.cu with pycuda (or loads fatbins selected by --fatbin_in/--fatbin_out).
Here every kernel is compiled ahead of time for sm_100a into librdf_b200.so, so the two flags are accepted for
command-line compatibility and ignored; `get_module` has no meaning any more and says so.
"""
from . import _capi


def add_args(parser):
    # same flags (and the reference's swapped help texts are not reproduced)
    parser.add_argument('--fatbin_in', nargs='?', required=False, type=str, help='ignored: kernels are prebuilt for sm_100a')
    parser.add_argument('--fatbin_out', nargs='?', required=False, type=str, help='ignored: kernels are prebuilt for sm_100a')


def config_compiler(args):
    f_in = getattr(args, 'fatbin_in', None)
    f_out = getattr(args, 'fatbin_out', None)
    assert not (f_in and f_out)          # src/cuda/py_nvcc_utils.py:16
    _capi.load()                         # fail here, loudly, if the library is missing


def get_module(n):
    raise NotImplementedError(
        f"get_module({n!r}): rdf_b200 has no runtime-compiled CUDA modules; the '{n}' kernels are C-ABI entry points of "
        f"{_capi.LIB_PATH} (see include/rdf_b200.h)")
