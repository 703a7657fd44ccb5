"""GPU parity of forest / tree evaluation through the C ABI: bit-exact label maps against the NumPy oracle, the C oracle
and the reference's own kernels (oracle/_ref), on the same seeded synthetic inputs."""
import numpy as np
import pytest

from conftest import to_dev, to_np, filled_u16

pytestmark = pytest.mark.gpu


def _api():
    from rdf_b200 import decision_tree as dt
    return dt


def _run_ours(forest_np, depth_np, labels_reduce=1, filt=None, fclass=None, scale=1.0, want_probs=False, prefill=65535):
    import torch
    dt = _api()
    T, NN, E = forest_np.shape
    D = int(np.log2(NN + 1)); C = (E - 7) // 2
    f = dt.DecisionForest(T, D, C)
    f.forest_cu.set(forest_np)
    ev = dt.DecisionTreeEvaluator()
    N, H, W = depth_np.shape
    labels = filled_u16((N, H // labels_reduce, W // labels_reduce), prefill)
    probs = torch.zeros((N, H // labels_reduce, W // labels_reduce, C), dtype=torch.float32, device='cuda') if want_probs else None
    ev.get_labels_forest(f, to_dev(depth_np), labels, labels_reduce=labels_reduce,
                         filter_images=to_dev(filt) if filt is not None else None, filter_images_class=fclass,
                         scale_factor=scale, probs_out=probs)
    torch.cuda.synchronize()
    return to_np(labels), (probs.cpu().numpy() if want_probs else None)


@pytest.mark.parametrize('kind', ['dense-smooth', 'dense-noise', 'live-mask'])
@pytest.mark.parametrize('T,D,C', [(3, 10, 4), (1, 6, 3), (4, 9, 11), (8, 7, 5), (5, 8, 2)])
@pytest.mark.parametrize('ragged', [False, True])
def test_forest_matches_oracles(kind, T, D, C, ragged):
    from rdf_b200 import synth
    from oracle import numpy_oracle as no, c_oracle as co
    depth = synth.depth_frames(kind, 2, 96, 160, seed=7)
    forest = synth.random_forest(T, D, C, seed=11, ragged=ragged)
    ours, probs = _run_ours(forest, depth, want_probs=True)
    exp = np.full(ours.shape, 65535, np.uint16)
    exp_p = np.zeros(ours.shape + (C,), np.float32)
    no.eval_forest(forest, depth, exp, probs_out=exp_p)
    exp_c = np.full(ours.shape, 65535, np.uint16)
    co.eval_forest(forest, depth, exp_c)
    assert np.array_equal(exp, exp_c)
    assert np.array_equal(ours, exp)                       # bit-exact labels
    assert np.abs(probs - exp_p).max() <= 1e-5             # leaf-probability maps: 1e-5 absolute (north_star)


@pytest.mark.parametrize('r,scale', [(1, 1.0), (2, 0.5), (2, 1.0), (3, 0.37)])
def test_labels_reduce_and_scale(r, scale):
    from rdf_b200 import synth
    from oracle import c_oracle as co
    depth = synth.depth_frames('dense-smooth', 1, 121, 213, seed=3)     # ragged sizes: not multiples of the tile or of r
    forest = synth.random_forest(3, 11, 4, seed=5, ragged=True)
    ours, _ = _run_ours(forest, depth, labels_reduce=r, scale=scale)
    exp = np.full(ours.shape, 65535, np.uint16)
    co.eval_forest(forest, depth, exp, labels_reduce=r, scale=scale)
    assert np.array_equal(ours, exp)


def test_filter_and_skip_semantics():
    """Filtered-out pixels and pixels with centre depth 0 / 65535 keep the caller's pre-fill (tree_eval.cu:81-89)."""
    from rdf_b200 import synth
    from oracle import numpy_oracle as no
    depth = synth.depth_frames('dense-noise', 1, 64, 96, seed=9)
    depth[0, 5:20, 7:30] = 0
    depth[0, 30:40, 50:90] = 65535
    forest = synth.random_forest(3, 8, 4, seed=2)
    rng = np.random.default_rng(0)
    filt = rng.integers(0, 3, size=(1, 64, 96)).astype(np.uint16)
    ours, _ = _run_ours(forest, depth, filt=filt, fclass=1, prefill=12345)
    exp = np.full(ours.shape, 12345, np.uint16)
    no.eval_forest(forest, depth, exp, filter_images=filt, filter_class=1)
    assert np.array_equal(ours, exp)
    assert (ours[0, 5:20, 7:30] == 12345).all() and (ours[0, 30:40, 50:90] == 12345).all()
    assert (ours[filt != 1] == 12345).all()


def test_zero_forest_gives_label_zero():
    """Zero-initialised nodes: floor(0) != -1 so both sides are leaves with an all-zero pdf -> label 0 (SURVEY note N2)."""
    from rdf_b200 import synth
    depth = synth.depth_frames('dense-smooth', 1, 32, 48)
    forest = np.zeros((3, 31, 7 + 8), np.float32)
    ours, _ = _run_ours(forest, depth)
    assert (ours == 0).all()


def test_special_float_nodes():
    """NaN / inf / denormal offsets and thresholds follow cvt.rmi + IEEE divide exactly like the oracle."""
    from rdf_b200 import synth
    from oracle import numpy_oracle as no
    depth = synth.depth_frames('dense-noise', 1, 48, 64, seed=4)
    forest = synth.random_forest(2, 6, 3, seed=8)
    rng = np.random.default_rng(1)
    specials = np.array([np.nan, np.inf, -np.inf, 1e-42, -1e-42, 0.0, -0.0, 3e38, -3e38, 0.5, -0.5], np.float32)
    sel = rng.random(forest[:, :, 0:5].shape) < 0.2
    forest[:, :, 0:5][sel] = rng.choice(specials, size=int(sel.sum()))
    ours, _ = _run_ours(forest, depth)
    exp = np.full(ours.shape, 65535, np.uint16)
    no.eval_forest(forest, depth, exp)
    assert np.array_equal(ours, exp)


def test_many_trees_uses_canonical_path():
    from rdf_b200 import synth
    from oracle import c_oracle as co
    depth = synth.depth_frames('dense-smooth', 1, 40, 72)
    forest = synth.random_forest(12, 6, 4, seed=6, ragged=True)
    ours, _ = _run_ours(forest, depth, labels_reduce=2, scale=0.5)
    exp = np.full(ours.shape, 65535, np.uint16)
    co.eval_forest(forest, depth, exp, labels_reduce=2, scale=0.5)
    assert np.array_equal(ours, exp)


@pytest.mark.parametrize('ragged', [False, True])
def test_single_tree(ragged):
    import torch
    from rdf_b200 import synth
    from oracle import numpy_oracle as no
    dt = _api()
    depth = synth.depth_frames('live-mask', 2, 60, 100, seed=5)
    forest = synth.random_forest(1, 9, 5, seed=3, ragged=ragged)
    if ragged:
        forest[0, -64:, 5:7] = -1.0            # some walks fall off the last level: those pixels stay untouched
    tree = dt.DecisionTree(9, 5)
    tree.tree_out_cu.set(forest[0])
    labels = filled_u16((2, 60, 100), 777)
    dt.DecisionTreeEvaluator().get_labels(tree, to_dev(depth), labels)
    torch.cuda.synchronize()
    exp = np.full((2, 60, 100), 777, np.uint16)
    no.eval_tree(forest[0], depth, exp)
    assert np.array_equal(to_np(labels), exp)


def test_forest_update_after_mutation():
    """forest_cu.set(...) after first use must be picked up (packed shadow is re-packed)."""
    import torch
    from rdf_b200 import synth
    from oracle import c_oracle as co
    dt = _api()
    depth = synth.depth_frames('dense-smooth', 1, 48, 80)
    f = dt.DecisionForest(3, 8, 4)
    ev = dt.DecisionTreeEvaluator()
    for seed in (1, 2):
        forest = synth.random_forest(3, 8, 4, seed=seed)
        f.forest_cu.set(forest)
        labels = filled_u16((1, 48, 80))
        ev.get_labels_forest(f, to_dev(depth), labels)
        torch.cuda.synchronize()
        exp = np.full((1, 48, 80), 65535, np.uint16)
        co.eval_forest(forest, depth, exp)
        assert np.array_equal(to_np(labels), exp)


def test_cfg1_full_size_vs_c_oracle_and_reference():
    """BASELINE config 1: one 848x480 frame, 3-tree depth-16 forest, 4 classes."""
    import torch
    from rdf_b200 import synth
    from oracle import c_oracle as co, ref_kernels as rk
    depth = synth.depth_frames('dense-smooth', 1, 480, 848)
    forest = synth.random_forest(3, 16, 4)
    ours, probs = _run_ours(forest, depth, want_probs=True)
    exp = np.full(ours.shape, 65535, np.uint16)
    exp_p = np.zeros(ours.shape + (4,), np.float32)
    co.eval_forest(forest, depth, exp, probs_out=exp_p)
    assert np.array_equal(ours, exp)
    assert np.abs(probs - exp_p).max() <= 1e-5
    assert len(np.unique(ours)) == 4                        # non-degenerate label map
    if True:                       # the reference build is required under -m gpu (tests/conftest.py)
        ref = filled_u16((1, 480, 848))
        rk.eval_forest(to_dev(forest), to_dev(depth), ref)
        torch.cuda.synchronize()
        assert np.array_equal(to_np(ref), ours)            # bit-exact against the reference's own kernel


def test_error_paths():
    import torch
    dt = _api()
    from rdf_b200 import _capi
    lib = _capi.load()
    rc = lib.rdf_eval_forest(None, None, 1, 8, 8, None, -1, None, None, 1, 1.0, None)
    assert rc == -1 and b'NULL' in lib.rdf_last_error()
    with pytest.raises(ValueError):
        _capi.check(rc)
    f = dt.DecisionForest(2, 3, 2)
    ev = dt.DecisionTreeEvaluator()
    with pytest.raises(AssertionError):                     # the reference's shape asserts are kept
        ev.get_labels_forest(f, filled_u16((1, 8, 8), 1000), filled_u16((1, 4, 8)))
    with pytest.raises(ValueError):
        ev.get_labels_forest(f, torch.zeros((1, 8, 8), dtype=torch.int16).view(torch.uint16), filled_u16((1, 8, 8)))  # CPU tensor


def test_device_generators_match_numpy_twins():
    import torch
    from rdf_b200 import synth, _capi
    lib = _capi.load()
    for kind, kid in (('dense-smooth', 0), ('dense-noise', 1), ('live-mask', 2)):
        out = torch.zeros((3, 120, 212), dtype=torch.int16, device='cuda').view(torch.uint16)
        _capi.check(lib.rdf_synth_depth(_capi.dptr(out), kid, 3, 212, 120, 99, 5, _capi.stream_ptr()))
        assert np.array_equal(to_np(out), synth.depth_frames(kind, 3, 120, 212, seed=99, first_frame=5))
    canon = torch.zeros((3, 255, 7 + 8), dtype=torch.float32, device='cuda')
    _capi.check(lib.rdf_synth_forest(_capi.dptr(canon), 3, 8, 4, 4321, _capi.stream_ptr()))
    exp = synth.hash_forest(3, 8, 4, seed=4321)
    assert np.array_equal(canon.cpu().numpy().view(np.uint32), exp.view(np.uint32))


def test_fast_divide_is_bit_identical_to_div_rn():
    """2.6e9 (numerator, divisor) pairs: RN(1/d) + one correction step == div.rn.f32 on the kernels' whole domain."""
    import ctypes
    from rdf_b200 import _capi
    lib = _capi.load()
    bad, bad_floor = ctypes.c_ulonglong(1), ctypes.c_ulonglong(1)
    _capi.check(lib.rdf_selftest_fastdiv(40000, 20261018, ctypes.byref(bad), ctypes.byref(bad_floor)))
    assert bad.value == 0 and bad_floor.value == 0


def test_golden_fixtures_from_reference_kernels():
    """CUDA path against the committed reference-kernel outputs (tests/golden/eval_forest.npz, eval_tree.npz)."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    z = np.load(os.path.join(gold, 'eval_forest.npz'))
    for name in [str(n) for n in z['names']]:
        r, scale, has_filter, fclass = z[f'{name}.params']
        filt = z[f'{name}.filter'] if has_filter else None
        got, _ = _run_ours(z[f'{name}.forest'], z[f'{name}.depth'], int(r), filt, int(fclass) if has_filter else None, float(scale))
        assert np.array_equal(got, z[f'{name}.labels']), name
    import torch
    from rdf_b200 import decision_tree as dt
    z = np.load(os.path.join(gold, 'eval_tree.npz'))
    for name in [str(n) for n in z['names']]:
        tree_np, depth, want = z[f'{name}.tree'], z[f'{name}.depth'], z[f'{name}.labels']
        tree = dt.DecisionTree(int(np.log2(tree_np.shape[0] + 1)), (tree_np.shape[1] - 7) // 2)
        tree.tree_out_cu.set(tree_np)
        out = filled_u16(want.shape, 65535)
        dt.DecisionTreeEvaluator().get_labels(tree, to_dev(depth), out)
        torch.cuda.synchronize()
        assert np.array_equal(to_np(out), want), name


@pytest.mark.parametrize('block', range(8))
def test_fuzz_against_c_oracle(block):
    """Random tiny forests with NaN / inf / huge / denormal-range offsets and thresholds, non-canonical child flags, zero nodes,
    1..9 trees (9 = canonical-layout kernel), depth frames full of 0 / 65535, random labels_reduce / scale / filter."""
    from fuzz_cases import make_case
    from oracle import c_oracle as co
    for seed in range(block * 40, block * 40 + 40):
        c = make_case(seed)
        N, h, w = c['shape']
        if h == 0 or w == 0:
            continue
        exp = np.full((N, h, w), 4321, np.uint16)
        exp_p = np.zeros((N, h, w, c['C']), np.float32)
        co.eval_forest(c['forest'], c['depth'], exp, c['r'], c['filt'], c['fclass'], c['scale'], probs_out=exp_p)
        got, probs = _run_ours(c['forest'], c['depth'], c['r'], c['filt'], c['fclass'], c['scale'], want_probs=True, prefill=4321)
        assert np.array_equal(got, exp), f'seed {seed}: {(got != exp).sum()} label(s) differ'
        written = exp != 4321
        assert np.allclose(probs[written], exp_p[written], atol=1e-5, rtol=0, equal_nan=True), f'seed {seed}'


def test_single_tree_packed_and_canonical_paths_agree():
    """get_labels runs the packed one-tree path (rdf_eval_tree_packed); rdf_eval_tree (canonical array) must agree, including
    'no write when no leaf is reached', and so must both with the C oracle."""
    import torch
    from rdf_b200 import _capi, synth
    from rdf_b200 import decision_tree as dt
    from oracle import c_oracle as co
    for ragged, kill_leaves in ((True, False), (False, True)):
        depth = synth.depth_frames('dense-noise', 2, 60, 80, seed=3)
        tree_np = synth.random_forest(1, 8, 5, seed=11, ragged=ragged)[0]
        if kill_leaves:
            tree_np[191:, 5:7] = -1.0                                  # half of the last level (rows 127..254) never reaches a leaf
        tree = dt.DecisionTree(8, 5)
        tree.tree_out_cu.set(tree_np)
        a = filled_u16((2, 60, 80), 4242)
        dt.DecisionTreeEvaluator().get_labels(tree, to_dev(depth), a)
        b = filled_u16((2, 60, 80), 4242)
        _capi.check(_capi.load().rdf_eval_tree(_capi.dptr(tree.tree_out_cu), 8, 5, _capi.dptr(to_dev(depth)), 2, 80, 60, _capi.dptr(b),
                                               _capi.stream_ptr()))
        torch.cuda.synchronize()
        exp = np.full((2, 60, 80), 4242, np.uint16)
        co.eval_tree(tree_np, depth, exp)
        assert np.array_equal(to_np(a), exp) and np.array_equal(to_np(b), exp)
        if kill_leaves:
            assert (exp == 4242).any() and (exp != 4242).any()


@pytest.mark.parametrize('kind,T,D,C,r', [('dense-smooth', 3, 8, 4, 1), ('dense-noise', 4, 7, 11, 1), ('live-mask', 2, 9, 3, 2), ('dense-noise', 8, 5, 4, 3)])
def test_texture_probe_variant_matches_oracle(kind, T, D, C, r):
    """rdf_eval_forest_tex (probes through the texture units, border addressing = the 65535 default) against the C oracle, with a
    filter image; forests with out-of-domain offsets are refused."""
    import ctypes
    import torch
    from rdf_b200 import _capi, synth
    from rdf_b200 import decision_tree as dt
    from oracle import c_oracle as co
    lib = _capi.load()
    N, H, W = 3, 61, 83
    depth_np = synth.depth_frames(kind, N, H, W, seed=T)
    depth_np[0, 3:6, 4:9] = 0
    forest_np = synth.random_forest(T, D, C, seed=D, ragged=True)
    filt_np = (np.arange(N * (H // r) * (W // r)).reshape(N, H // r, W // r) % 3).astype(np.uint16)
    f = dt.DecisionForest(T, D, C)
    f.forest_cu.set(forest_np)
    d, filt = to_dev(depth_np), to_dev(filt_np)
    lab = filled_u16((N, H // r, W // r), 777)
    h = ctypes.c_void_p()
    _capi.check(lib.rdf_depth_tex_create(N, W, H, ctypes.byref(h)))
    try:
        _capi.check(lib.rdf_depth_tex_upload(h, _capi.dptr(d), N, _capi.stream_ptr()))
        _capi.check(lib.rdf_eval_forest_tex(f.handle(), h, _capi.dptr(d), N, _capi.dptr(filt), 1, _capi.dptr(lab), r, _capi.stream_ptr()))
        torch.cuda.synchronize()
        exp = np.full((N, H // r, W // r), 777, np.uint16)
        co.eval_forest(forest_np, depth_np, exp, r, filt_np, 1)
        assert np.array_equal(to_np(lab), exp)
        forest_np[0, 0, 0] = np.inf                                 # exact-divide node: the texture variant must refuse
        f.forest_cu.set(forest_np)
        assert lib.rdf_eval_forest_tex(f.handle(), h, _capi.dptr(d), N, None, -1, _capi.dptr(lab), r, _capi.stream_ptr()) == -3
    finally:
        lib.rdf_depth_tex_destroy(h)


_BLOCK_LAYOUT_SCRIPT = r'''
import sys
sys.path.insert(0, r"%(root)s"); sys.path.insert(0, r"%(root)s/3d-beats_b200")
import numpy as np, torch
from rdf_b200 import synth
from rdf_b200 import decision_tree as dt
from oracle import c_oracle as co
for (T, D, C, ragged, kind, r, scale) in [(3, 9, 4, True, 'dense-noise', 1, 1.0), (4, 12, 5, False, 'dense-smooth', 2, 0.5), (1, 7, 3, True, 'live-mask', 1, 1.0),
                                          (8, 10, 4, True, 'dense-noise', 1, 1.0), (2, 5, 4, False, 'dense-smooth', 1, 1.0)]:
    f = synth.random_forest(T, D, C, seed=11 + D, ragged=ragged)
    d = synth.depth_frames(kind, 2, 96, 160, seed=5)
    forest = dt.DecisionForest(T, D, C); forest.forest_cu.set(f)
    out = dt.cu_array.GPUArray((2, 96 // r, 160 // r), dtype=np.uint16).fill(65535)
    dt.DecisionTreeEvaluator().get_labels_forest(forest, dt.cu_array.to_gpu(d), out, labels_reduce=r, scale_factor=scale)
    exp = np.full((2, 96 // r, 160 // r), 65535, np.uint16)
    co.eval_forest(f, d, exp, r, None, None, scale)
    assert np.array_equal(out.get(), exp), (T, D, C)
print('ok')
'''


def test_block_layout_is_bit_exact():
    """RDF_PACK_LAYOUT=blocks (subtree blocks below level 6, csrc/rdf_common.cuh:rdf_blocks_row) gives the same label maps as the
    heap order: odd and even numbers of paired levels, ragged trees, trees shallower than the heap-ordered top, 8 trees.  The layout
    is chosen once per process, hence the subprocess.  (Measured slower than heap order: profiles/r02_layout.md.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-c', _BLOCK_LAYOUT_SCRIPT % {'root': root}], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, RDF_PACK_LAYOUT='blocks'))
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]


@pytest.mark.parametrize('N,chunk,ramp', [(61, 16, 2), (61, 16, 0), (5, 16, 8), (64, 8, 1), (200, 32, 4)])
def test_host_batch_evaluator_matches_resident_run(N, chunk, ramp):
    """HostBatchEvaluator (pinned host frames in, pinned host label maps out, ramped chunk schedule on three streams) returns the
    label maps of one resident get_labels_forest call, skipped pixels holding the pre-fill (test_on_saved_model.py:46-58)."""
    import torch
    from rdf_b200 import synth
    from rdf_b200.pipeline import HostBatchEvaluator, pinned_like
    dt = _api()
    H, W = 40, 72
    depth = synth.depth_frames('live-mask', N, H, W, seed=21)
    forest_np = synth.random_forest(3, 9, 4, seed=4, ragged=True)
    want, _ = _run_ours(forest_np, depth, prefill=4321)
    f = dt.DecisionForest(3, 9, 4)
    f.forest_cu.set(forest_np)
    hb = HostBatchEvaluator(dt.DecisionTreeEvaluator(), f, (H, W), chunk_frames=chunk, ramp=ramp)
    sizes = hb.chunk_sizes(N)
    assert sum(sizes) == N and max(sizes) <= chunk and min(sizes) >= 1
    depth_host = pinned_like((N, H, W), np.uint16)
    labels_host = pinned_like((N, H, W), np.uint16)
    depth_host.view(torch.int16).copy_(torch.from_numpy(depth.view(np.int16)))
    for _ in range(2):                                               # the second run reuses buffers the first one left busy
        labels_host.view(torch.int16).fill_(-1)
        hb.run(depth_host, labels_host, prefill=4321)
        torch.cuda.synchronize()
        got = labels_host.view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got, want)
    assert hb.bytes_h2d == depth.nbytes and hb.bytes_d2h == want.nbytes
