"""Size-independent properties of the CUDA path at (or near) BASELINE.json's full sizes, where the NumPy oracle is too slow to
be the checker for everything: batch invariance, labels_reduce consistency, filter = subset, idempotence, histogram linearity
and conservation, plus a C-oracle spot check of the cfg3 forest shape (T=4, D=20)."""
import ctypes

import numpy as np
import pytest

from conftest import to_dev, to_np, filled_u16

pytestmark = pytest.mark.gpu


def _forest(T, D, C, seed=1234):
    import torch
    from rdf_b200 import _capi
    from rdf_b200 import decision_tree as dt
    f = dt.DecisionForest(T, D, C)
    _capi.check(_capi.load().rdf_synth_forest(_capi.dptr(f.forest_cu), T, D, C, seed, _capi.stream_ptr()))
    torch.cuda.synchronize()
    return f


def _depth(kind, N, H, W, first=0, seed=1234):
    import torch
    from rdf_b200 import _capi
    from rdf_b200 import decision_tree as dt
    d = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16)
    _capi.check(_capi.load().rdf_synth_depth(_capi.dptr(d), {'dense-smooth': 0, 'dense-noise': 1, 'live-mask': 2}[kind], N, W, H, seed,
                                             first, _capi.stream_ptr()))
    torch.cuda.synchronize()
    return d


@pytest.mark.parametrize('kind', ['dense-smooth', 'dense-noise'])
def test_cfg3_shape_batch_invariance_and_oracle_spot_check(kind):
    """cfg3 forest (T=4, D=20, C=4) over 24 full-size frames: one launch == per-frame launches == shifted batch; frame 5 == C oracle."""
    import torch
    from rdf_b200 import decision_tree as dt
    from oracle import c_oracle as co
    H, W, N = 480, 848, 24
    forest = _forest(4, 20, 4)
    ev = dt.DecisionTreeEvaluator()
    depth = _depth(kind, N, H, W)
    a = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16).fill(65535)
    ev.get_labels_forest(forest, depth, a)
    b = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16).fill(65535)
    for n in range(N):
        ev.get_labels_forest(forest, depth[n:n + 1], b[n:n + 1])
    torch.cuda.synchronize()
    assert torch.equal(a.tensor.view(torch.int16), b.tensor.view(torch.int16))
    # a batch that starts at frame 7 (what another rank of a frame-sharded job sees)
    shifted = _depth(kind, 8, H, W, first=7)
    c = dt.cu_array.GPUArray((8, H, W), dtype=np.uint16).fill(65535)
    ev.get_labels_forest(forest, shifted, c)
    torch.cuda.synchronize()
    assert torch.equal(c.tensor.view(torch.int16), a.tensor[7:15].view(torch.int16))
    exp = np.full((1, H, W), 65535, np.uint16)
    co.eval_forest(forest.forest_cu.get(), depth[5:6].get(), exp)
    assert np.array_equal(a[5:6].get(), exp)
    assert len(np.unique(exp)) >= 3                               # non-degenerate label map


def test_labels_reduce_is_subsampling_and_filter_is_subset():
    import torch
    from rdf_b200 import decision_tree as dt
    H, W = 480, 848
    forest = _forest(3, 16, 4)
    ev = dt.DecisionTreeEvaluator()
    depth = _depth('live-mask', 1, H, W)
    full = dt.cu_array.GPUArray((1, H, W), dtype=np.uint16).fill(65535)
    ev.get_labels_forest(forest, depth, full)
    for r in (2, 3, 4):
        red = dt.cu_array.GPUArray((1, H // r, W // r), dtype=np.uint16).fill(65535)
        ev.get_labels_forest(forest, depth, red, labels_reduce=r)
        torch.cuda.synchronize()
        # labels pixel (x, y) samples depth pixel (x*r, y*r) with the same features (tree_eval.cu:69-70)
        assert np.array_equal(red.get()[0], full.get()[0, ::r, ::r][:H // r, :W // r])
    # filter: only pixels whose filter label equals the class are written; they get the unfiltered label
    filt = (full.get() % 3).astype(np.uint16)
    out = dt.cu_array.GPUArray((1, H, W), dtype=np.uint16).fill(777)
    ev.get_labels_forest(forest, depth, out, filter_images=dt.cu_array.to_gpu(filt), filter_images_class=1)
    torch.cuda.synchronize()
    got, ref = out.get(), full.get()
    sel = (filt == 1) & (ref != 65535)
    assert np.array_equal(got[sel], ref[sel]) and (got[~sel] == 777).all()
    # idempotence: a second run over an already written map changes nothing
    ev.get_labels_forest(forest, depth, full)
    torch.cuda.synchronize()
    assert np.array_equal(full.get(), ref)


def test_cfg5_shape_deep_forest_matches_c_oracle_on_a_band():
    """8 trees of depth 22 (cfg5 is depth 24: same code path, 4x fewer nodes to keep the host copy small): rows 300..307 of a
    1280x720 frame against the C oracle."""
    import torch
    from rdf_b200 import decision_tree as dt
    from oracle import c_oracle as co
    H, W = 720, 1280
    forest = _forest(8, 22, 4, seed=99)
    depth = _depth('dense-noise', 1, H, W, seed=99)
    out = dt.cu_array.GPUArray((1, H, W), dtype=np.uint16).fill(65535)
    dt.DecisionTreeEvaluator().get_labels_forest(forest, depth, out)
    torch.cuda.synchronize()
    canon = forest.forest_cu.get()
    d = depth.get()
    # the oracle evaluates a band by treating it as the labels image of a filter: use filter_images to restrict rows
    filt = np.zeros((1, H, W), np.uint16)
    filt[0, 300:308] = 1
    exp = np.full((1, H, W), 65535, np.uint16)
    co.eval_forest(canon, d, exp, 1, filt, 1)
    assert np.array_equal(out.get()[0, 300:308], exp[0, 300:308])


@pytest.mark.parametrize('kind', ['dense-smooth', 'dense-noise'])
def test_cfg5_at_depth_24_matches_the_reference_kernel(kind):
    """BASELINE configs[4] at its stated size: 8 trees of depth 24 (134 M nodes: 8.05 GB canonical, 8.6 GB packed; leaf ids up
    to 2.7e8, header byte offsets up to 4.3e9) on a whole 1280x720 frame, against the reference's own
    evaluate_image_using_forest (src/cuda/tree_eval.cu:24-137, compiled unchanged) run on the same device - no host copy of the
    forest is needed.  This is where an overflow in the packed ids (csrc/rdf_capi.cu pack, mad.wide.u32 header addressing in
    csrc/rdf_traverse.cuh) would show."""
    import torch
    from rdf_b200 import decision_tree as dt
    from oracle import ref_kernels as rk
    H, W = 720, 1280
    forest = _forest(8, 24, 4, seed=77)
    depth = _depth(kind, 1, H, W, seed=77)
    out = dt.cu_array.GPUArray((1, H, W), dtype=np.uint16).fill(65535)
    dt.DecisionTreeEvaluator().get_labels_forest(forest, depth, out)
    ref = filled_u16((1, H, W), 65535)
    rk.eval_forest(forest.forest_cu.tensor, depth.tensor, ref)
    torch.cuda.synchronize()
    got, want = to_np(out.tensor), to_np(ref)
    assert (want != 65535).all()                                   # dense frames: every pixel is evaluated
    assert len(np.unique(want)) >= 3                               # not a degenerate label map
    assert np.array_equal(got, want), f'{(got != want).sum()} of {got.size} labels differ from the reference kernel at D=24'
    # the deepest level is really reached through the packed ids: the last tree's last-level rows are far above 2^31 bytes
    info = [ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()]
    from rdf_b200 import _capi
    _capi.check(_capi.load().rdf_forest_info(forest.handle(), *[ctypes.byref(i) for i in info]))
    assert info[1].value == 24 and info[3].value == 8 * ((1 << 24) - 1) * (32 + 2 * 4 * 4)      # 8.59e9 packed bytes
    del forest, depth, out, ref
    torch.cuda.empty_cache()


def test_histogram_linearity_and_conservation_at_cfg4_width():
    """hist(images A + B) == hist(A) + hist(B) (what the multi-GPU allreduce relies on) and every feature row counts every
    active pixel exactly once, at 848x480 with 200 features x 64 thresholds."""
    import torch
    from rdf_b200 import _capi, synth
    lib = _capi.load()
    H, W, C, F, NT, level = 480, 848, 4, 200, 64, 5
    N = 4
    depth = synth.depth_frames('dense-smooth', N, H, W)
    labels = synth.train_labels(N, H, W)
    labels[2, 100:140] = 0
    nodes = synth.random_node_assignment(labels, level)
    off, th = synth.random_proposals(F, NT)
    S = 1 << level
    slot = torch.arange(S, dtype=torch.int32, device='cuda')
    offd, thd = to_dev(off), to_dev(th)

    def hist_of(n0, n1):
        d, l, nd = to_dev(depth[n0:n1]), to_dev(labels[n0:n1]), to_dev(np.ascontiguousarray(nodes[n0:n1]))
        npx = (n1 - n0) * H * W
        need = ctypes.c_size_t()
        _capi.check(lib.rdf_train_bucket_workspace_bytes(npx, S, ctypes.byref(need)))
        ws = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')
        h = torch.zeros((S, F, NT + 1, C), dtype=torch.int32, device='cuda')
        _capi.check(lib.rdf_train_bucket(_capi.dptr(nd), npx, _capi.dptr(slot), S, _capi.dptr(ws), need.value, _capi.stream_ptr()))
        _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(d), _capi.dptr(l), n1 - n0, W, H, _capi.dptr(ws), S, _capi.dptr(offd),
                                                _capi.dptr(thd), F, NT, C, _capi.dptr(h), _capi.stream_ptr()))
        torch.cuda.synchronize()
        return h

    whole, a, b = hist_of(0, N), hist_of(0, 1), hist_of(1, N)
    assert torch.equal(whole, a + b)
    per_node = np.bincount(nodes[nodes >= 0], minlength=S)
    got = whole.sum(dim=(2, 3)).cpu().numpy()
    assert np.array_equal(got, np.repeat(per_node[:, None], F, axis=1))
    assert int(whole[..., 0].sum()) == 0
