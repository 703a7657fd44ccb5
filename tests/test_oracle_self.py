"""CPU tests of the oracles themselves: the NumPy restatement and the C restatement must agree with each other on seeded
inputs, and both must show the invariants the reference's device asserts / quirks imply (note N2)."""
import numpy as np
import pytest

from rdf_b200 import synth
from oracle import numpy_oracle as no, c_oracle as co


@pytest.mark.parametrize('kind', ['dense-smooth', 'dense-noise', 'live-mask'])
@pytest.mark.parametrize('T,D,C,ragged', [(3, 8, 4, False), (2, 7, 11, True), (5, 5, 3, True)])
def test_numpy_and_c_eval_agree(kind, T, D, C, ragged):
    depth = synth.depth_frames(kind, 2, 48, 64, seed=3)
    forest = synth.random_forest(T, D, C, seed=9, ragged=ragged)
    a = np.full((2, 48, 64), 65535, np.uint16)
    b = a.copy()
    pa = np.zeros((2, 48, 64, C), np.float32)
    pb = pa.copy()
    no.eval_forest(forest, depth, a, probs_out=pa)
    co.eval_forest(forest, depth, b, probs_out=pb)
    assert np.array_equal(a, b)
    assert np.array_equal(pa, pb)
    if kind != 'live-mask':
        assert (a != 65535).all()


@pytest.mark.parametrize('r,scale', [(2, 0.5), (3, 0.37), (1, 2.0)])
def test_labels_reduce_scale_filter(r, scale):
    H, W = 60, 90
    depth = synth.depth_frames('dense-smooth', 1, H, W, seed=4)
    forest = synth.random_forest(3, 7, 4, seed=2, ragged=True)
    h, w = H // r, W // r
    filt = (np.arange(h * w).reshape(1, h, w) % 3).astype(np.uint16)
    a = np.full((1, h, w), 7, np.uint16)
    b = a.copy()
    no.eval_forest(forest, depth, a, r, filt, 1, scale)
    co.eval_forest(forest, depth, b, r, filt, 1, scale)
    assert np.array_equal(a, b)
    assert (a[filt != 1] == 7).all()                  # filtered-out pixels keep the caller's pre-fill (tree_eval.cu:81-85)
    assert (a[filt == 1] < 4).all()


def test_skip_and_probe_semantics():
    """centre depth 0 / 65535 -> untouched (tree_eval.cu:88-89); out-of-image probes read 65535; a probed 0 is used as 0."""
    depth = np.full((1, 8, 8), 1000, np.uint16)
    depth[0, 0, 0] = 0
    depth[0, 0, 1] = 65535
    C = 3
    forest = np.zeros((1, 1, 7 + 2 * C), np.float32)
    # u probes 2 px to the right (offset 2000/1000), v probes the centre; feature = d(x+2) - d(x)
    forest[0, 0, 0:5] = (2000.0, 0.0, 0.0, 0.0, 1.0)
    forest[0, 0, 7:7 + C] = (0.0, 0.5, 0.25)          # left  (f < 1): in-image probe, equal depths -> label 1
    forest[0, 0, 7 + C:] = (0.0, 0.25, 0.5)           # right (f >= 1): probe fell off the image (65535 - 1000) -> label 2
    for fn in (no.eval_forest, co.eval_forest):
        out = np.full((1, 8, 8), 9, np.uint16)
        fn(forest, depth, out)
        assert out[0, 0, 0] == 9 and out[0, 0, 1] == 9
        assert (out[0, 1:, :6] == 1).all()
        assert (out[0, :, 6:] == 2).all()
    depth[0, 4, 6] = 0                                # probed zero: feature = 0 - 1000 < 1 -> left
    for fn in (no.eval_forest, co.eval_forest):
        out = np.full((1, 8, 8), 9, np.uint16)
        fn(forest, depth, out)
        assert out[0, 4, 4] == 1 and out[0, 4, 6] == 9


def test_zero_forest_gives_label_zero():
    """zero-initialised node: floor(0) != -1 -> leaf with an all-zero pdf -> label 0 (note N2)."""
    depth = synth.depth_frames('dense-smooth', 1, 16, 16)
    forest = np.zeros((2, 7, 15), np.float32)
    for fn in (no.eval_forest, co.eval_forest):
        out = np.full((1, 16, 16), 65535, np.uint16)
        fn(forest, depth, out)
        assert (out == 0).all()


def test_first_strict_max_argmax():
    acc = np.array([[0.0, 0.5, 0.5, 0.25], [0.0, 0.0, 0.0, 0.0], [-1.0, -2.0, 0.0, -0.5]], np.float32)
    assert list(no.best_pdf_chance(acc)) == [1, 0, 0]


def test_float2int_rd_matches_cvt_rmi():
    x = np.array([0.5, -0.5, -1.0, -1.000001, 3e9, -3e9, np.nan, np.inf, -np.inf, -0.0], np.float32)
    want = [0, -1, -1, -2, 2**31 - 1, -2**31, 0, 2**31 - 1, -2**31, 0]
    assert list(no.float2int_rd(x)) == want


def test_single_tree_agree_and_fall_off():
    depth = synth.depth_frames('dense-noise', 2, 30, 40, seed=8)
    tree = synth.random_forest(1, 6, 4, seed=5, ragged=True)[0]
    a = np.full((2, 30, 40), 65535, np.uint16)
    b = a.copy()
    no.eval_tree(tree, depth, a)
    co.eval_tree(tree, depth, b)
    assert np.array_equal(a, b)
    tree[:, 5:7] = -1.0
    a[:] = 1234
    co.eval_tree(tree, depth, a)
    assert (a == 1234).all()                          # falls off the last level: nothing written (tree_eval.cu:174-210)


def test_composite_walk():
    cond = np.array([[1, 2], [0, 11], [0, 1], [0, 2]], np.int32)
    l0 = np.array([[1, 2, 0, 65535, 1, 1]], np.uint16)
    l1 = np.array([[1, 2, 1, 1, 0, 65535]], np.uint16)
    for fn in (no.composite, co.composite):
        out = np.full((1, 6), 777, np.uint16)
        fn([l0, l1], cond, out)
        assert list(out[0]) == [1, 11, 777, 777, 777, 777]


def test_mean_shift_agree_and_nan_pattern():
    rng = np.random.default_rng(1)
    lab = np.full((40, 50), 65535, np.uint16)
    lab[5:15, 5:20] = 1
    lab[20:35, 30:45] = 3
    lab[rng.integers(0, 40, 30), rng.integers(0, 50, 30)] = 2
    var = np.array([4.0, 6.0, 8.0, 5.0], np.float32)
    a = no.mean_shift(lab, 4, var, 6)
    b = co.mean_shift(lab, 4, var, 6)
    assert np.isnan(a[3]).all() and np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) <= 1e-9
    r0 = no.mean_shift(lab, 4, var, 1)
    ys, xs = np.nonzero(lab == 1)
    assert abs(r0[0, 0] - xs.mean()) < 1e-12 and abs(r0[0, 1] - ys.mean()) < 1e-12     # round 0 = plain centroid (mean_shift.cu:31-34)


@pytest.mark.parametrize('level,NT', [(0, 8), (3, 64), (5, 1)])
def test_train_hist_agree_and_conservation(level, NT):
    N, H, W, C, F = 2, 40, 56, 4, 6
    depth = synth.depth_frames('dense-smooth', N, H, W, seed=6)
    depth[0, 3:5, 3:5] = 0
    labels = synth.train_labels(N, H, W)
    labels[1, :4, :] = 0
    nodes = synth.random_node_assignment(labels, level, seed=3)
    n_nodes = 1 << level
    slot = np.arange(n_nodes, dtype=np.int32)
    off, th = synth.random_proposals(F, NT, seed=2)
    a = no.train_hist(depth, labels, nodes, slot, n_nodes, off, th, C)
    b = co.train_hist(depth, labels, nodes, slot, n_nodes, off, th, C)
    assert np.array_equal(a, b)
    per_node = np.bincount(nodes[nodes >= 0], minlength=n_nodes)
    assert np.array_equal(a.sum(axis=(2, 3)), np.repeat(per_node[:, None], F, axis=1))     # left + right == parent
    assert a[..., 0].sum() == 0                                                            # class 0 is never counted


def test_gini_gain_oracles_agree():
    rng = np.random.default_rng(0)
    for _ in range(200):
        left = rng.integers(0, 5000, size=4).astype(np.uint64)
        right = rng.integers(0, 5000, size=4).astype(np.uint64)
        if left.sum() == 0 or right.sum() == 0:
            continue
        parent = left + right
        a = no.gini_gain(parent, left, right)
        b = co.gini_gain(parent, left, right)
        assert abs(float(a) - float(b)) <= 1e-6
        assert -1e-6 <= float(b) <= 1.0
    pure_l = np.array([0, 10, 0, 0], np.uint64)
    pure_r = np.array([0, 0, 10, 0], np.uint64)
    assert abs(float(co.gini_gain(pure_l + pure_r, pure_l, pure_r)) - 0.5) <= 1e-7


def test_generators_are_deterministic_and_in_range():
    a = synth.depth_frames('dense-smooth', 2, 20, 30, seed=1, first_frame=5)
    b = synth.depth_frames('dense-smooth', 7, 20, 30, seed=1)[5:7]
    assert np.array_equal(a, b)                                   # frame n depends only on (seed, n, y, x): shardable
    assert a.min() >= 3000 and a.max() <= 4054
    n = synth.depth_frames('dense-noise', 1, 20, 30)
    assert n.min() >= 1 and n.max() <= 65534
    m = synth.depth_frames('live-mask', 1, 48, 84)
    assert (m == 65535).any() and (m != 65535).any()
    f = synth.random_forest(2, 5, 3)
    assert f.shape == (2, 31, 13) and (f[:, 15:, 5:7] == 0).all() and (f[:, :15, 5:7] == -1).all()
    assert np.array_equal(f[:, :, 7:] * 1024, np.round(f[:, :, 7:] * 1024))              # dyadic pdfs (note N1)
    h = synth.hash_forest(3, 4, 4, seed=7)
    assert np.array_equal(h[1:2], synth.hash_forest(3, 4, 4, seed=7, trees=[1]))
