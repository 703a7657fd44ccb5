"""GPU parity of the stacked (layered) forest and mean shift: fused launch vs NumPy oracle vs the reference's own kernels."""
import json
import os

import numpy as np
import pytest

from conftest import to_dev, to_np, filled_u16

pytestmark = pytest.mark.gpu


def _layered(tmp_path, depth_dims, r, max_depth=10, seed=1234):
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    forests, cfg, variances = synth.layered_cfg2(seed=seed, max_depth=max_depth)
    path = synth.write_layered_model(str(tmp_path), forests, cfg)
    ldf = dt.LayeredDecisionForest.load(path, depth_dims, r)
    return ldf, forests, cfg, variances


@pytest.mark.parametrize('H,W,r,scale', [(240, 424, 2, 0.5), (120, 212, 1, 0.25), (97, 131, 2, 1.0)])
def test_layered_fused_matches_oracle_and_unfused(tmp_path, H, W, r, scale):
    import torch
    from rdf_b200 import synth
    from rdf_b200.buffers import GpuBuffer
    from oracle import numpy_oracle as no
    depth = synth.depth_frames('live-mask', 1, H, W, seed=21)
    ldf, forests, cfg, _ = _layered(tmp_path, (H, W), r)
    assert ldf.num_layered_classes == 11 and ldf.labels_dims == (H // r, W // r)
    depth_buf = GpuBuffer((1, H, W), np.uint16)
    depth_buf.cu().set(depth)
    labels_buf = GpuBuffer((1, H // r, W // r), np.uint16)
    labels_buf.cu().fill(4242)                                   # run() must overwrite every pixel
    ldf.run(depth_buf, labels_buf, scale)
    torch.cuda.synchronize()
    comp = labels_buf.cu().get()[0]
    layers = [b.cu().get() for b in ldf.label_images]
    exp_comp, exp_layers = no.layered_run(forests, [(None, None), (0, 1)], cfg['conditions'], depth[0], r, scale)
    assert np.array_equal(comp, exp_comp)
    for a, b in zip(layers, exp_layers):
        assert np.array_equal(a, b)
    assert len(np.unique(comp)) > 4                              # several finger classes + 65535 present
    # the reference's launch sequence through the same ABI (fills + per-layer eval with filter + composite)
    labels2 = GpuBuffer((1, H // r, W // r), np.uint16)
    ldf.run_unfused(depth_buf, labels2, scale)
    torch.cuda.synchronize()
    assert np.array_equal(labels2.cu().get()[0], exp_comp)


def test_layered_matches_reference_kernels(tmp_path):
    import torch
    from rdf_b200 import synth
    from rdf_b200.buffers import GpuBuffer
    from oracle import ref_kernels as rk
    H, W, r, scale = 480, 848, 2, 1.0
    depth = synth.depth_frames('live-mask', 1, H, W)
    ldf, forests, cfg, variances = _layered(tmp_path, (H, W), r, max_depth=16)
    depth_buf = GpuBuffer((1, H, W), np.uint16)
    depth_buf.cu().set(depth)
    labels_buf = GpuBuffer((1, H // r, W // r), np.uint16)
    ldf.run(depth_buf, labels_buf, scale)
    ref_comp, ref_layers = rk.layered_run([to_dev(f) for f in forests], [(None, None), (0, 1)], cfg['conditions'],
                                          to_dev(depth[0]), r, scale)
    torch.cuda.synchronize()
    assert np.array_equal(labels_buf.cu().get()[0], to_np(ref_comp))
    for mine, ref in zip(ldf.label_images, ref_layers):
        assert np.array_equal(mine.cu().get(), to_np(ref))
    # mean shift on the composite: ours vs the reference kernel + its host loop, 1e-5 absolute, same NaN pattern
    from rdf_b200.mean_shift import MeanShift
    ms = MeanShift().run(6, labels_buf.cu(), ldf.num_layered_classes, variances)
    ref_ms = rk.mean_shift(ref_comp.reshape(1, H // r, W // r), ldf.num_layered_classes, variances, 6)
    assert np.array_equal(np.isnan(ms), np.isnan(ref_ms))
    assert np.nanmax(np.abs(ms - ref_ms)) <= 1e-5


@pytest.mark.parametrize('h,w,K', [(240, 424, 11), (33, 57, 3), (480, 848, 11), (720, 1280, 20), (7, 5, 2), (120, 160, 5), (200, 212, 30)])
def test_mean_shift_matches_oracle(h, w, K):
    from rdf_b200.mean_shift import MeanShift
    from oracle import numpy_oracle as no
    rng = np.random.default_rng(h * 1000 + w)
    labels = np.full((1, h, w), 65535, np.uint16)
    # blobs of classes, some background 0, class K left empty on purpose (-> NaN row)
    yy, xx = np.mgrid[0:h, 0:w]
    for k in range(1, K):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        rad = max(2, min(h, w) // 6)
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= rad * rad
        labels[0][m] = k
    labels[0][rng.random((h, w)) < 0.05] = 0
    variances = rng.uniform(4.0, 60.0, size=K).astype(np.float32)
    ms = MeanShift()
    for rounds in (1, 6):
        got = ms.run(rounds, to_dev(labels), K, variances)
        exp = no.mean_shift(labels, K, variances, rounds)
        assert got.shape == (K, 2) and got.dtype == np.float64
        assert np.array_equal(np.isnan(got), np.isnan(exp))
        assert np.isnan(exp[K - 1]).all()
        assert np.nanmax(np.abs(got - exp)) <= 1e-5
    # bitwise reproducible run to run
    again = ms.run(6, to_dev(labels), K, variances)
    assert np.array_equal(np.nan_to_num(again, nan=-1.0), np.nan_to_num(got, nan=-1.0))


def test_mean_shift_dense_labels():
    """every pixel labelled (maximum list size) and a single class"""
    from rdf_b200.mean_shift import MeanShift
    from oracle import c_oracle as co
    labels = np.ones((1, 480, 848), np.uint16)
    labels[0, :, 424:] = 2
    variances = np.array([30.0, 500.0], np.float32)
    got = MeanShift().run(6, to_dev(labels), 2, variances)
    exp = co.mean_shift(labels, 2, variances, 6)
    assert np.abs(got - exp).max() <= 1e-5


def test_mean_shift_ignores_labels_above_num_labels():
    """labels above num_labels: the reference kernel dereferences a null pointer for them (cu_utils.hpp:110-114, found when
    generating tests/golden); here they are ignored, like 0 and 65535.  Also the golden blobs fixture (reference output)."""
    import os
    from rdf_b200.mean_shift import MeanShift
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layered_meanshift.npz'))
    lab, var, want = z['blobs.labels'].copy(), z['blobs.variances'], z['blobs.means']
    got = MeanShift().run(5, to_dev(lab[None]), 4, var)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.nanmax(np.abs(got - want)) <= 1e-5
    lab[1, 0:5] = 9
    lab[2, 0:5] = 5
    got2 = MeanShift().run(5, to_dev(lab[None]), 4, var)
    assert np.array_equal(np.nan_to_num(got2, nan=-1.0), np.nan_to_num(got, nan=-1.0))


@pytest.mark.parametrize('name', ['layered_240x424_r2', 'layered_97x131_r1'])
def test_layered_golden_fixture(tmp_path, name):
    """CUDA path against the committed reference-kernel outputs (tests/golden/layered_meanshift.npz)."""
    import os
    import torch
    from rdf_b200 import decision_tree as dt
    from rdf_b200.buffers import GpuBuffer
    from rdf_b200.mean_shift import MeanShift
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layered_meanshift.npz'))
    depth = z[f'{name}.depth']
    r, scale = int(z[f'{name}.params'][0]), float(z[f'{name}.params'][1])
    H, W = depth.shape[1:]
    layers = []
    for i in range(2):
        f = z[f'{name}.forest{i}']
        m = dt.DecisionForest(f.shape[0], int(np.log2(f.shape[1] + 1)), (f.shape[2] - 7) // 2)
        m.forest_cu.set(f)
        layers.append(m)
    cfg = {'layers': [{'model': layers[0]}, {'model': layers[1], 'filter_model': 0, 'filter_model_class': 1}],
           'conditions': z[f'{name}.conditions'].tolist(), 'label_colors': [[0, 0, 0, 255]] * 11, 'root': '.'}
    ldf = dt.LayeredDecisionForest(cfg, (H, W), r)
    d = GpuBuffer((1, H, W), np.uint16)
    d.cu().set(depth)
    out = GpuBuffer((1, H // r, W // r), np.uint16)
    ldf.run(d, out, scale)
    torch.cuda.synchronize()
    assert np.array_equal(out.cu().get()[0], z[f'{name}.composite'])
    assert np.array_equal(ldf.label_images[0].cu().get(), z[f'{name}.layer0'])
    assert np.array_equal(ldf.label_images[1].cu().get(), z[f'{name}.layer1'])
    means = MeanShift().run(6, out.cu(), 11, z[f'{name}.variances'])
    want = z[f'{name}.means']
    assert np.array_equal(np.isnan(means), np.isnan(want)) and np.nanmax(np.abs(means - want)) <= 1e-5


def test_composite_standalone_untouched_pixels():
    import torch
    from rdf_b200 import decision_tree as dt
    from oracle import numpy_oracle as no
    rng = np.random.default_rng(3)
    h, w = 50, 70
    l1 = rng.choice(np.array([0, 1, 2, 65535], np.uint16), size=(h, w))
    l2 = rng.choice(np.array([0, 1, 2, 3, 65535], np.uint16), size=(h, w))
    cond = np.array([[1, 2], [0, 7], [0, 1], [0, 2], [0, 3]], np.int32)
    a, b = to_dev(l1), to_dev(l2)
    ptrs = to_dev(np.array([a.data_ptr(), b.data_ptr()], np.int64))
    comp = filled_u16((1, h, w), 999)
    dt.DecisionTreeEvaluator().make_composite_labels_image(ptrs, w, h, to_dev(cond), comp)
    torch.cuda.synchronize()
    exp = np.full((h, w), 999, np.uint16)
    no.composite([l1, l2], cond, exp)
    assert np.array_equal(to_np(comp)[0], exp)
