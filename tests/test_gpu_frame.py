"""GPU parity of the live-frame conditioning: the fused kernels of csrc/rdf_frame.cu against the C
oracle (oracle/frame_oracle.c) and, live, against the reference's own kernels (oracle/_ref/libref_points.so); then the whole
product frame (HandsFramePipeline) against the reference's host sequence replayed with the oracles."""
import numpy as np
import pytest

from conftest import to_dev, to_np

pytestmark = pytest.mark.gpu


def _condition(scene, sigma=2.0, k=5, level=3, thresh=None, with_mm=True):
    import torch
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.buffers import GPUArray
    d = scene['depth_raw']
    H, W = d.shape
    ops = PointsOps()
    raw = GPUArray((H, W), dtype=np.uint16); raw.set(d)
    out = GPUArray((H, W), dtype=np.uint16); out.fill(7777)
    mm = GPUArray((H >> level, W >> level), dtype=np.uint16) if with_mm else None
    if mm is not None:
        mm.fill(7777)
    plane = GPUArray((4, 4), dtype=np.float32); plane.set(np.ascontiguousarray(scene['plane'], dtype=np.float32))
    ops.condition_depth(raw, out, mm, scene['pp'], scene['focal'], plane, scene['plane_z_threshold'] if thresh is None else thresh,
                        sigma, k, level)
    torch.cuda.synchronize()
    assert np.array_equal(raw.get(), d)                               # the raw frame stays intact for the read-out
    return out.get(), (mm.get() if mm is not None else None)


@pytest.mark.parametrize('H,W,seed,k,sigma,level', [(480, 848, 1234, 5, 2.0, 3), (120, 208, 31, 5, 2.0, 3), (96, 160, 32, 3, 0.7, 2),
                                                    (61, 101, 33, 5, 0.05, 2), (120, 208, 34, 7, 1.3, 3), (75, 93, 35, 9, 2.5, 1),
                                                    (33, 47, 36, 41, 9.0, 0), (1, 1, 37, 5, 2.0, 3), (5, 3, 38, 5, 2.0, 1),
                                                    (480, 848, 99, 5, 2.0, 4)])
def test_condition_depth_matches_oracle_and_reference_kernels(H, W, seed, k, sigma, level):
    from rdf_b200 import synth
    from oracle import frame_oracle as fo, ref_points as rp
    s = synth.live_scene(H, W, seed=seed)
    got, got_mm = _condition(s, sigma, k, level)
    exp, exp_mm = fo.condition_frame(s['depth_raw'], s['pp'], s['focal'], s['plane'], s['plane_z_threshold'], sigma, k, level)
    assert np.array_equal(got, exp)
    assert np.array_equal(got_mm, exp_mm)
    if True:                       # reference build required (tests/conftest.py)
        gk = fo.gaussian_kernel(k, sigma) if sigma > 0.1 else None
        ref, ref_mm = rp.condition_frame(s['depth_raw'], s['pp'], s['focal'], s['plane'], s['plane_z_threshold'], gk, level)
        assert np.array_equal(got, ref)
        assert np.array_equal(got_mm, ref_mm)
    if H >= 60:
        assert 0 < (got > 0).sum() < got.size


def test_condition_depth_matches_golden_fixture():
    import os
    from tests_golden_cases import CASES
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'frame.npz')
    if not os.path.exists(path):
        pytest.skip('tests/golden/frame.npz not generated yet')
    g = np.load(path)
    for name, H, W, seed, k, sigma, level, thresh in CASES:
        s = dict(depth_raw=g[name + '.depth_raw'], pp=g[name + '.pp'], focal=g[name + '.focal'], plane=g[name + '.plane'],
                 plane_z_threshold=thresh)
        got, got_mm = _condition(s, sigma, k, level)
        assert np.array_equal(got, g[name + '.depth']) and np.array_equal(got_mm, g[name + '.mm']), name


@pytest.mark.parametrize('seed', range(6))
def test_condition_depth_fuzz_planes_and_frames(seed):
    """Arbitrary matrices (incl. a last row that is not (0,0,0,1): w' != 1 keeps the sample, w' == 0 drops it), noise frames,
    extreme thresholds."""
    from oracle import frame_oracle as fo, ref_points as rp
    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.integers(20, 90)), int(rng.integers(20, 130))
    d = rng.integers(0, 65536, size=(H, W)).astype(np.uint16)
    d[rng.random((H, W)) < 0.3] = 0
    if seed == 1:
        d[:] = 0
    plane = rng.normal(size=(4, 4)).astype(np.float32)
    if seed % 3 == 0:
        plane[3] = [0, 0, 0, 1]
    if seed == 4:
        plane[3] = [0, 0, 0, 0]                                        # w' == 0 everywhere: every sample is "missing"
    if seed == 5:
        plane[3] = [0, 0, 1.0 / 4096, 0]                               # w' == 1 only where d == 4096
        d[rng.random((H, W)) < 0.3] = 4096
    thresh = [40.0, -1e30, 1e30, 0.0, 12.5, 300.0][seed]
    s = dict(depth_raw=d, pp=rng.normal(size=2).astype(np.float32) * 10 + [W / 2, H / 2], focal=np.float32(50 + 400 * rng.random()),
             plane=plane, plane_z_threshold=np.float32(thresh))
    s['pp'] = s['pp'].astype(np.float32)
    for sigma, k in [(2.0, 5), (0.0, 5), (1.1, 11)]:
        got, got_mm = _condition(s, sigma, k, 2)
        exp, exp_mm = fo.condition_frame(d, s['pp'], s['focal'], plane, thresh, sigma, k, 2)
        assert np.array_equal(got, exp) and np.array_equal(got_mm, exp_mm)
        if True:                   # reference build required (tests/conftest.py)
            ref, ref_mm = rp.condition_frame(d, s['pp'], s['focal'], plane, thresh, fo.gaussian_kernel(k, sigma) if sigma > 0.1 else None, 2)
            assert np.array_equal(got, ref) and np.array_equal(got_mm, ref_mm)


def test_gaussian_depth_filter_alone_matches_reference():
    import torch
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.buffers import GPUArray
    from oracle import frame_oracle as fo
    rng = np.random.default_rng(3)
    d = rng.integers(0, 9000, size=(77, 123)).astype(np.uint16)
    d[rng.random(d.shape) < 0.4] = 0
    a = GPUArray(d.shape, dtype=np.uint16); a.set(d)
    b = GPUArray(d.shape, dtype=np.uint16)
    PointsOps().gaussian_depth_filter(a, b, 1.7, 7)
    torch.cuda.synchronize()
    exp = np.zeros_like(d)
    fo.gaussian_depth_filter(d, exp, 1.7, 7)
    assert np.array_equal(b.get(), exp)


@pytest.mark.parametrize('H,W,seed,level', [(480, 848, 1234, 3), (120, 208, 31, 3), (61, 101, 33, 2), (96, 160, 32, 0)])
def test_grouping_grow_stencil_flip_match_oracle_and_reference(H, W, seed, level):
    import torch
    from rdf_b200 import synth
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.grouping import CppGrouping
    from rdf_b200.buffers import GPUArray
    from oracle import frame_oracle as fo, grouping_oracle as go, ref_points as rp
    s = synth.live_scene(H, W, seed=seed)
    depth, mm = fo.condition_frame(s['depth_raw'], s['pp'], s['focal'], s['plane'], s['plane_z_threshold'], 2.0, 5, level)
    if mm.size <= 16384:
        stencil, _ = go.make_groups(mm, 0.06)
    else:                                                              # level 0: the grouping kernel's size limit; any group image will do
        stencil = ((mm > 0) * (1 + (np.arange(W)[None, :] >= W // 2))).astype(np.uint16)
    ops = PointsOps()
    d_dev = GPUArray((H, W), dtype=np.uint16); d_dev.set(depth)
    st_dev = GPUArray(mm.shape, dtype=np.uint16); st_dev.set(stencil)
    if mm.size <= 16384:
        mm_dev = GPUArray(mm.shape, dtype=np.uint16); mm_dev.set(mm)
        g_info = GPUArray((2, 3), dtype=np.float32)
        st2 = GPUArray(mm.shape, dtype=np.uint16)
        CppGrouping().make_groups_cu(mm_dev, st2, g_info, 0.06)
        assert np.array_equal(st2.get(), stencil)
    grown_dev = GPUArray(mm.shape, dtype=np.uint16)
    ops.grow_groups(st_dev, grown_dev)
    grown = fo.grow_groups(stencil)
    assert np.array_equal(grown_dev.get(), grown)
    hands = [(1, False), (2, True), (2, False), (3, True)]
    out = GPUArray((len(hands), H, W), dtype=np.uint16); out.fill(1)
    ops.stencil_hands(d_dev, st_dev, level, hands, out, grow=True)
    out2 = GPUArray((2, H, W), dtype=np.uint16); out2.fill(1)
    ops.stencil_hands(d_dev, grown_dev, level, hands[:2], out2, grow=False)
    torch.cuda.synchronize()
    got = out.get()
    for i, (gid, flip) in enumerate(hands):
        exp = fo.hand_depth_image(depth, grown, level, gid, flip)
        assert np.array_equal(got[i], exp), (gid, flip)
        if i < 2:
            assert np.array_equal(got[i], rp.hand_depth_image(depth, grown, level, gid, flip))
    assert np.array_equal(out2.get(), got[:2])
    assert (got[0] != 65535).sum() > 0 and (got[1] != 65535).sum() > 0 and (got[3] != 65535).sum() == 0
    if True:                       # reference build required (tests/conftest.py)
        assert np.array_equal(grown, rp.grow_groups(stencil))
    # flip_x + display helpers
    lab = np.random.default_rng(seed).integers(0, 13, size=(H // 2, W // 2)).astype(np.uint16)
    lab[lab == 12] = 65535
    a = GPUArray(lab.shape, dtype=np.uint16); a.set(lab)
    b = GPUArray(lab.shape, dtype=np.uint16)
    ops.flip_x(a, b)
    assert np.array_equal(b.get(), fo.flip_x(lab)) and np.array_equal(b.get(), lab[:, ::-1])
    colors = np.random.default_rng(seed + 1).integers(0, 256, size=(11, 4)).astype(np.uint8)
    rgba0 = np.random.default_rng(seed + 2).integers(0, 256, size=lab.shape + (4,)).astype(np.uint8)
    c_dev = GPUArray(colors.shape, dtype=np.uint8); c_dev.set(colors)
    r_dev = GPUArray(rgba0.shape, dtype=np.uint8); r_dev.set(rgba0)
    ops.make_rgba_from_labels(a, c_dev, r_dev)
    exp_rgba = rgba0.copy()
    fo.make_rgba_from_labels(lab, colors, exp_rgba)
    assert np.array_equal(r_dev.get(), exp_rgba)
    dr = GPUArray((H, W, 4), dtype=np.uint8)
    ops.make_depth_rgba(d_dev, 2000, 6000, dr)
    assert np.array_equal(dr.get(), fo.make_depth_rgba(depth, 2000, 6000))
    if True:                       # reference build required (tests/conftest.py)
        assert np.array_equal(r_dev.get(), rp.make_rgba_from_labels(lab, colors, rgba0))
        assert np.array_equal(dr.get(), rp.make_depth_rgba(depth, 2000, 6000))


def test_fingertip_z_matches_oracle():
    import torch
    from rdf_b200 import synth
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.buffers import GPUArray
    from oracle import frame_oracle as fo
    s = synth.live_scene(480, 848, seed=8)
    rng = np.random.default_rng(8)
    K = 11
    means = np.stack([rng.random(K) * 430, rng.random(K) * 245], axis=1)
    means[3] = np.nan
    means[5] = [424.0, 10.0]          # x * 2 == 848: just outside
    means[6] = [-0.4, 239.9]          # truncates to (0, 239): inside
    means[7] = [1e12, 5.0]
    means[8] = [-3.0, 5.0]
    idx = [2, 3, 4, 5, 6, 7, 8, 9, 1, 11]
    ops = PointsOps()
    m_dev = GPUArray((K, 2), dtype=np.float64); m_dev.set(means)
    raw = GPUArray((480, 848), dtype=np.uint16); raw.set(s['depth_raw'])
    plane = GPUArray((4, 4), dtype=np.float32); plane.set(s['plane'])
    z_dev = GPUArray((len(idx),), dtype=np.float64)
    z_host = torch.empty(len(idx), dtype=torch.float64, pin_memory=True)
    m_host = torch.empty((K, 2), dtype=torch.float64, pin_memory=True)
    ops.fingertip_z(m_dev, idx, 2, raw, s['pp'], s['fx'], s['fy'], plane, z_dev)
    ops.fingertip_z(m_dev, idx, 2, raw, s['pp'], s['fx'], s['fy'], plane, z_host, means_copy=m_host)
    torch.cuda.synchronize()
    exp = fo.fingertip_z(means, idx, 2, s['depth_raw'], s['pp'], s['fx'], s['fy'], s['plane'])
    got = z_dev.get()
    assert np.array_equal(np.isnan(got), np.isnan(exp)) and 4 <= np.isnan(exp).sum() < len(idx)
    ok = ~np.isnan(exp)
    assert np.max(np.abs(got[ok] - exp[ok])) <= 1e-9 * np.max(np.abs(exp[ok]))      # tolerance: 1e-9 relative (fp64 sums of 4 terms)
    assert np.array_equal(z_host.numpy(), got, equal_nan=True)
    assert np.array_equal(m_host.numpy(), means, equal_nan=True)


def _expected_frame(scene, forests, cfg, variances, r, rounds, fingertips, scale=1.0):
    """src/3d_bz.py:159-260 + run_per_hand_pipeline(1, False), (2, True) replayed with the oracles."""
    from oracle import frame_oracle as fo, grouping_oracle as go, numpy_oracle as no, c_oracle as co
    depth, mm = fo.condition_frame(scene['depth_raw'], scene['pp'], scene['focal'], scene['plane'], scene['plane_z_threshold'])
    stencil, g_info = go.make_groups(mm, 0.06)
    grown = fo.grow_groups(stencil) if g_info[:, 0].sum() > 0 else np.zeros_like(stencil)
    K = max(c[1] for c in cfg['conditions'] if c[0] == 0)
    means, zs, labels = [], [], []
    for gid, flip in [(1, False), (2, True)]:
        hand = fo.hand_depth_image(depth, grown, 3, gid, flip)
        comp, _ = no.layered_run(forests, [(None, None), (0, 1)], cfg['conditions'], hand, r, scale)
        if flip:
            comp = fo.flip_x(comp)
        m = co.mean_shift(comp[None], K, variances, rounds)
        means.append(m)
        zs.append(fo.fingertip_z(m, fingertips, r, scene['depth_raw'], scene['pp'], scene['fx'], scene['fy'], scene['plane']))
        labels.append(comp)
    return np.stack(means), np.stack(zs), labels, depth, grown


@pytest.mark.parametrize('use_graph,batch,concurrent,num_hands,upload,fused_readout', [
    (True, True, False, 2, 'kernel', True), (False, True, False, 2, 'kernel', False), (True, False, True, 2, 'kernel', False),
    (False, False, False, 2, 'kernel', False), (True, True, False, 1, 'fused', True), (True, True, False, 0, 'kernel', True),
    (True, False, True, 1, 'fused', False), (True, True, False, 2, 'kernel', False)])
def test_hands_frame_pipeline_matches_reference_sequence(tmp_path, use_graph, batch, concurrent, num_hands, upload, fused_readout):
    import torch
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from rdf_b200.pipeline import HandsFramePipeline
    H, W, r = 480, 848, 2
    forests, cfg, variances = synth.layered_cfg2(seed=77, max_depth=12)
    path = synth.write_layered_model(str(tmp_path), forests, cfg)
    ldf = dt.LayeredDecisionForest.load(path, (H, W), r)
    scene = synth.live_scene(H, W, seed=4321, num_hands=max(num_hands, 1))
    if num_hands == 0:
        scene['depth_raw'] = np.where(scene['depth_raw'] > 0, np.uint16(60000), np.uint16(0)).astype(np.uint16)  # far below the table
    fingertips = (2, 3, 4, 5, 6)
    pipe = HandsFramePipeline(ldf, variances, scene['pp'], scene['focal'], scene['plane'], fx=scene['fx'], fy=scene['fy'],
                              fingertip_idxes=fingertips, use_graph=use_graph, batch_hands=batch, concurrent_hands=concurrent, upload=upload, fused_readout=fused_readout)
    for rep in range(2):                                              # replay twice: no state may leak between frames
        means, z = pipe.run(scene['depth_raw'])
    exp_means, exp_z, exp_labels, exp_depth, exp_grown = _expected_frame(scene, forests, cfg, variances, r, 6, fingertips)
    assert np.array_equal(pipe.depth_image.cu().get(), exp_depth)
    for i in range(2):
        assert np.array_equal(pipe.labels_image(i).get(), exp_labels[i]), i
    assert np.array_equal(np.isnan(means), np.isnan(exp_means))
    ok = ~np.isnan(exp_means)
    if ok.any():
        assert np.max(np.abs(means[ok] - exp_means[ok])) <= 1e-5      # centroid tolerance of north_star
    assert np.array_equal(np.isnan(z), np.isnan(exp_z))
    okz = ~np.isnan(exp_z)
    if okz.any():
        # a centroid within 1e-5 of an integer boundary could pick the neighbouring pixel; none does in these scenes
        assert np.max(np.abs(z[okz] - exp_z[okz])) <= 1e-6 * max(1.0, np.max(np.abs(exp_z[okz])))
    if num_hands == 2:
        assert ok[0].any() and ok[1].any() and okz.any()
    if num_hands == 0:
        assert not ok.any() and not okz.any()


@pytest.mark.parametrize('h,w,K,N', [(240, 424, 11, 2), (120, 212, 11, 3), (33, 57, 3, 2), (30, 53, 30, 4), (480, 848, 11, 2)])
def test_batched_mean_shift_equals_per_image(h, w, K, N):
    """rdf_mean_shift_batch: one launch over N label images == N single calls (the batched class-parallel kernel where the image
    size allows 128-bit rows per image, a loop of launches otherwise); the fused read-out == the separate one."""
    import torch
    from rdf_b200.mean_shift import MeanShift
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.buffers import GPUArray
    from oracle import numpy_oracle as no
    rng = np.random.default_rng(h * 1000 + w)
    labels = np.full((N, h, w), 65535, dtype=np.uint16)
    for n in range(N):
        m = rng.random((h, w)) < 0.12
        labels[n][m] = rng.integers(0, K + 2, size=int(m.sum())).astype(np.uint16)       # 0 and K+1 are ignored
    variances = (4.0 + 10.0 * rng.random(K)).astype(np.float32)
    L = GPUArray(labels.shape, dtype=np.uint16); L.set(labels)
    batched = MeanShift().run_async(5, L, K, variances, batch=True).get()
    single = MeanShift()
    for n in range(N):
        one = single.run(5, L[n], K, variances)
        exp = no.mean_shift(labels[n], K, variances, 5)
        assert np.array_equal(np.isnan(one), np.isnan(exp)) and np.nanmax(np.abs(one - exp)) <= 1e-5
        assert np.array_equal(np.isnan(batched[n]), np.isnan(exp)) and np.nanmax(np.abs(batched[n] - exp)) <= 1e-5
    # fused read-out against the separate kernel on the same centroids
    r = 2
    raw = rng.integers(0, 9000, size=(h * r, w * r)).astype(np.uint16)
    raw_dev = GPUArray(raw.shape, dtype=np.uint16); raw_dev.set(raw)
    plane = GPUArray((4, 4), dtype=np.float32); plane.set(rng.normal(size=(4, 4)).astype(np.float32))
    idx = [1, K, 2, 1] if K >= 2 else [1]
    z_fused = GPUArray((N, len(idx)), dtype=np.float64)
    z_sep = GPUArray((N, len(idx)), dtype=np.float64)
    mc = GPUArray((N, K, 2), dtype=np.float64)
    ms = MeanShift()
    means = ms.run_fingertips_async(5, L, K, variances, idx, r, raw_dev, (w * r / 2 + 0.5, h * r / 2 - 0.25), 400.0, 401.0, plane, z_fused,
                                    means_copy=mc, batch=True)
    PointsOps().fingertip_z(means, idx, r, raw_dev, (w * r / 2 + 0.5, h * r / 2 - 0.25), 400.0, 401.0, plane, z_sep)
    torch.cuda.synchronize()
    assert np.array_equal(z_fused.get(), z_sep.get(), equal_nan=True)
    assert np.array_equal(mc.get(), means.get(), equal_nan=True) and np.array_equal(means.get(), batched, equal_nan=True)


@pytest.mark.parametrize('H,W,r,N,scale', [(120, 212, 2, 3, 0.5), (97, 131, 1, 2, 1.0)])
def test_batched_layered_equals_per_image(tmp_path, H, W, r, N, scale):
    """rdf_layered_run_batch over N images with per-image mirroring == N single runs (+ flip_x of the composite)."""
    import torch
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from rdf_b200.buffers import GPUArray
    forests, cfg, _ = synth.layered_cfg2(seed=5, max_depth=9)
    ldf = dt.LayeredDecisionForest.load(synth.write_layered_model(str(tmp_path), forests, cfg), (H, W), r)
    depth = np.concatenate([synth.depth_frames('live-mask', 1, H, W, seed=40 + n) for n in range(N)])
    flips = [bool(n & 1) for n in range(N)]
    h, w = H // r, W // r
    d = GPUArray((N, H, W), dtype=np.uint16); d.set(depth)
    comp = GPUArray((N, h, w), dtype=np.uint16); comp.fill(1234)
    layers = [GPUArray((N, h, w), dtype=np.uint16) for _ in range(ldf.num_models)]
    ldf.run(d, comp, scale, composite_flip_x=flips, label_images=layers)
    got, got_layers = comp.get(), [l.get() for l in layers]
    one = GPUArray((1, h, w), dtype=np.uint16)
    for n in range(N):
        ldf.run(d[n], one, scale)
        torch.cuda.synchronize()
        exp = one.get()[0]
        assert np.array_equal(got[n], exp[:, ::-1] if flips[n] else exp), n
        for li in range(ldf.num_models):
            assert np.array_equal(got_layers[li][n], ldf.label_images[li].cu().get()), (n, li)     # per-layer images stay unmirrored


def test_hands_frame_pipeline_matches_reference_kernels_and_host_sequence():
    """The whole product frame against the REFERENCE's own kernels (points_ops.cu, calibrated_plane.cu, tree_eval.cu, mean_shift.cu
    compiled unchanged) and C++ flood fill, launched in the order and with the host round trips of src/3d_bz.py."""
    import os
    import sys
    from oracle import ref_points as rp, ref_kernels as rk, grouping_oracle as go
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import bench_hands_frame as b
    pipe, scene, forests, cfg, variances = b.build(depth=12, seed=99)
    means, z = pipe.run(scene['depth_raw'])
    _, (ref_means, ref_z) = b.reference_sequence(scene, forests, cfg, variances, iters=1)
    assert np.array_equal(np.isnan(means), np.isnan(ref_means))
    ok = ~np.isnan(ref_means)
    assert ok.any() and np.max(np.abs(means[ok] - ref_means[ok])) <= 1e-5
    assert np.array_equal(np.isnan(z), np.isnan(ref_z))
    okz = ~np.isnan(ref_z)
    assert okz.any() and np.max(np.abs(z[okz] - ref_z[okz])) <= 1e-6 * max(1.0, np.max(np.abs(ref_z[okz])))


@pytest.mark.parametrize('seed', range(8))
def test_condition_and_stencil_fuzz_shapes(seed):
    """Random small shapes, window sizes, reduction levels and hand lists against the C oracle."""
    import torch
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.buffers import GPUArray
    from rdf_b200 import synth
    from oracle import frame_oracle as fo
    rng = np.random.default_rng(5000 + seed)
    H, W = int(rng.integers(1, 140)), int(rng.integers(1, 200))
    if seed % 2:
        W = (W + 7) // 8 * 8                                            # the 128-bit stencil path
    level = int(rng.integers(0, 4))
    k = int(rng.choice([1, 3, 5, 7, 9, 13]))
    sigma = float(rng.choice([0.05, 0.5, 1.0, 2.0, 3.3]))
    s = synth.live_scene(max(H, 2), max(W, 2), seed=seed)
    d = np.ascontiguousarray(s['depth_raw'][:H, :W])
    d[rng.random(d.shape) < 0.1] = 0
    scene = dict(s, depth_raw=d)
    got, got_mm = _condition(scene, sigma, k, level)
    exp, exp_mm = fo.condition_frame(d, s['pp'], s['focal'], s['plane'], s['plane_z_threshold'], sigma, k, level)
    assert np.array_equal(got, exp) and np.array_equal(got_mm, exp_mm)
    gh, gw = H >> level, W >> level
    if gh == 0 or gw == 0:
        return
    groups = rng.integers(0, 4, size=(gh, gw)).astype(np.uint16)
    groups[rng.random(groups.shape) < 0.6] = 0
    hands = [(int(rng.integers(0, 4)), bool(rng.integers(0, 2))) for _ in range(int(rng.integers(1, 5)))]
    ops = PointsOps()
    d_dev = GPUArray((H, W), dtype=np.uint16); d_dev.set(exp)
    g_dev = GPUArray((gh, gw), dtype=np.uint16); g_dev.set(groups)
    out = GPUArray((len(hands), H, W), dtype=np.uint16); out.fill(3)
    ops.stencil_hands(d_dev, g_dev, level, hands, out, grow=True)
    torch.cuda.synchronize()
    grown = fo.grow_groups(groups)
    for i, (gid, flip) in enumerate(hands):
        assert np.array_equal(out.get()[i], fo.hand_depth_image(exp, grown, level, gid, flip)), (i, gid, flip)


@pytest.mark.parametrize('H,W,level', [(61, 101, 2), (120, 208, 3), (37, 64, 1), (1, 1, 0), (480, 848, 3)])
def test_frame_kernels_do_not_write_outside_their_outputs(H, W, level):
    """Every output sits between two canary regions of one larger allocation (compute-sanitizer is not available on the GPU pool)."""
    import torch
    from rdf_b200 import synth
    from rdf_b200.points_ops import PointsOps
    from rdf_b200.grouping import CppGrouping
    from rdf_b200.buffers import GPUArray
    G = 4096                                                     # canary elements on each side

    def guarded(shape, np_dtype, canary):
        n = int(np.prod(shape))
        whole = GPUArray((n + 2 * G,), dtype=np_dtype)
        whole.fill(canary)
        mid = GPUArray(tuple(shape), tensor=whole.tensor[G:G + n].view(*shape))
        return whole, mid, n

    def intact(whole, n, canary):
        a = whole.get()
        return bool((a[:G] == canary).all() and (a[G + n:] == canary).all())

    s = synth.live_scene(max(H, 2), max(W, 2), seed=H + W)
    d = np.ascontiguousarray(s['depth_raw'][:H, :W])
    ops = PointsOps()
    raw = GPUArray((H, W), dtype=np.uint16); raw.set(d)
    plane = GPUArray((4, 4), dtype=np.float32); plane.set(s['plane'])
    gh, gw = H >> level, W >> level
    w_out, out, n_out = guarded((H, W), np.uint16, 0xABCD)
    w_mm, mm, n_mm = guarded((max(gh, 1), max(gw, 1)), np.uint16, 0xABCD)
    mm_arg = mm if gh * gw > 0 else None
    ops.condition_depth(raw, out, mm_arg, s['pp'], s['focal'], plane, s['plane_z_threshold'], 2.0, 5, level)
    torch.cuda.synchronize()
    assert intact(w_out, n_out, 0xABCD) and intact(w_mm, n_mm, 0xABCD)
    if gh * gw == 0:
        return
    w_st, st, n_st = guarded((gh, gw), np.uint16, 0xABCD)
    w_gi, gi, n_gi = guarded((2, 3), np.float32, -7.0)
    if gh * gw <= 16384:
        CppGrouping().make_groups_cu(mm, st, gi, 0.01)
    else:
        st.fill(1)
    w_gr, gr, n_gr = guarded((gh, gw), np.uint16, 0xABCD)
    ops.grow_groups(st, gr)
    w_h, hands, n_h = guarded((3, H, W), np.uint16, 0xABCD)
    ops.stencil_hands(out, st, level, [(1, False), (2, True), (1, True)], hands, grow=True)
    w_f, fl, n_f = guarded((H, W), np.uint16, 0xABCD)
    ops.flip_x(out, fl)
    torch.cuda.synchronize()
    for whole, n, c in [(w_st, n_st, 0xABCD), (w_gi, n_gi, -7.0), (w_gr, n_gr, 0xABCD), (w_h, n_h, 0xABCD), (w_f, n_f, 0xABCD), (w_out, n_out, 0xABCD),
                        (w_mm, n_mm, 0xABCD)]:
        assert intact(whole, n, c)
    assert not (hands.get() == 0xABCD).all()
