#!/usr/bin/env python
"""Latency of one WHOLE product frame (src/3d_bz.py tick + two run_per_hand_pipeline calls): raw 848x480 camera frame in pinned
host memory -> conditioning -> hand grouping -> per hand (stencil, 2-layer stacked forest, 6-round mean shift, fingertip depths)
-> centroids and fingertip z of both hands in pinned host memory.  One CUDA-graph replay per frame.  Beside it: the reference's own
kernels in the reference's order with its host choreography (D2H -> C++ flood fill -> H2D, per-round mean-shift transfers) on the
same GPU, and a parity check of exactly what was timed.   python tools/bench_hands_frame.py [--iters 1000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)


def build(H=480, W=848, r=2, depth=16, seed=1234, **kw):
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from rdf_b200.pipeline import HandsFramePipeline
    forests, cfg, variances = synth.layered_cfg2(max_depth=depth)
    cfg = dict(cfg, layers=[dict(l) for l in cfg['layers']])
    for layer, f in zip(cfg['layers'], forests):
        m = dt.DecisionForest(f.shape[0], depth, (f.shape[2] - 7) // 2)
        m.forest_cu.set(f)
        layer['model'] = m
    cfg['root'] = '.'
    ldf = dt.LayeredDecisionForest(cfg, (H, W), r)
    scene = synth.live_scene(H, W, seed=seed)
    pipe = HandsFramePipeline(ldf, variances, scene['pp'], scene['focal'], scene['plane'], fx=scene['fx'], fy=scene['fy'], **kw)
    return pipe, scene, forests, cfg, variances


def percentiles(a):
    a = np.asarray(a)
    return {'p50_us': float(np.percentile(a, 50)), 'p95_us': float(np.percentile(a, 95)), 'p99_us': float(np.percentile(a, 99))}


def measure(pipe, scene, iters, warm):
    import torch
    pipe.run(scene['depth_raw'])
    plain, devt = [], []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        pipe.submit()
        pipe.stream.synchronize()
        t1 = time.perf_counter()
        if i >= warm:
            plain.append((t1 - t0) * 1e6)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warm + iters):
        with torch.cuda.stream(pipe.stream):
            e0.record()
        pipe.submit()
        with torch.cuda.stream(pipe.stream):
            e1.record()
        pipe.stream.synchronize()
        if i >= warm:
            devt.append(e0.elapsed_time(e1) * 1e3)
    out = percentiles(plain)
    out['device_p50_us'] = float(np.percentile(devt, 50))
    out['device_p99_us'] = float(np.percentile(devt, 99))
    return out


def stages(pipe, n=300):
    import ctypes
    import torch
    from rdf_b200 import _capi

    def stage(fn):
        with torch.cuda.stream(pipe.stream):
            fn()
        pipe.stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=pipe.stream):
            fn()
        with torch.cuda.stream(pipe.stream):
            for _ in range(20):
                g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                g.replay()
            b.record()
        pipe.stream.synchronize()
        return a.elapsed_time(b) * 1e3 / n

    def upload():
        _capi.check(_capi.load().rdf_upload_frame(ctypes.c_void_p(pipe.depth_host.data_ptr()), _capi.dptr(pipe.depth_raw.cu()),
                                                  pipe.depth_host.numel() * 2, _capi.stream_ptr()))
    p = pipe
    return {
        'upload_us': stage(upload),
        'condition_from_host_us': stage(lambda: p.ops.condition_depth(p.depth_host, p.depth_image, p.depth_image_mm, p.pp, p.focal, p.plane,
                                                                      p.thresh, p.sigma, p.k_size, p.mm_level)),
        'condition_us': stage(lambda: p.ops.condition_depth(p.depth_raw, p.depth_image, p.depth_image_mm, p.pp, p.focal, p.plane, p.thresh,
                                                            p.sigma, p.k_size, p.mm_level)),
        'group_hands_us': stage(lambda: p.grouping.make_groups_cu(p.depth_image_mm, p.depth_image_mm_groups_2, p.g_info, p.group_min_size)),
        'stencil_hands_us': stage(lambda: p.ops.stencil_hands(p.depth_image, p.depth_image_mm_groups_2, p.mm_level, p.hands,
                                                              p.depth_image_hands, grow=True)),
        'layered_both_hands_us': stage(lambda: p.ldf.run(p.depth_image_hands, p.labels_images, p.scale, composite_flip_x=[f for _, f in p.hands],
                                                         label_images=p.layer_images)),
        'mean_shift_both_hands_us': stage(lambda: p.mean_shift[0].run_async(p.rounds, p.labels_images.cu(), p.K, p.variances, batch=True)),
        'layered_one_hand_us': stage(lambda: p.ldf.run(p.depth_image_hands.cu()[0], p.labels_images.cu()[0], p.scale, composite_flip_x=[False],
                                                       label_images=[b.cu()[0] for b in p.layer_images])),
        'mean_shift_one_hand_us': stage(lambda: p.mean_shift[1].run_async(p.rounds, p.labels_images.cu()[0], p.K, p.variances)),
        'mean_shift_fingertips_both_hands_us': stage(lambda: p.mean_shift[0].run_fingertips_async(
            p.rounds, p.labels_images.cu(), p.K, p.variances, p.fingertips, p.ldf.labels_reduce, p.depth_raw, p.pp, p.fx, p.fy, p.plane,
            p.z_host, means_copy=p.means_host, batch=True)),
        'fingertip_z_us': stage(lambda: p.ops.fingertip_z(p.mean_shift[0].means, p.fingertips, p.ldf.labels_reduce, p.depth_raw, p.pp, p.fx,
                                                          p.fy, p.plane, p.z_host, means_copy=p.means_host)),
        'note': 'each stage replayed alone as a 1-node CUDA graph, back to back; includes per-graph launch latency',
    }


def expected(scene, forests, cfg, variances, r=2, rounds=6, fingertips=(2, 3, 4, 5, 6)):
    from oracle import frame_oracle as fo, grouping_oracle as go, numpy_oracle as no, c_oracle as co
    depth, mm = fo.condition_frame(scene['depth_raw'], scene['pp'], scene['focal'], scene['plane'], scene['plane_z_threshold'])
    stencil, g_info = go.make_groups(mm, 0.06)
    grown = fo.grow_groups(stencil)
    K = max(c[1] for c in cfg['conditions'] if c[0] == 0)
    means, zs, labels = [], [], []
    for gid, flip in [(1, False), (2, True)]:
        hand = fo.hand_depth_image(depth, grown, 3, gid, flip)
        comp, _ = no.layered_run(forests, [(None, None), (0, 1)], cfg['conditions'], hand, r, 1.0)
        if flip:
            comp = fo.flip_x(comp)
        m = co.mean_shift(comp[None], K, variances, rounds)
        means.append(m)
        zs.append(fo.fingertip_z(m, fingertips, r, scene['depth_raw'], scene['pp'], scene['fx'], scene['fy'], scene['plane']))
        labels.append(comp)
    return np.stack(means), np.stack(zs), labels


def reference_sequence(scene, forests, cfg, variances, iters=20, r=2, rounds=6, fingertips=(2, 3, 4, 5, 6)):
    """The reference's own kernels + host choreography for the same frame (src/3d_bz.py:156-260,387-522), wall time per frame."""
    import torch
    from oracle import ref_points as rp, ref_kernels as rk, grouping_oracle as go, frame_oracle as fo
    if not (rp.available() and rk.available() and go.ref_available()):
        return None
    import ctypes
    L = rp.lib()
    H, W = scene['depth_raw'].shape
    level, mh, mw = 3, H >> 3, W >> 3
    dev = 'cuda'
    f_dev = [torch.from_numpy(f).to(dev) for f in forests]
    K = max(c[1] for c in cfg['conditions'] if c[0] == 0)
    gk = torch.from_numpy(fo.gaussian_kernel(5, 2.0)).to(dev)
    plane_h = np.ascontiguousarray(scene['plane'], dtype=np.float32)
    depth_host = torch.from_numpy(scene['depth_raw'].view(np.int16)).pin_memory()
    depth = torch.zeros((H, W), dtype=torch.int16, device=dev)
    depth_2, group = torch.zeros_like(depth), torch.zeros_like(depth)
    pts = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    mm = torch.zeros((mh, mw), dtype=torch.int16, device=dev)
    groups_2, groups = torch.zeros_like(mm), torch.zeros_like(mm)
    labels_2 = torch.zeros((H // r, W // r), dtype=torch.int16, device=dev)
    st, p, f32 = rp._st, rp._p, rp._f
    pp, focal, thresh = scene['pp'], scene['focal'], scene['plane_z_threshold']
    times, last = [], None
    for it in range(iters + 3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        depth.copy_(depth_host)
        L.ref_deproject_points(W, H, f32(pp[0]), f32(pp[1]), f32(focal), p(depth), p(pts), st())
        L.ref_transform_points(W * H, p(pts), plane_h.ctypes.data_as(ctypes.c_void_p), st())
        L.ref_filter_points_by_plane(W * H, f32(thresh), p(pts), st())
        L.ref_remove_missing(W * H, p(pts), p(depth), st())
        depth_2.copy_(depth)
        L.ref_gaussian_depth_filter(W, H, 5, p(gk), p(depth_2), p(depth), st())
        L.ref_shrink_image(W, H, level, p(depth), p(mm), st())
        mm_cpu = mm.cpu().numpy().view(np.uint16)                                   # D2H + sync
        coords, g_info = go.ref_make_groups(mm_cpu, 0.06)                           # the reference's C++ flood fill
        groups_2.zero_()
        if len(coords):
            stencil = go.stencil_from_coords(coords, mh, mw)                        # stands in for the H2D of the coordinate list +
            groups_2.copy_(torch.from_numpy(stencil.view(np.int16)))               # write_pixel_groups_to_stencil_image
            L.ref_grow_groups(mw, mh, p(groups_2), p(groups), st())
        else:
            groups.zero_()
        out_means, out_z = [], []
        for gid, flip in [(1, False), (2, True)]:
            group.zero_()
            L.ref_stencil_depth_image_by_group(W, H, level, gid, p(groups), p(depth), p(group), st())
            if flip:
                L.ref_flip_x(W, H, p(group), p(depth_2), st())
            else:
                depth_2.copy_(group)
            L.ref_convert_0s_to_maxuint(W * H, p(depth_2), st())
            comp, _ = rk.layered_run(f_dev, [(None, None), (0, 1)], cfg['conditions'], depth_2.view(torch.uint16), r, 1.0)
            if flip:
                labels_2.copy_(comp.view(torch.int16))
                L.ref_flip_x(W // r, H // r, p(labels_2), p(comp), st())
            m = rk.mean_shift(comp.reshape(1, H // r, W // r), K, variances, rounds)
            out_means.append(m)
            out_z.append(fo.fingertip_z(m, fingertips, r, scene['depth_raw'], pp, scene['fx'], scene['fy'], scene['plane']))
        torch.cuda.synchronize()
        if it >= 3:
            times.append((time.perf_counter() - t0) * 1e6)
        last = (np.stack(out_means), np.stack(out_z))
    return percentiles(times), last


def run(iters=1000, per_hand=None, upload='kernel', with_ref=True, fused_readout=True):
    pipe, scene, forests, cfg, variances = build(batch_hands=per_hand is None, concurrent_hands=per_hand == 'concurrent', upload=upload,
                                                 fused_readout=fused_readout)
    means, z = pipe.run(scene['depth_raw'])
    res = {'workload': 'whole product frame: raw 848x480 frame -> plane clip + 5x5 zero-aware gaussian + 1/8 image -> hand grouping -> '
                       'per hand: stencil(+mirror), L1 (T3 D16 C3) -> L2 (T3 D16 C11), labels_reduce 2, mean shift 6 rounds x 11 classes, '
                       '5 fingertip depths; one CUDA-graph replay per frame',
           'e2e_host_frame': measure(pipe, scene, iters, min(100, iters)),
           'h2d_bytes': pipe.h2d_bytes, 'd2h_bytes': pipe.d2h_bytes, 'kernels_per_frame': pipe.kernels_per_frame, 'upload': upload,
           'hands': per_hand or 'batched', 'stages': stages(pipe)}
    exp_means, exp_z, exp_labels = expected(scene, forests, cfg, variances)
    ok = ~np.isnan(exp_means)
    res['parity'] = {
        'labels_bit_exact_vs_oracle': bool(all(np.array_equal(pipe.labels_image(i).get(), exp_labels[i]) for i in range(2))),
        'centroids_within_1e-5': bool(np.array_equal(np.isnan(means), np.isnan(exp_means)) and np.max(np.abs(means[ok] - exp_means[ok])) <= 1e-5),
        'fingertip_z_max_abs_err': float(np.nanmax(np.abs(z - exp_z))) if np.isfinite(exp_z).any() else None,
        'fingertip_nan_pattern_equal': bool(np.array_equal(np.isnan(z), np.isnan(exp_z))),
        'classes_found_per_hand': [int(np.isfinite(means[i, :, 0]).sum()) for i in range(2)],
        'fingertips_found_per_hand': [int(np.isfinite(z[i]).sum()) for i in range(2)],
        'evaluated_pixels_per_hand': [int((pipe.depth_image_hands.cu().get()[i][::2, ::2] != 65535).sum()) for i in range(2)],
    }
    if with_ref:
        try:
            ref = reference_sequence(scene, forests, cfg, variances)
            if ref is not None:
                t, (rm, rz) = ref
                okr = ~np.isnan(rm)
                res['reference_kernels_same_gpu'] = dict(t, what='reference kernels (points_ops.cu, calibrated_plane.cu, tree_eval.cu, mean_shift.cu '
                                                         'compiled unchanged) + grouping.cpp, launched in the order and with the host round trips of '
                                                         'src/3d_bz.py; wall time per frame',
                                                         centroids_agree_1e_5=bool(np.array_equal(np.isnan(rm), np.isnan(means)) and
                                                                                   np.max(np.abs(rm[okr] - means[okr])) <= 1e-5),
                                                         fingertip_z_max_abs_diff=float(np.nanmax(np.abs(rz - z))) if np.isfinite(rz).any() else None)
        except Exception as e:
            res['reference_kernels_same_gpu'] = {'error': repr(e)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=1000)
    ap.add_argument('--no-ref', action='store_true')
    ap.add_argument('--upload', default='kernel', choices=['kernel', 'fused'])
    ap.add_argument('--separate-readout', action='store_true')
    ap.add_argument('--per-hand', choices=['concurrent', 'sequential'], help='one launch per hand and stage instead of batched hands')
    args = ap.parse_args()
    print(json.dumps({'hands_frame': run(args.iters, args.per_hand, args.upload, not args.no_ref, not args.separate_readout)}))


if __name__ == '__main__':
    main()
