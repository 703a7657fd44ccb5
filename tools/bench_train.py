#!/usr/bin/env python
"""cfg4 (BASELINE.json configs[3]) split-search sweep: 42 dense-smooth 848x480 frames = 17.1 M labelled pixels,
F candidate features x 64 sorted thresholds, C = 4; one full histogram pass at levels 0 / 4 / 8 / 12
(1 / 16 / 256 / 4096 active nodes, node ids from a hash of 32x32 pixel tiles).  Under torchrun the images are sharded over
the ranks and the histograms are NCCL sum-allreduced (the path's only exchange step); time = max over ranks.
Prints one JSON line per level on rank 0.  `--check` compares a feature sub-block with the C oracle (level 4, 2 frames)."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=42)
    ap.add_argument('--features', type=int, default=2000)
    ap.add_argument('--thresholds', type=int, default=64)
    ap.add_argument('--levels', default='0,4,8,12')
    ap.add_argument('--feature-block', type=int, default=0, help='features per rdf_train_hist call (0 = pick by memory)')
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--raster', action='store_true', help='use the un-bucketed rdf_train_hist kernel')
    args = ap.parse_args()
    from rdf_b200 import _capi, synth, dist as rdist
    import torch.distributed as dist
    rank, world, local = rdist.init_from_env()
    torch.cuda.set_device(local)
    lib = _capi.load()
    H, W, C, F, NT = 480, 848, 4, args.features, args.thresholds
    n0, n1 = rdist.shard_range(args.frames, rank, world)
    N = n1 - n0
    depth_np = synth.depth_frames('dense-smooth', N, H, W, first_frame=n0)
    labels_np = synth.train_labels(N, H, W, first_frame=n0)
    depth = torch.from_numpy(depth_np.view(np.int16)).cuda()
    labels = torch.from_numpy(labels_np.view(np.int16)).cuda()
    offsets_np, thresholds_np = synth.random_proposals(F, NT)
    offsets = torch.from_numpy(offsets_np).cuda()
    thresholds = torch.from_numpy(thresholds_np).cuda()
    st = _capi.stream_ptr
    for level in [int(x) for x in args.levels.split(',')]:
        S = 1 << level
        # all frames' node ids come from one generator call so that sharding does not change them
        nodes_all = synth.random_node_assignment(synth.train_labels(args.frames, H, W), level)
        nodes = torch.from_numpy(np.ascontiguousarray(nodes_all[n0:n1])).cuda()
        slot = torch.arange(S, dtype=torch.int32, device='cuda')
        per_feature = S * (NT + 1) * C * 4
        fb = args.feature_block or int(max(1, min(F, (6 << 30) // per_feature)))
        hist = torch.zeros((S, fb, NT + 1, C), dtype=torch.int32, device='cuda')
        blocks = [(f0, min(F, f0 + fb)) for f0 in range(0, F, fb)]

        need = ctypes.c_size_t()
        _capi.check(lib.rdf_train_bucket_workspace_bytes(N * H * W, S, ctypes.byref(need)))
        ws = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')

        def sweep(do_allreduce=True):
            if not args.raster:                              # counting sort of the active pixels by node: once per level
                _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes), N * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
            for f0, f1 in blocks:
                h = hist if f1 - f0 == fb else hist.view(-1)[:S * (f1 - f0) * (NT + 1) * C].view(S, f1 - f0, NT + 1, C)
                h.zero_()
                if args.raster:
                    _capi.check(lib.rdf_train_hist(_capi.dptr(depth), _capi.dptr(labels), _capi.dptr(nodes), N, W, H, _capi.dptr(slot), S,
                                                   _capi.dptr(offsets[f0:f1]), _capi.dptr(thresholds[f0:f1]), f1 - f0, NT, C, _capi.dptr(h), st()))
                else:
                    _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(ws), S,
                                                            _capi.dptr(offsets[f0:f1]), _capi.dptr(thresholds[f0:f1]), f1 - f0, NT, C,
                                                            _capi.dptr(h), st()))
                if do_allreduce and world > 1:
                    dist.all_reduce(h)
        sweep()
        torch.cuda.synchronize()
        rdist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            sweep()
        e1.record()
        torch.cuda.synchronize()
        ms = rdist.max_over_ranks(e0.elapsed_time(e1) / args.iters)
        ms_compute = None
        if world > 1:
            e0.record()
            for _ in range(args.iters):
                sweep(False)
            e1.record()
            torch.cuda.synchronize()
            ms_compute = rdist.max_over_ranks(e0.elapsed_time(e1) / args.iters)
        if rank == 0:
            px = args.frames * H * W
            b_alg = px * 8 + px * F * 8                      # 8 B/px/level + 8 B per (px x feature)
            print(json.dumps({'cfg4_level': level, 'active_nodes': S, 'n_gpus': world, 'ms_per_level': ms, 'ms_compute_only': ms_compute,
                              'g_feature_evals_per_s': px * F / ms / 1e6, 'feature_block': fb, 'hist_bytes_per_block': S * fb * (NT + 1) * C * 4,
                              'algorithmic_GBps': b_alg / ms / 1e6, 'kernel': 'raster' if args.raster else 'bucketed', 'labelled_pixels': px, 'features': F, 'thresholds': NT}), flush=True)
        del hist, nodes
    if args.check and rank == 0:
        from oracle import c_oracle as co
        level, S, nf = 4, 16, 8
        nodes_np = synth.random_node_assignment(synth.train_labels(args.frames, H, W), level)[n0:n0 + 2]
        hist = torch.zeros((S, nf, NT + 1, C), dtype=torch.int32, device='cuda')
        d2, l2, nd2 = depth[:2].contiguous(), labels[:2].contiguous(), torch.from_numpy(np.ascontiguousarray(nodes_np)).cuda()
        slot = torch.arange(S, dtype=torch.int32, device='cuda')
        need = ctypes.c_size_t()
        _capi.check(lib.rdf_train_bucket_workspace_bytes(2 * H * W, S, ctypes.byref(need)))
        ws = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')
        _capi.check(lib.rdf_train_bucket(_capi.dptr(nd2), 2 * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
        _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(d2), _capi.dptr(l2), 2, W, H, _capi.dptr(ws), S,
                                                _capi.dptr(offsets[:nf]), _capi.dptr(thresholds[:nf]), nf, NT, C, _capi.dptr(hist), st()))
        torch.cuda.synchronize()
        exp = co.train_hist(depth_np[:2], labels_np[:2], nodes_np, np.arange(S, dtype=np.int32), S, offsets_np[:nf], thresholds_np[:nf], C)
        print(json.dumps({'cfg4_check_vs_c_oracle': bool(np.array_equal(hist.cpu().numpy().view(np.uint32), exp))}), flush=True)
    rdist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
