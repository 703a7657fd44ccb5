#!/usr/bin/env python
"""cfg4 level step on N GPUs (torchrun), three exchange modes.  Images sharded over the ranks:
  allreduce: bucket + zero + rdf_train_hist_bucketed + NCCL allreduce of the whole histogram + rdf_train_pick_best (all features)
  p2p      : bucket + zero + barrier + rdf_train_hist_bucketed_p2p (reduce-scatter fused into the flush over NVLink) + barrier +
             rdf_train_pick_candidates (own feature slice) + all-gather of the per-node winners + rdf_train_pick_finalize
Dataset replicated on every rank (42 frames = 34 MB):
  features : bucket (all pixels) + zero + rdf_train_hist_bucketed for the rank's OWN feature slice (complete histograms, nothing
             crosses NVLink) + rdf_train_pick_candidates + all-gather of the per-node winners + rdf_train_pick_finalize
Time = max over ranks (CUDA events).  Checks that all modes write the same node records."""
import argparse
import ctypes
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def run(frames=42, features=2000, levels=(0, 8, 12), iters=3, emit=None):
    """Needs an initialised process group with world_size > 1 (bench.py's multi-GPU leg and main() below).  Returns one dict per
    level on every rank (times are max over ranks)."""
    import types
    args = types.SimpleNamespace(frames=frames, features=features, iters=iters)
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    from rdf_b200 import _capi, synth, dist as rdist
    rank, world, local = rdist.env_rank_world()
    lib = _capi.load()
    results = []
    H, W, C, F, NT, D = 480, 848, 4, args.features, 64, 16
    n0, n1 = rdist.shard_range(args.frames, rank, world)
    N = n1 - n0
    depth = torch.from_numpy(synth.depth_frames('dense-smooth', N, H, W, first_frame=n0).view(np.int16)).cuda()
    labels_all = synth.train_labels(args.frames, H, W)
    labels = torch.from_numpy(labels_all[n0:n1].view(np.int16).copy()).cuda()
    NF = args.frames                                                        # 'features' mode: the whole dataset on every rank
    depth_full = torch.from_numpy(synth.depth_frames('dense-smooth', NF, H, W).view(np.int16)).cuda()
    labels_full = torch.from_numpy(labels_all.view(np.int16).copy()).cuda()
    off_np, th_np = synth.random_proposals(F, NT)
    offsets, thresholds = torch.from_numpy(off_np).cuda(), torch.from_numpy(th_np).cuda()
    st = _capi.stream_ptr
    E = 7 + 2 * C
    Fo = (F + world - 1) // world
    for level in levels:
        S = 1 << level
        nodes_all = synth.random_node_assignment(labels_all, level)
        nodes = torch.from_numpy(np.ascontiguousarray(nodes_all[n0:n1])).cuda()
        slot = torch.arange(S, dtype=torch.int32, device='cuda')
        active = torch.arange(S, dtype=torch.int32, device='cuda')
        parent = torch.zeros((1 << D, C), dtype=torch.int64, device='cuda')      # global parent counts (all ranks' pixels)
        na = torch.from_numpy(nodes_all.reshape(-1).astype(np.int64))
        la = torch.from_numpy(labels_all.reshape(-1).astype(np.int64))
        parent.view(-1).index_add_(0, (na * C + la).cuda(), torch.ones(na.numel(), dtype=torch.int64, device='cuda'))
        need = ctypes.c_size_t()
        _capi.check(lib.rdf_train_bucket_workspace_bytes(N * H * W, S, ctypes.byref(need)))
        ws = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')
        hist = torch.zeros((S, F, NT + 1, C), dtype=torch.int32, device='cuda')
        nodes_full = torch.from_numpy(np.ascontiguousarray(nodes_all)).cuda()
        _capi.check(lib.rdf_train_bucket_workspace_bytes(NF * H * W, S, ctypes.byref(need)))
        ws_full = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')
        need_full = need.value
        _capi.check(lib.rdf_train_bucket_workspace_bytes(N * H * W, S, ctypes.byref(need)))
        f_lo = min(F, rank * Fo)
        nloc = max(0, min(F, (rank + 1) * Fo) - f_lo)
        hist_feat = torch.zeros((S, max(nloc, 1), NT + 1, C), dtype=torch.int32, device='cuda')
        sym = symm_mem.empty(S * Fo * (NT + 1) * C, dtype=torch.int32, device=torch.device('cuda', local))
        hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
        ptrs = torch.tensor([int(x) for x in hdl.buffer_ptrs], dtype=torch.int64, device='cuda')
        cg = torch.zeros((S,), dtype=torch.float32, device='cuda')
        ci = torch.zeros((S,), dtype=torch.int32, device='cuda')
        cc = torch.zeros((S, 2, C), dtype=torch.int64, device='cuda')
        ag = torch.empty((world, S), dtype=torch.float32, device='cuda')
        ai = torch.empty((world, S), dtype=torch.int32, device='cuda')
        ac = torch.empty((world, S, 2, C), dtype=torch.int64, device='cuda')

        def fresh():
            return (torch.zeros(((1 << D) - 1, E), dtype=torch.float32, device='cuda'), torch.zeros_like(parent),
                    torch.full((1 << D,), -1.0, dtype=torch.float32, device='cuda'))

        def step_allreduce(tree, nxt, gain):
            _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes), N * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
            hist.zero_()
            _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(ws), S, _capi.dptr(offsets),
                                                    _capi.dptr(thresholds), F, NT, C, _capi.dptr(hist), st()))
            if world > 1:
                dist.all_reduce(hist)
            _capi.check(lib.rdf_train_pick_best(S, _capi.dptr(active), _capi.dptr(slot), _capi.dptr(parent), _capi.dptr(hist), S,
                                                _capi.dptr(offsets), _capi.dptr(thresholds), F, NT, C, level, D, _capi.dptr(tree),
                                                _capi.dptr(nxt), _capi.dptr(gain), st()))

        def step_p2p(tree, nxt, gain):
            _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes), N * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
            sym.zero_()
            hdl.barrier(channel=0)
            _capi.check(lib.rdf_train_hist_bucketed_p2p(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(ws), S, _capi.dptr(offsets),
                                                        _capi.dptr(thresholds), F, NT, C, _capi.dptr(ptrs), world, st()))
            hdl.barrier(channel=0)
            nloc = max(0, min(F, (rank + 1) * Fo) - rank * Fo)
            _capi.check(lib.rdf_train_pick_candidates(S, _capi.dptr(active), _capi.dptr(slot), _capi.dptr(parent), _capi.dptr(sym), S, nloc, Fo,
                                                      rank * Fo, NT, C, _capi.dptr(cg), _capi.dptr(ci), _capi.dptr(cc), st()))
            if world > 1:
                dist.all_gather_into_tensor(ag, cg)
                dist.all_gather_into_tensor(ai, ci)
                dist.all_gather_into_tensor(ac, cc)
            else:
                ag.copy_(cg[None]); ai.copy_(ci[None]); ac.copy_(cc[None])
            _capi.check(lib.rdf_train_pick_finalize(S, _capi.dptr(active), _capi.dptr(slot), _capi.dptr(parent), world, _capi.dptr(ag),
                                                    _capi.dptr(ai), _capi.dptr(ac), _capi.dptr(offsets), _capi.dptr(thresholds), NT, C, level, D,
                                                    _capi.dptr(tree), _capi.dptr(nxt), _capi.dptr(gain), st()))

        def step_features(tree, nxt, gain):
            _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes_full), NF * H * W, _capi.dptr(slot), S, _capi.dptr(ws_full), need_full, st()))
            hist_feat.zero_()
            if nloc > 0:
                _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth_full), _capi.dptr(labels_full), NF, W, H, _capi.dptr(ws_full), S,
                                                        _capi.dptr(offsets[f_lo:f_lo + nloc]), _capi.dptr(thresholds[f_lo:f_lo + nloc]), nloc, NT, C,
                                                        _capi.dptr(hist_feat), st()))
                _capi.check(lib.rdf_train_pick_candidates(S, _capi.dptr(active), _capi.dptr(slot), _capi.dptr(parent), _capi.dptr(hist_feat), S, nloc,
                                                          nloc, f_lo, NT, C, _capi.dptr(cg), _capi.dptr(ci), _capi.dptr(cc), st()))
            else:
                cg.fill_(-1.0); ci.fill_(0x7fffffff); cc.zero_()
            dist.all_gather_into_tensor(ag, cg)
            dist.all_gather_into_tensor(ai, ci)
            dist.all_gather_into_tensor(ac, cc)
            _capi.check(lib.rdf_train_pick_finalize(S, _capi.dptr(active), _capi.dptr(slot), _capi.dptr(parent), world, _capi.dptr(ag),
                                                    _capi.dptr(ai), _capi.dptr(ac), _capi.dptr(offsets), _capi.dptr(thresholds), NT, C, level, D,
                                                    _capi.dptr(tree), _capi.dptr(nxt), _capi.dptr(gain), st()))

        out = {}
        trees = {}
        for name, fn in (('allreduce', step_allreduce), ('p2p', step_p2p), ('features', step_features)):
            t, nx, g = fresh()
            fn(t, nx, g)
            torch.cuda.synchronize()
            rdist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                g.fill_(-1.0)
                fn(t, nx, g)
            e1.record()
            torch.cuda.synchronize()
            out[name] = rdist.max_over_ranks(e0.elapsed_time(e1) / args.iters)
            trees[name] = (t, nx)
        same = bool(torch.equal(trees['allreduce'][0].view(torch.int32), trees['p2p'][0].view(torch.int32)) and
                    torch.equal(trees['allreduce'][1], trees['p2p'][1]) and
                    torch.equal(trees['features'][0].view(torch.int32), trees['p2p'][0].view(torch.int32)) and
                    torch.equal(trees['features'][1], trees['p2p'][1]))
        same = bool(rdist.max_over_ranks(0.0 if same else 1.0) == 0.0)          # on every rank
        digest = hashlib.md5(trees['p2p'][0].cpu().numpy().tobytes() + trees['p2p'][1].cpu().numpy().tobytes()).hexdigest()
        px = args.frames * H * W
        rec = {'cfg4_level': level, 'active_nodes': S, 'n_gpus': world, 'ms_per_level_allreduce': out['allreduce'],
               'ms_per_level_p2p': out['p2p'], 'ms_per_level_features': out['features'], 'node_records_identical': same, 'records_md5': digest,
               'hist_GB': S * F * (NT + 1) * C * 4 / 1e9, 'g_feature_evals_per_s_p2p': px * F / out['p2p'] / 1e6}
        results.append(rec)
        if rank == 0 and emit is not None:
            emit(rec)
        del hist, sym, hdl, ws, ws_full, hist_feat, nodes_full
        torch.cuda.empty_cache()
    rdist.barrier()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=42)
    ap.add_argument('--features', type=int, default=2000)
    ap.add_argument('--levels', default='0,8,12')
    ap.add_argument('--iters', type=int, default=3)
    args = ap.parse_args()
    import torch.distributed as dist
    from rdf_b200 import dist as rdist
    rank, world, local = rdist.init_from_env()
    torch.cuda.set_device(local)
    run(args.frames, args.features, [int(x) for x in args.levels.split(',')], args.iters, emit=lambda r: print(json.dumps(r), flush=True))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
