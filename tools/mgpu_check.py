#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
  1. frame-sharded forest eval: every rank evaluates its shard; the gathered label maps equal rank 0's single-GPU result;
  2. image-sharded training, both exchange modes (reduction fused into the histogram kernel over peer-mapped buffers, and
     NCCL sum-allreduce of the split histograms): every rank ends with the tree a single GPU trains on the whole dataset
     (bit-identical canonical array).
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run():
    """Needs torch.distributed initialised when WORLD_SIZE > 1 (bench.py's multi-GPU leg, tests/test_gpu_multi.py, main() below).
    Returns the result dict on every rank (the parity flags are only meaningful on rank 0)."""
    from rdf_b200 import dist as rdist, synth
    from rdf_b200 import decision_tree as dt
    rank, world, local = rdist.env_rank_world()
    ev = dt.DecisionTreeEvaluator()

    # ---- 1. eval ----
    N, H, W, T, D, C = 8 * world + 3, 120, 160, 4, 10, 4
    forest_np = synth.random_forest(T, D, C, seed=5, ragged=True)
    forest = dt.DecisionForest(T, D, C)
    forest.forest_cu.set(forest_np)
    f0, f1 = rdist.shard_range(N, rank, world)
    depth = dt.cu_array.to_gpu(synth.depth_frames('dense-smooth', f1 - f0, H, W, seed=3, first_frame=f0))
    labels = dt.cu_array.GPUArray((f1 - f0, H, W), dtype=np.uint16).fill(65535)
    ev.get_labels_forest(forest, depth, labels)
    sizes = [rdist.shard_range(N, r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    pad = torch.zeros((maxn, H, W), dtype=torch.int16, device='cuda')
    pad[:f1 - f0] = labels.tensor.view(torch.int16)
    gathered = [torch.empty_like(pad) for _ in range(world)]
    if world > 1:                                            # NCCL has no int16: gather the bytes
        dist.all_gather([g.view(torch.uint8) for g in gathered], pad.view(torch.uint8))
    else:
        gathered = [pad]
    eval_ok = True
    if rank == 0:
        full_depth = dt.cu_array.to_gpu(synth.depth_frames('dense-smooth', N, H, W, seed=3))
        full = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16).fill(65535)
        ev.get_labels_forest(forest, full_depth, full)
        got = torch.cat([g[:b - a] for g, (a, b) in zip(gathered, sizes)])
        eval_ok = bool(torch.equal(got, full.tensor.view(torch.int16)))

    # ---- 2. training ----
    Nt, Ht, Wt, Dt, F, NT = 4 * world, 64, 96, 6, 32, 8
    props = {lvl: synth.random_proposals(F, NT, seed=100 + lvl) for lvl in range(Dt)}

    modes = {}

    def train(n0, n1, group_world, exchange=None):
        if exchange == 'features':                           # replicated dataset: every rank trains on ALL images
            n0, n1 = 0, Nt
        d = synth.depth_frames('dense-smooth', n1 - n0, Ht, Wt, seed=9, first_frame=n0)
        l = synth.train_labels(n1 - n0, Ht, Wt, first_frame=n0)
        ds = dt.DecisionTreeDatasetConfig.from_arrays(d, l, C)
        tr = dt.DecisionTreeTrainer(n1 - n0, F, thresholds_per_feature=NT, proposal_fn=lambda lvl, b: props[lvl],
                                    process_group=None if group_world > 1 else False, exchange=exchange)
        tr.allocate(ds, F, Dt)
        if group_world > 1:
            modes[str(exchange)] = ('features (replicated dataset, feature-sharded search, winners all-gathered)' if tr._feat is not None else
                                    'p2p (reduction fused into the histogram kernel)' if tr._p2p is not None else
                                    'nccl allreduce ' + getattr(tr, '_p2p_error', ''))
        tree = dt.DecisionTree(Dt, C)
        tr.train(ds, tree)
        torch.cuda.synchronize()
        return tree.tree_out_cu.get()

    i0, i1 = rdist.shard_range(Nt, rank, world)
    results = {}
    for exchange in ((None, 'allreduce', 'features') if world > 1 else (None,)):
        sharded = train(i0, i1, world, exchange)
        trees = [None] * world
        if world > 1:
            dist.all_gather_object(trees, sharded.tobytes())
        else:
            trees = [sharded.tobytes()]
        results[exchange] = trees
    train_ok = True
    if rank == 0:
        single = train(0, Nt, 1)
        train_ok = all(t == single.tobytes() for trees in results.values() for t in trees) and bool((single[:, 5:7] == -1).any())
    rdist.barrier()
    return {'world': world, 'eval_shards_match_single_gpu': eval_ok, 'sharded_training_matches_single_gpu': train_ok,
            'exchange_modes_tested': modes, 'frames': N, 'train_images': Nt}


def main():
    from rdf_b200 import dist as rdist
    rank, world, local = rdist.init_from_env()
    torch.cuda.set_device(local)
    res = run()
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and not (res['eval_shards_match_single_gpu'] and res['sharded_training_matches_single_gpu']):
        sys.exit(1)


if __name__ == '__main__':
    main()
