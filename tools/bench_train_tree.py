#!/usr/bin/env python
"""Whole-tree training: this repo's DecisionTreeTrainer (rdf_train_* through the C ABI) against the reference's own training
kernels (src/cuda/tree_train.cu compiled unchanged, driven by oracle/ref_kernels.train_tree with the reference's loop), same
images, same proposal stream (reference form: one threshold per proposal), identical resulting tree required.
    python tools/bench_train_tree.py [--frames 8] [--depth 12] [--proposals 64] [--blocks 4]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--depth', type=int, default=12)
    ap.add_argument('--proposals', type=int, default=64)
    ap.add_argument('--blocks', type=int, default=4)
    args = ap.parse_args()
    from conftest import to_dev
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from oracle import ref_kernels as rk
    N, H, W, C, D, P, B = args.frames, 480, 848, 4, args.depth, args.proposals, args.blocks
    depth = synth.depth_frames('dense-smooth', N, H, W)
    labels = synth.train_labels(N, H, W)
    rng = np.random.default_rng(7)
    stream = {lvl: [np.concatenate(synth.random_proposals(P, 1, seed=int(rng.integers(1 << 30))), axis=1).astype(np.float32)
                    for _ in range(B)] for lvl in range(D)}

    ds = dt.DecisionTreeDatasetConfig.from_arrays(depth, labels, C)
    trainer = dt.DecisionTreeTrainer(N, P, proposal_fn=lambda lvl, b: (stream[lvl][b][:, 0:4], stream[lvl][b][:, 4:5]))
    trainer.allocate(ds, P * B, D)
    tree = dt.DecisionTree(D, C)

    def ours():
        trainer.train(ds, tree)
        torch.cuda.synchronize()
    ours()
    t0 = time.perf_counter()
    ours()
    t_ours = time.perf_counter() - t0
    mine = tree.tree_out_cu.get()

    out = {'frames': N, 'pixels': N * H * W, 'depth': D, 'proposals_per_level': P * B, 'ours_s': t_ours,
           'nodes_split': int((mine[:, 5:7] == -1).any(axis=1).sum())}
    if rk.available():
        d_dev, l_dev = to_dev(depth), to_dev(labels)

        def ref():
            return rk.train_tree(d_dev, l_dev, C, D, lambda lvl: stream[lvl])
        ref()
        t0 = time.perf_counter()
        theirs = ref()
        t_ref = time.perf_counter() - t0
        out.update({'reference_kernels_s': t_ref, 'speedup': t_ref / t_ours,
                    'identical_split_records': bool(np.array_equal(mine[:, 0:7], theirs[:, 0:7])),
                    'max_pdf_diff': float(np.abs(mine[:, 7:] - theirs[:, 7:]).max())})
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
