#!/usr/bin/env python
"""Whole-tree training: this repo's DecisionTreeTrainer (rdf_train_* through the C ABI) against the reference's own training
kernels (src/cuda/tree_train.cu compiled unchanged, driven by oracle/ref_kernels.train_tree with the reference's loop), same
images, same proposal stream (reference form: one threshold per proposal), identical resulting tree required.
    python tools/bench_train_tree.py [--frames 8] [--depth 12] [--proposals 64] [--blocks 4]
--thresholds NT > 1 switches to the cfg4 form (F features x NT sorted thresholds per block; the reference has no such form, so no
reference run):  python tools/bench_train_tree.py --frames 42 --depth 16 --proposals 2000 --blocks 1 --thresholds 64"""
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
    ap.add_argument('--thresholds', type=int, default=1)
    ap.add_argument('--no-ref', action='store_true')
    args = ap.parse_args()
    from conftest import to_dev
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from oracle import ref_kernels as rk
    N, H, W, C, D, P, B = args.frames, 480, 848, 4, args.depth, args.proposals, args.blocks
    depth = synth.depth_frames('dense-smooth', N, H, W)
    labels = synth.train_labels(N, H, W)
    rng = np.random.default_rng(7)
    NT = args.thresholds
    if NT == 1:
        stream = {lvl: [np.concatenate(synth.random_proposals(P, 1, seed=int(rng.integers(1 << 30))), axis=1).astype(np.float32)
                        for _ in range(B)] for lvl in range(D)}
        fn = lambda lvl, b: (stream[lvl][b][:, 0:4], stream[lvl][b][:, 4:5])
    else:
        wide = {lvl: [synth.random_proposals(P, NT, seed=int(rng.integers(1 << 30))) for _ in range(B)] for lvl in range(D)}
        fn = lambda lvl, b: wide[lvl][b]

    ds = dt.DecisionTreeDatasetConfig.from_arrays(depth, labels, C)
    trainer = dt.DecisionTreeTrainer(N, P, thresholds_per_feature=NT, proposal_fn=fn)
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

    out = {'frames': N, 'pixels': N * H * W, 'depth': D, 'proposals_per_level': P * B, 'thresholds_per_feature': NT, 'ours_s': t_ours,
           'ms_per_level': t_ours * 1e3 / D, 'g_candidate_evals_per_s': N * H * W * P * B * D / t_ours / 1e9,
           'nodes_split': int((mine[:, 5:7] == -1).any(axis=1).sum())}
    if rk.available() and NT == 1 and not args.no_ref:
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
