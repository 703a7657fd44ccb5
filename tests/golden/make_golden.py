#!/usr/bin/env python
"""Generate tests/golden/*.npz: outputs of the REFERENCE's own kernels (src/cuda/tree_eval.cu, tree_train.cu,
mean_shift.cu compiled unchanged for sm_100a -> oracle/_ref/libref_kernels.so) on small seeded inputs.

The reference ships no golden vectors, so these fixtures are what pins the CPU oracle: they are produced
by the reference's kernels with the reference's launch geometry on a B200 and re-checked on CPU by
tests/test_oracle_golden.py.  Every fixture stores its INPUTS as well as the reference outputs, so a later change of
rdf_b200/synth.py cannot silently move the goalposts.

Run on a GPU box (the reference .so travels with the snapshot; /root/reference is not needed at run time):
    gpurun -- python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def main(out_dir):
    import torch
    from conftest import to_dev, to_np, filled_u16
    from rdf_b200 import synth
    from oracle import ref_kernels as rk
    assert torch.cuda.is_available() and rk.available()
    os.makedirs(out_dir, exist_ok=True)
    meta = dict(gpu=torch.cuda.get_device_name(0), generator='tests/golden/make_golden.py',
                source='reference kernels compiled unchanged for sm_100a (oracle/ref_kernels/ref_launch.cu)')

    # ---- forest eval: evaluate_image_using_forest (tree_eval.cu:24-137) ----
    cases = [
        # name, kind, N, H, W, T, D, C, ragged, r, scale, with_filter, seed
        ('smooth_t3d8c4', 'dense-smooth', 2, 60, 80, 3, 8, 4, False, 1, 1.0, False, 11),
        ('noise_t4d7c11_ragged', 'dense-noise', 2, 48, 64, 4, 7, 11, True, 1, 1.0, False, 12),
        ('mask_t3d8c3_r2_s05', 'live-mask', 1, 96, 128, 3, 8, 3, True, 2, 0.5, False, 13),
        ('smooth_t2d6c5_r3_s037_filter', 'dense-smooth', 2, 63, 90, 2, 6, 5, True, 3, 0.37, True, 14),
        ('noise_t1d9c2', 'dense-noise', 1, 50, 70, 1, 9, 2, False, 1, 1.0, False, 15),
        ('smooth_t8d5c4', 'dense-smooth', 1, 40, 56, 8, 5, 4, True, 1, 2.0, False, 16),
    ]
    out = {}
    for name, kind, N, H, W, T, D, C, ragged, r, scale, with_filter, seed in cases:
        depth = synth.depth_frames(kind, N, H, W, seed=seed)
        depth[0, 3:6, 4:9] = 0                      # centre depth 0 -> skipped; also a probed 0 is used as 0
        depth[0, 10:12, 10:14] = 65535              # centre 65535 -> skipped
        forest = synth.random_forest(T, D, C, seed=seed, ragged=ragged)
        h, w = H // r, W // r
        filt = None
        fclass = None
        if with_filter:
            rng = np.random.default_rng(seed)
            filt = rng.integers(0, 3, size=(N, h, w)).astype(np.uint16)
            fclass = 1
        labels = filled_u16((N, h, w), 65535)
        rk.eval_forest(to_dev(forest), to_dev(depth), labels, r, to_dev(filt) if filt is not None else None, fclass, scale)
        torch.cuda.synchronize()
        out[f'{name}.depth'] = depth
        out[f'{name}.forest'] = forest
        out[f'{name}.params'] = np.array([r, scale, 1 if with_filter else 0, fclass if fclass is not None else -1], dtype=np.float64)
        if filt is not None:
            out[f'{name}.filter'] = filt
        out[f'{name}.labels'] = to_np(labels)
    metak = {'meta.' + k: np.array(v) for k, v in meta.items()}          # provenance travels inside every fixture
    np.savez_compressed(os.path.join(out_dir, 'eval_forest.npz'), names=np.array([c[0] for c in cases]), **out, **metak)

    # ---- single tree: evaluate_image_using_tree (tree_eval.cu:140-212) ----
    out = {}
    names = []
    for name, kind, H, W, D, C, ragged, seed in [('tree_d7c4', 'dense-smooth', 50, 66, 7, 4, False, 21),
                                                 ('tree_d8c3_ragged', 'dense-noise', 44, 60, 8, 3, True, 22),
                                                 ('tree_d4c4_falls_off', 'dense-smooth', 30, 40, 4, 4, False, 23)]:
        depth = synth.depth_frames(kind, 2, H, W, seed=seed)
        depth[1, 5:7, 5:9] = 0
        tree = synth.random_forest(1, D, C, seed=seed, ragged=ragged)[0]
        if 'falls_off' in name:
            tree[:, 5:7] = -1.0                      # no leaf anywhere: walks fall off the last level, nothing is written
        labels = filled_u16((2, H, W), 65535)
        rk.eval_tree(to_dev(tree), to_dev(depth), labels)
        torch.cuda.synchronize()
        names.append(name)
        out[f'{name}.depth'] = depth
        out[f'{name}.tree'] = tree
        out[f'{name}.labels'] = to_np(labels)
    np.savez_compressed(os.path.join(out_dir, 'eval_tree.npz'), names=np.array(names), **out, **metak)

    # ---- layered run + composite (decision_tree.py:233-264, tree_eval.cu:214-248) and mean shift ----
    out = {}
    names = []
    for name, H, W, r, scale, D, seed in [('layered_240x424_r2', 240, 424, 2, 0.5, 8, 31), ('layered_97x131_r1', 97, 131, 1, 1.0, 7, 32)]:
        depth = synth.depth_frames('live-mask', 1, H, W, seed=seed)
        forests, cfg, variances = synth.layered_cfg2(seed=seed, max_depth=D)
        comp, imgs = rk.layered_run([to_dev(f) for f in forests], [(None, None), (0, 1)], cfg['conditions'], to_dev(depth[0]), r, scale)
        torch.cuda.synchronize()
        means = rk.mean_shift(comp.reshape(1, H // r, W // r), 11, variances, 6)
        names.append(name)
        out[f'{name}.depth'] = depth
        out[f'{name}.forest0'] = forests[0]
        out[f'{name}.forest1'] = forests[1]
        out[f'{name}.conditions'] = np.asarray(cfg['conditions'], dtype=np.int32)
        out[f'{name}.params'] = np.array([r, scale], dtype=np.float64)
        out[f'{name}.composite'] = to_np(comp)
        out[f'{name}.layer0'] = to_np(imgs[0])
        out[f'{name}.layer1'] = to_np(imgs[1])
        out[f'{name}.variances'] = variances
        out[f'{name}.means'] = means
    # mean shift on a synthetic label image with an empty class (NaN row) and out-of-range labels
    rng = np.random.default_rng(5)
    lab = np.full((60, 90), 65535, np.uint16)
    for k, (cx, cy) in enumerate([(20, 15), (60, 40), (45, 20)]):
        pts = np.round(rng.normal((cx, cy), 4.0, size=(300, 2))).astype(int)
        ok = (pts[:, 0] >= 0) & (pts[:, 0] < 90) & (pts[:, 1] >= 0) & (pts[:, 1] < 60)
        lab[pts[ok, 1], pts[ok, 0]] = [1, 2, 4][k]           # class index 2 (label 3) stays empty
    lab[0, 0:5] = 0
    # NOTE: labels above num_labels must NOT be fed to the reference kernel: Array2d::get_ptr returns nullptr for them
    # (cu_utils.hpp:110-114) and the atomicAdd faults.  Our kernel ignores such labels; that case is tested separately.
    var = np.array([3.0, 5.0, 4.0, 6.0], dtype=np.float32)
    out['blobs.labels'] = lab
    out['blobs.variances'] = var
    out['blobs.means'] = rk.mean_shift(to_dev(lab).reshape(1, 60, 90), 4, var, 5)
    np.savez_compressed(os.path.join(out_dir, 'layered_meanshift.npz'), names=np.array(names), **out, **metak)

    # ---- training: evaluate_random_features histograms of one level + whole-tree training (tree_train.cu) ----
    out = {}
    N, H, W, C = 2, 48, 64, 4
    depth = synth.depth_frames('dense-smooth', N, H, W, seed=41)
    depth[0, 4:8, 4:8] = 0                                   # compute_feature returns 0.0 for d == 0 (reachable here)
    labels = synth.train_labels(N, H, W)
    labels[1, 20:24, :] = 0
    level = 3
    nodes = synth.random_node_assignment(labels, level, seed=7)
    P = 24
    off, th = synth.random_proposals(P, 1, seed=43)
    proposals = np.concatenate([off, th], axis=1).astype(np.float32)
    n_children = 1 << (level + 1)
    counts = torch.zeros((P, n_children, C), dtype=torch.int64, device='cuda')
    L = rk.lib()
    labels_d, depth_d, props_d, nodes_d = to_dev(labels), to_dev(depth), to_dev(proposals), to_dev(nodes)   # keep alive
    rk._ok(L.ref_evaluate_random_features(N, W, H, P, C, 6, n_children, 0, n_children, rk._p(labels_d), rk._p(depth_d),
                                          rk._p(props_d), rk._p(nodes_d), rk._p(counts), rk._st()))
    torch.cuda.synchronize()
    out['hist.depth'] = depth
    out['hist.labels'] = labels
    out['hist.nodes'] = nodes
    out['hist.proposals'] = proposals
    out['hist.counts'] = counts.cpu().numpy().astype(np.uint64)      # [P][child][class]

    D, blocks = 6, 2
    rng = np.random.default_rng(77)
    stream = {}
    for lvl in range(D):
        stream[lvl] = []
        for _ in range(blocks):
            o, t = synth.random_proposals(32, 1, seed=int(rng.integers(1 << 30)))
            stream[lvl].append(np.concatenate([o, t], axis=1).astype(np.float32))
    tree = rk.train_tree(to_dev(depth), to_dev(labels), C, D, lambda lvl: stream[lvl])
    out['train.depth'] = depth
    out['train.labels'] = labels
    out['train.proposals'] = np.stack([np.stack(stream[lvl]) for lvl in range(D)])   # [D][blocks][P][5]
    out['train.tree'] = tree
    np.savez_compressed(os.path.join(out_dir, 'train.npz'), **out, **metak)
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'golden'))
