#!/usr/bin/env python
"""Generate tests/golden/frame.npz: outputs of the REFERENCE's own frame kernels (src/cuda/points_ops.cu, calibrated_plane.cu
compiled unchanged for sm_100a against oracle/ref_kernels/glm_min -> oracle/_ref/libref_points.so) and of its C++ flood fill
(oracle/_ref/libref_grouping.so), run in the order of src/3d_bz.py:159-260,390-456 on small seeded scenes.  Inputs are stored with
the outputs.  Run on a GPU box:   gpurun -- python tests/golden/make_golden_frame.py gpurun_out/golden
then copy gpurun_out/golden/frame.npz into tests/golden/ and commit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

# name, H, W, seed, k_size, sigma, mm_level, plane_z_threshold
CASES = [
    ('k5_l3', 120, 208, 31, 5, 2.0, 3, 40.0),
    ('k3_l2', 96, 160, 32, 3, 0.7, 2, 25.0),
    ('nofilter_odd_l2', 61, 101, 33, 5, 0.05, 2, 40.0),
    ('k7_l3', 120, 208, 34, 7, 1.3, 3, 60.0),
]


def main(out_dir):
    import torch
    from rdf_b200 import synth
    from oracle import ref_points as rp, grouping_oracle as go, frame_oracle as fo
    assert torch.cuda.is_available() and rp.available() and go.ref_available()
    os.makedirs(out_dir, exist_ok=True)
    out = {}
    for name, H, W, seed, k, sigma, level, thresh in CASES:
        s = synth.live_scene(H, W, seed=seed)
        gk = fo.gaussian_kernel(k, sigma) if sigma > 0.1 else None
        depth, mm = rp.condition_frame(s['depth_raw'], s['pp'], s['focal'], s['plane'], thresh, gk, level)
        coords, g_info = go.ref_make_groups(mm, 0.06)
        stencil = go.stencil_from_coords(coords, *mm.shape)
        grown = rp.grow_groups(stencil)
        hands = np.stack([rp.hand_depth_image(depth, grown, level, 1, False), rp.hand_depth_image(depth, grown, level, 2, True)])
        rng = np.random.default_rng(seed)
        labels = rng.integers(0, 13, size=(H // 2, W // 2)).astype(np.uint16)
        labels[labels == 12] = 65535
        colors = rng.integers(0, 256, size=(11, 4)).astype(np.uint8)
        rgba0 = rng.integers(0, 256, size=(H // 2, W // 2, 4)).astype(np.uint8)
        pre = name + '.'
        out.update({pre + 'depth_raw': s['depth_raw'], pre + 'pp': s['pp'], pre + 'focal': s['focal'], pre + 'plane': s['plane'],
                    pre + 'thresh': np.float32(thresh), pre + 'k': np.int32(k), pre + 'sigma': np.float64(sigma),
                    pre + 'level': np.int32(level), pre + 'gauss': gk if gk is not None else np.zeros((0, 0), np.float32),
                    pre + 'depth': depth, pre + 'mm': mm, pre + 'stencil': stencil, pre + 'g_info': g_info, pre + 'grown': grown,
                    pre + 'hands': hands, pre + 'labels': labels, pre + 'colors': colors, pre + 'rgba0': rgba0,
                    pre + 'labels_flipped': rp.flip_x(labels), pre + 'rgba': rp.make_rgba_from_labels(labels, colors, rgba0),
                    pre + 'mm_rgba': rp.make_depth_rgba(grown, 0, 2), pre + 'depth_rgba': rp.make_depth_rgba(depth, 2000, 6000)})
        print(name, 'kept', int((depth > 0).sum()), 'groups', g_info[:, 0])
    np.savez_compressed(os.path.join(out_dir, 'frame.npz'), names=np.array([c[0] for c in CASES]),
                        **{'meta.gpu': np.array(torch.cuda.get_device_name(0)), 'meta.generator': np.array('tests/golden/make_golden_frame.py'),
                           'meta.source': np.array('reference points_ops.cu / calibrated_plane.cu compiled unchanged for sm_100a '
                                                   '(oracle/ref_kernels/ref_points.cu) + grouping.cpp compiled by g++ '
                                                   '(oracle/ref_kernels/ref_grouping.cpp)')}, **out)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'golden'))
