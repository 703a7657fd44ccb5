#!/usr/bin/env python
"""Per-phase device time of one training level at cfg4 width (42 frames, F features x 64 thresholds, C = 4):
bucket (counting sort by node) | histogram | pick-best | next-active | advance-pixels.  One JSON line per level."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=42)
    ap.add_argument('--features', type=int, default=2000)
    ap.add_argument('--levels', default='0,8,12')
    ap.add_argument('--offset-scale', type=float, default=1.0, help='multiply the feature offsets (1e6: every probe leaves the image, every pixel lands in one bin: worst case for colliding shared-memory updates)')
    args = ap.parse_args()
    from rdf_b200 import _capi, synth
    lib = _capi.load()
    N, H, W, C, F, NT, D = args.frames, 480, 848, 4, args.features, 64, 16
    depth = torch.from_numpy(synth.depth_frames('dense-smooth', N, H, W).view(np.int16)).cuda()
    labels_np = synth.train_labels(N, H, W)
    labels = torch.from_numpy(labels_np.view(np.int16)).cuda()
    off_np, th_np = synth.random_proposals(F, NT)
    off_np = (off_np * np.float32(args.offset_scale)).astype(np.float32)
    offsets, thresholds = torch.from_numpy(off_np).cuda(), torch.from_numpy(th_np).cuda()
    st = _capi.stream_ptr
    E = 7 + 2 * C
    for level in [int(x) for x in args.levels.split(',')]:
        S = 1 << level
        nodes = torch.from_numpy(synth.random_node_assignment(labels_np, level)).cuda()
        slot = torch.arange(S, dtype=torch.int32, device='cuda')
        active = torch.arange(S, dtype=torch.int32, device='cuda')
        # parent counts per node from the node assignment (what the previous level's pick-best would have written)
        flat_nodes, flat_labels = nodes.view(-1).long(), labels.view(-1).long()
        parent = torch.zeros((1 << D, C), dtype=torch.int64, device='cuda')
        parent.view(-1).index_add_(0, flat_nodes * C + flat_labels, torch.ones_like(flat_nodes))
        next_counts = torch.zeros_like(parent)
        best_gain = torch.full((1 << D,), -1.0, dtype=torch.float32, device='cuda')
        tree = torch.zeros(((1 << D) - 1, E), dtype=torch.float32, device='cuda')
        next_active = torch.zeros((1 << D,), dtype=torch.int32, device='cuda')
        num_next = torch.zeros((1,), dtype=torch.int32, device='cuda')
        fb = int(max(1, min(F, (6 << 30) // (S * (NT + 1) * C * 4))))
        hist = torch.zeros((S, fb, NT + 1, C), dtype=torch.int32, device='cuda')
        need = ctypes.c_size_t()
        _capi.check(lib.rdf_train_bucket_workspace_bytes(N * H * W, S, ctypes.byref(need)))
        ws = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        t = {k: 0.0 for k in ('bucket', 'zero', 'hist', 'pick_best', 'next_active', 'advance')}
        for rep in range(2):                                  # second repetition is the timed one
            nodes_work = nodes.clone()
            torch.cuda.synchronize()
            ev[0].record()
            _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes_work), N * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
            ev[1].record()
            tz = th = tp = 0.0
            for f0 in range(0, F, fb):
                f1 = min(F, f0 + fb)
                h = hist if f1 - f0 == fb else hist.view(-1)[:S * (f1 - f0) * (NT + 1) * C].view(S, f1 - f0, NT + 1, C)
                a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
                a.record()
                h.zero_()
                b.record()
                _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(ws), S,
                                                        _capi.dptr(offsets[f0:f1]), _capi.dptr(thresholds[f0:f1]), f1 - f0, NT, C, _capi.dptr(h), st()))
                c.record()
                _capi.check(lib.rdf_train_pick_best(S, _capi.dptr(active), _capi.dptr(slot), _capi.dptr(parent), _capi.dptr(h), S,
                                                    _capi.dptr(offsets[f0:f1]), _capi.dptr(thresholds[f0:f1]), f1 - f0, NT, C, level, D,
                                                    _capi.dptr(tree), _capi.dptr(next_counts), _capi.dptr(best_gain), st()))
                d.record()
                torch.cuda.synchronize()
                tz += a.elapsed_time(b); th += b.elapsed_time(c); tp += c.elapsed_time(d)
            ev[2].record()
            _capi.check(lib.rdf_train_next_active(_capi.dptr(tree), level, D, C, _capi.dptr(active), S, _capi.dptr(next_active), _capi.dptr(num_next), st()))
            ev[3].record()
            _capi.check(lib.rdf_train_advance_pixels(_capi.dptr(depth), _capi.dptr(nodes_work), N, W, H, _capi.dptr(tree), level, D, C, st()))
            ev[4].record()
            torch.cuda.synchronize()
            t = {'bucket': ev[0].elapsed_time(ev[1]), 'zero': tz, 'hist': th, 'pick_best': tp,
                 'next_active': ev[2].elapsed_time(ev[3]), 'advance': ev[3].elapsed_time(ev[4])}
        print(json.dumps({'level': level, 'active_nodes': S, 'features': F, 'ms': {k: round(v, 3) for k, v in t.items()},
                          'next_active_nodes': int(num_next.item()), 'hist_GB': S * F * (NT + 1) * C * 4 / 1e9,
                          'pick_best_GBps': S * F * (NT + 1) * C * 4 / 1e6 / max(t['pick_best'], 1e-9),
                          'pick_screen': not os.environ.get('RDF_PICK_NO_SCREEN'),
                          # identical split records whatever the kernel variant: digest of the level's node records + child counts
                          'records_md5': hashlib.md5(tree.cpu().numpy().tobytes() + next_counts.cpu().numpy().tobytes()).hexdigest()}), flush=True)
        del hist, ws, nodes


if __name__ == '__main__':
    main()
