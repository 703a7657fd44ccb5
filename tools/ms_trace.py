#!/usr/bin/env python
"""Phase timeline of rdf_mean_shift_v2_kernel (rank 0 of the cluster) from %globaltimer stamps: RDF_MS_TRACE=1 makes the
kernel write them into the head of its workspace.  Run on a GPU box: RDF_MS_TRACE=1 python tools/ms_trace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault('RDF_MS_TRACE', '1')

import numpy as np  # noqa: E402

from rdf_b200 import synth  # noqa: E402
from rdf_b200.buffers import GPUArray  # noqa: E402
from rdf_b200.mean_shift import MeanShift  # noqa: E402
from oracle import numpy_oracle as no  # noqa: E402

H, W, r = 480, 848, 2
_, _, var = synth.layered_cfg2()
lab = np.full((1, H // r, W // r), 65535, np.uint16)          # cfg2-like label image: ellipse blob, 11 classes in stripes
m = synth.ellipse_mask(H, W)[::r, ::r]
xs = np.arange(W // r)[None, :].repeat(H // r, 0)
lab[0][m] = (1 + (xs[m] // 14) % 11).astype(np.uint16)
L = GPUArray(lab.shape, dtype=np.uint16)
L.set(lab)
ms = MeanShift()
for _ in range(5):
    out = ms.run(6, L, 11, var)
t = ms._workspace.get()[:64].view(np.uint64).astype(np.int64)
print('labelled px', int(m.sum()))
print('setup (ns): load+count', int(t[1] - t[0]), 'scan', int(t[2] - t[1]), 'scatter', int(t[3] - t[2]))
print('rounds (ns):', [int(t[5 + i] - t[4 + i]) for i in range(5)], 'last:', int(t[14] - t[9]), 'tail:', int(t[15] - t[14]), 'total:', int(t[15] - t[0]))
if os.environ.get('RDF_MS_V2'):
    print('round 1 detail (ns): items..lastwarp', int(t[22] - t[5]), 'class sums + DSMEM', int(t[23] - t[22]), 'cluster.sync', int(t[24] - t[23]),
          'means', int(t[25] - t[24]), 'to next round', int(t[6] - t[25]))
else:   # v3 (class-parallel): stamps of class 0, rank 0
    print('round 1 detail (ns): entries+exp', int(t[20] - t[5]), 'warp sums', int(t[21] - t[20]), 'CTA sum (+DSMEM)', int(t[22] - t[21]),
          'cluster.sync', int(t[23] - t[22]), 'mean + block sync', int(t[24] - t[23]), 'to next round', int(t[6] - t[24]))
print('max |ours - oracle|', float(np.nanmax(np.abs(out - no.mean_shift(lab, 11, var, 6)))))
