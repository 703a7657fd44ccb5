#!/usr/bin/env python
"""Hand grouping per frame (106 x 60 = the product's 848x480 depth image shrunk by 8): device kernel vs the reference's C++
flood fill on the host (oracle/_ref/libref_grouping.so, its D2H / H2D copies not included)."""
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

from rdf_b200.grouping import CppGrouping  # noqa: E402
from oracle import grouping_oracle as go  # noqa: E402
from test_grouping_oracle import blob_image  # noqa: E402

img = blob_image(60, 106, 3, density=0.01)
g = CppGrouping()
d = torch.from_numpy(img.view(np.int16)).cuda().view(torch.uint16)
st = torch.zeros_like(d)
gi = torch.zeros((2, 3), device='cuda')
for _ in range(10):
    g.make_groups_cu(d, st, gi, 0.005)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    g.make_groups_cu(d, st, gi, 0.005)
    s.synchronize()
    with torch.cuda.graph(graph, stream=s):
        g.make_groups_cu(d, st, gi, 0.005)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300):
        graph.replay()
    e1.record()
s.synchronize()
t_gpu = e0.elapsed_time(e1) / 300 * 1e3
out = {'device_us_per_frame': round(t_gpu, 2), 'pixels': int(img.size), 'foreground': int((img != 0).sum())}
if go.ref_available():
    t0 = time.perf_counter()
    for _ in range(300):
        coords, g_ref = go.ref_make_groups(img, 0.005)
    out['reference_cpp_host_us_per_frame'] = round((time.perf_counter() - t0) / 300 * 1e6, 2)
    out['stencil_identical'] = bool(np.array_equal(st.cpu().view(torch.int16).numpy().view(np.uint16), go.stencil_from_coords(coords, 60, 106)))
print(json.dumps(out))
