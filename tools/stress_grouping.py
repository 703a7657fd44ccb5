"""Race hunt for rdf_group_hands: tall narrow images (long union-find chains), every image 5 times, against the NumPy oracle.
   python tools/stress_grouping.py [iterations]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT+'/3d-beats_b200', ROOT+'/tests'): sys.path.insert(0,p)
import numpy as np, torch
from oracle import grouping_oracle as go
from test_grouping_oracle import blob_image
from rdf_b200.grouping import CppGrouping
from rdf_b200.buffers import GPUArray
g = CppGrouping()
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 400):
    h, w = [(200,33),(150,40),(60,106),(120,128)][it % 4]
    img = blob_image(h, w, 7000 + it, density=0.02 if it % 2 else 0.0)
    dev = GPUArray((h,w), dtype=np.uint16); dev.set(img)
    st = GPUArray((h,w), dtype=np.uint16); gi = GPUArray((2,3), dtype=np.float32)
    for rep in range(5):
        g.make_groups_cu(dev, st, gi, 0.005)
        a, b = st.get(), gi.get()
        so, go_ = go.make_groups(img, 0.005) if rep == 0 else (so, go_)
        if not (np.array_equal(a, so) and np.array_equal(b, go_)):
            bad += 1
            print('MISMATCH', it, rep, h, w, 'stencil', np.array_equal(a, so), 'ginfo', b.tolist(), go_.tolist(), flush=True)
print('done bad =', bad)
