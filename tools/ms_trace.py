import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/3d-beats_b200')
import numpy as np, torch
from rdf_b200 import synth
from rdf_b200.mean_shift import MeanShift
from rdf_b200.buffers import GPUArray
from oracle import numpy_oracle as no
H,W,r=480,848,2
depth=synth.depth_frames('live-mask',1,H,W)
forests,cfg,var=synth.layered_cfg2()
comp,_=no.layered_run([f[:, :1023] for f in forests] if False else forests,[(None,None),(0,1)],cfg['conditions'],depth[0],r,1.0) if False else (None,None)
# cheap label image: ellipse blob with 11 classes in vertical stripes
lab=np.full((1,H//r,W//r),65535,np.uint16)
m=synth.ellipse_mask(H,W)[::r,::r]
xs=np.arange(W//r)[None,:].repeat(H//r,0)
lab[0][m]=(1+(xs[m]//14)%11).astype(np.uint16)
L=GPUArray(lab.shape,dtype=np.uint16); L.set(lab)
ms=MeanShift()
import time
for i in range(int(os.environ.get('ITERS','5'))):
    out=ms.run(6,L,11,var)
tr=ms._workspace.get()[:64].view(np.uint64)
t=tr.astype(np.int64)
print('phases(ns):', [int(t[i+1]-t[i]) for i in range(0,3)], 'rounds:', [int(t[5+i]-t[4+i]) for i in range(5)], 'last round+:', int(t[14]-t[9]), 'tail:', int(t[15]-t[14]), 'total:', int(t[15]-t[0]))
print('round1 detail (ns): start->items', int(t[20]-t[5]), 'warpsum', int(t[21]-t[20]), 'sync', int(t[22]-t[21]), 'classsum+dsmem', int(t[23]-t[22]), 'cluster.sync', int(t[24]-t[23]), 'means', int(t[25]-t[24]), 'to next', int(t[6]-t[25]))
print('cycles', int(t[17]-t[16]), 'MHz', (t[17]-t[16])/(t[15]-t[0])*1e3)
print(np.nanmax(np.abs(out-no.mean_shift(lab,11,var,6))))
