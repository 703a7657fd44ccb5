"""Per level of a cfg3-shaped walk (hash forest tree 0, one 848x480 frame): how many distinct nodes the 32 lanes of a 16x2 warp patch
sit on (CPU, NumPy).  usage: python tools/uniform_stats.py [dense-smooth|dense-noise]  -> profiles/r02_micro_l1_groups.md"""
import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, '3d-beats_b200')); sys.path.insert(0, ROOT)
from rdf_b200 import synth
H,W,D=480,848,20
kind=sys.argv[1] if len(sys.argv)>1 else 'dense-smooth'
depth=synth.depth_frames(kind,1,H,W)[0].astype(np.int64)
forest=synth.hash_forest(1,D,4,trees=[0])[0]   # [2^D-1,15]
yy,xx=np.mgrid[0:H,0:W]
node=np.zeros((H,W),np.int64)
df=depth.astype(np.float32)
def probe(ox,oy):
    px=xx+ox; py=yy+oy
    ok=(px>=0)&(px<W)&(py>=0)&(py<H)
    v=np.full((H,W),65535,np.int64)
    v[ok]=depth[py[ok],px[ok]]
    return v
print(kind)
for lev in range(D):
    row=(1<<lev)-1+node
    nd=forest[row]
    ox=np.floor((nd[...,0]/df).astype(np.float32)).astype(np.int64); oy=np.floor((nd[...,1]/df).astype(np.float32)).astype(np.int64)
    vx=np.floor((nd[...,2]/df).astype(np.float32)).astype(np.int64); vy=np.floor((nd[...,3]/df).astype(np.float32)).astype(np.int64)
    f=(probe(ox,oy)-probe(vx,vy)).astype(np.float32)
    # stats on node distinctness per 16x2 warp patch BEFORE stepping (this level's header load)
    n=node.reshape(H//2,2,W//16,16).transpose(0,2,1,3).reshape(-1,32)
    ns=np.sort(n,axis=1); distinct=1+(np.diff(ns,axis=1)!=0).sum(1)
    print(f'level {lev:2d}: uniform warps {np.mean(distinct==1)*100:5.1f}%  mean distinct {distinct.mean():5.2f}  <=2: {np.mean(distinct<=2)*100:5.1f}% <=4: {np.mean(distinct<=4)*100:5.1f}%')
    side=(~(f<nd[...,4])).astype(np.int64)
    node=2*node+side
