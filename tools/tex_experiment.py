#!/usr/bin/env python
"""Texture-probe forest eval vs the global-load path: parity (vs C oracle on small cases, vs the global path at cfg3 size) + speed."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rdf_b200 import _capi, synth  # noqa: E402
from rdf_b200 import decision_tree as dt  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

lib = _capi.load()
st = _capi.stream_ptr


def tex_eval(forest, depth, labels, r=1, filt=None, fclass=-1):
    N, H, W = depth.shape
    h = ctypes.c_void_p()
    _capi.check(lib.rdf_depth_tex_create(N, W, H, ctypes.byref(h)))
    _capi.check(lib.rdf_depth_tex_upload(h, _capi.dptr(depth), N, st()))
    _capi.check(lib.rdf_eval_forest_tex(forest.handle(), h, _capi.dptr(depth), N, _capi.dptr(filt), fclass, _capi.dptr(labels), r, st()))
    torch.cuda.synchronize()
    return h


# ---- parity on small cases ----
ok = True
for seed, (kind, T, D, C, r) in enumerate([('dense-smooth', 3, 8, 4, 1), ('dense-noise', 4, 7, 11, 1), ('live-mask', 2, 9, 3, 2), ('dense-noise', 8, 5, 4, 3)]):
    depth_np = synth.depth_frames(kind, 3, 61, 83, seed=seed)
    depth_np[0, 3:6, 4:9] = 0
    forest_np = synth.random_forest(T, D, C, seed=seed, ragged=True)
    f = dt.DecisionForest(T, D, C)
    f.forest_cu.set(forest_np)
    d = dt.cu_array.to_gpu(depth_np)
    lab = dt.cu_array.GPUArray((3, 61 // r, 83 // r), dtype=np.uint16).fill(65535)
    hh = tex_eval(f, d, lab, r)
    lib.rdf_depth_tex_destroy(hh)
    exp = np.full((3, 61 // r, 83 // r), 65535, np.uint16)
    co.eval_forest(forest_np, depth_np, exp, r)
    same = bool(np.array_equal(lab.get(), exp))
    ok = ok and same
    print('parity', kind, T, D, C, r, same)

# ---- speed at cfg3 shape ----
out = {'small_case_parity': ok}
for kind in ('dense-smooth', 'dense-noise'):
    N, H, W, T, D, C = 512, 480, 848, 4, 20, 4
    forest = dt.DecisionForest(T, D, C)
    _capi.check(lib.rdf_synth_forest(_capi.dptr(forest.forest_cu), T, D, C, 1234, st()))
    depth = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16)
    _capi.check(lib.rdf_synth_depth(_capi.dptr(depth), {'dense-smooth': 0, 'dense-noise': 1}[kind], N, W, H, 1234, 0, st()))
    a = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16).fill(65535)
    b = dt.cu_array.GPUArray((N, H, W), dtype=np.uint16).fill(65535)
    ev = dt.DecisionTreeEvaluator()
    ev.get_labels_forest(forest, depth, a)
    h = tex_eval(forest, depth, b)
    same = bool(torch.equal(a.tensor.view(torch.int16), b.tensor.view(torch.int16)))
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    e0.record()
    for _ in range(3):
        ev.get_labels_forest(forest, depth, a)
    e1.record()
    for _ in range(3):
        _capi.check(lib.rdf_eval_forest_tex(forest.handle(), h, _capi.dptr(depth), N, None, -1, _capi.dptr(b), 1, st()))
    e2.record()
    for _ in range(3):
        _capi.check(lib.rdf_depth_tex_upload(h, _capi.dptr(depth), N, st()))
    e3.record()
    torch.cuda.synchronize()
    px = N * H * W
    out[kind] = {'identical_labels': same, 'global_path_mpx_s': px * 3 / e0.elapsed_time(e1) / 1e3, 'tex_path_mpx_s': px * 3 / e1.elapsed_time(e2) / 1e3,
                 'tex_upload_ms': e2.elapsed_time(e3) / 3}
    lib.rdf_depth_tex_destroy(h)
print(json.dumps(out))
