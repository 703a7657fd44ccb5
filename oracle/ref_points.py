"""Launch the REFERENCE's own pre-/post-forest kernels (src/cuda/points_ops.cu, calibrated_plane.cu compiled unchanged for sm_100a
into oracle/_ref/libref_points.so) in the order src/3d_bz.py runs them.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, '_ref', 'libref_points.so')
_lib = None


def available():
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f'{SO_PATH} missing: run `make -C oracle ref` where /root/reference exists')
        _lib = ctypes.CDLL(SO_PATH)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f(v):
    return ctypes.c_float(float(v))


def _ok(rc):
    if rc != 0:
        raise RuntimeError(f'reference kernel launch failed ({rc})')


def _dev_u16(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint16).view(np.int16)).cuda()


def _host_u16(t):
    return t.cpu().numpy().view(np.uint16)


def condition_frame(depth_raw, pp, focal, plane, plane_z_threshold, gauss_kernel=None, mm_level=3):
    """src/3d_bz.py:159-220 with the reference's kernels; returns (depth, depth_mm) as numpy uint16."""
    H, W = depth_raw.shape
    depth = _dev_u16(depth_raw)
    pts = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    plane_h = np.ascontiguousarray(plane, dtype=np.float32)
    L = lib()
    _ok(L.ref_deproject_points(W, H, _f(pp[0]), _f(pp[1]), _f(focal), _p(depth), _p(pts), _st()))
    _ok(L.ref_transform_points(W * H, _p(pts), plane_h.ctypes.data_as(ctypes.c_void_p), _st()))
    _ok(L.ref_filter_points_by_plane(W * H, _f(plane_z_threshold), _p(pts), _st()))
    _ok(L.ref_remove_missing(W * H, _p(pts), _p(depth), _st()))
    if gauss_kernel is not None:
        k = torch.from_numpy(np.ascontiguousarray(gauss_kernel, dtype=np.float32)).cuda()
        depth_2 = depth.clone()
        _ok(L.ref_gaussian_depth_filter(W, H, int(gauss_kernel.shape[0]), _p(k), _p(depth_2), _p(depth), _st()))
    mm = torch.zeros((H >> mm_level, W >> mm_level), dtype=torch.int16, device='cuda')
    if mm.numel():                                         # an empty grid is a launch error
        _ok(L.ref_shrink_image(W, H, mm_level, _p(depth), _p(mm), _st()))
    torch.cuda.synchronize()
    return _host_u16(depth), _host_u16(mm)


def live_frame_for_forest(depth_raw, pp, focal, plane, plane_z_threshold):
    """src/run_live.py:86-121 / src/run_live_layered.py:87-122 with the reference's kernels: deproject -> transform -> plane filter ->
    setup_depth_image_for_forest.  Returns (pts float32[H,W,4], depth uint16[H,W]) as numpy."""
    H, W = depth_raw.shape
    depth = _dev_u16(depth_raw)
    pts = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    plane_h = np.ascontiguousarray(plane, dtype=np.float32)
    L = lib()
    _ok(L.ref_deproject_points(W, H, _f(pp[0]), _f(pp[1]), _f(focal), _p(depth), _p(pts), _st()))
    _ok(L.ref_transform_points(W * H, _p(pts), plane_h.ctypes.data_as(ctypes.c_void_p), _st()))
    _ok(L.ref_filter_points_by_plane(W * H, _f(plane_z_threshold), _p(pts), _st()))
    _ok(L.ref_setup_depth_image_for_forest(W * H, _p(pts), _p(depth), _st()))
    torch.cuda.synchronize()
    return pts.cpu().numpy(), _host_u16(depth)


def grow_groups(g_in):
    h, w = g_in.shape
    a = _dev_u16(g_in)
    out = torch.zeros_like(a)
    _ok(lib().ref_grow_groups(w, h, _p(a), _p(out), _st()))
    torch.cuda.synchronize()
    return _host_u16(out)


def hand_depth_image(depth, groups_grown, mm_level, g_id, flip_x):
    """src/3d_bz.py:390-420 with the reference's kernels."""
    H, W = depth.shape
    L = lib()
    d = _dev_u16(depth)
    g = _dev_u16(groups_grown)
    group = torch.zeros_like(d)
    _ok(L.ref_stencil_depth_image_by_group(W, H, mm_level, int(g_id), _p(g), _p(d), _p(group), _st()))
    if flip_x:
        out = torch.zeros_like(d)
        _ok(L.ref_flip_x(W, H, _p(group), _p(out), _st()))
    else:
        out = group.clone()
    _ok(L.ref_convert_0s_to_maxuint(W * H, _p(out), _st()))
    torch.cuda.synchronize()
    return _host_u16(out)


def flip_x(img):
    h, w = img.shape
    a = _dev_u16(img)
    out = torch.zeros_like(a)
    _ok(lib().ref_flip_x(w, h, _p(a), _p(out), _st()))
    torch.cuda.synchronize()
    return _host_u16(out)


def make_rgba_from_labels(labels, colors, rgba_init):
    h, w = labels.shape
    a = _dev_u16(labels)
    c = torch.from_numpy(np.ascontiguousarray(colors, dtype=np.uint8)).cuda()
    out = torch.from_numpy(np.ascontiguousarray(rgba_init, dtype=np.uint8)).cuda()
    _ok(lib().ref_make_rgba_from_labels(w, h, colors.shape[0], _p(a), _p(c), _p(out), _st()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def make_depth_rgba(depth, d_min, d_max):
    h, w = depth.shape
    a = _dev_u16(depth)
    out = torch.zeros((h, w, 4), dtype=torch.uint8, device='cuda')
    _ok(lib().ref_make_depth_rgba(w, h, int(d_min), int(d_max), _p(a), _p(out), _st()))
    torch.cuda.synchronize()
    return out.cpu().numpy()
