"""CPU oracle of the live frame loop around the forest (src/3d_bz.py:156-260,387-522): ctypes front end of oracle/frame_oracle.c
(one C function per reference kernel) plus the reference's HOST sequence restated call by call.  TEST INFRASTRUCTURE ONLY.

Parity pin: tests/golden/frame.npz holds outputs of the reference's own kernels (oracle/ref_points.py, compiled from
/root/reference against oracle/ref_kernels/glm_min) on a B200 for the seeded scenes of rdf_b200.synth.live_scene;
tests/test_frame_oracle.py checks this restatement against them on the CPU."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, '_build', 'libframe_oracle.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FileNotFoundError(f'{SO_PATH} missing: run `make -C oracle`')
        _lib = ctypes.CDLL(SO_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f(v):
    return ctypes.c_float(float(v))


def gaussian_kernel(k_size, sigma):
    """src/cuda/points_ops.py:9-14 (scipy.stats.norm.pdf written out: exp(-x^2/2) / sqrt(2 pi) / sigma)."""
    assert k_size % 2 == 1, 'kernel must be odd'
    l = k_size // 2
    x = (np.linspace(-l, l, k_size) - 0.) / sigma
    kern1d = np.exp(-x ** 2 / 2.0) / np.sqrt(2 * np.pi) / sigma
    kern2d = np.outer(kern1d, kern1d)
    return (kern2d / kern2d.sum()).astype(np.float32)


def deproject_points(depth, pp, focal, pts):
    H, W = depth.shape
    lib().fo_deproject_points(W, H, _f(pp[0]), _f(pp[1]), _f(focal), _p(depth), _p(pts))


def transform_points(pts, plane):
    plane = np.ascontiguousarray(plane, dtype=np.float32)
    lib().fo_transform_points(pts.size // 4, _p(pts), _p(plane))


def filter_points_by_plane(pts, thresh):
    lib().fo_filter_points_by_plane(pts.size // 4, _f(thresh), _p(pts))


def remove_missing_3d_points_from_depth_image(pts, depth):
    lib().fo_remove_missing(depth.size, _p(pts), _p(depth))


def gaussian_depth_filter(d_in, d_out, sigma, k_size=5):
    H, W = d_in.shape
    k = np.ascontiguousarray(gaussian_kernel(k_size, sigma))
    lib().fo_gaussian_depth_filter(W, H, k_size, _p(k), _p(d_in), _p(d_out))


def shrink_image(d_in, level):
    H, W = d_in.shape
    out = np.zeros((H >> level, W >> level), dtype=np.uint16)
    lib().fo_shrink_image(W, H, level, _p(d_in), _p(out))
    return out


def grow_groups(g_in):
    h, w = g_in.shape
    out = np.zeros_like(g_in)
    lib().fo_grow_groups(w, h, _p(np.ascontiguousarray(g_in)), _p(out))
    return out


def condition_frame(depth_raw, pp, focal, plane, plane_z_threshold, gauss_sigma=2.0, k_size=5, mm_level=3, pts=None):
    """src/3d_bz.py:159-220: returns (conditioned depth image, 1/2^level image).  `pts` float32[H,W,4] is the persistent point
    buffer of the reference (stale where depth is 0); zeros when not given."""
    depth = np.ascontiguousarray(depth_raw, dtype=np.uint16).copy()
    H, W = depth.shape
    if pts is None:
        pts = np.zeros((H, W, 4), dtype=np.float32)
    deproject_points(depth, pp, focal, pts)
    transform_points(pts, plane)
    filter_points_by_plane(pts, plane_z_threshold)
    remove_missing_3d_points_from_depth_image(pts, depth)
    if gauss_sigma > 0.1:
        depth_2 = depth.copy()
        gaussian_depth_filter(depth_2, depth, gauss_sigma, k_size)
    return depth, shrink_image(depth, mm_level)


def hand_depth_image(depth, groups_grown, mm_level, g_id, flip_x):
    """run_per_hand_pipeline's pre-processing, src/3d_bz.py:390-420."""
    H, W = depth.shape
    group = np.zeros_like(depth)
    lib().fo_stencil_depth_image_by_group(W, H, mm_level, int(g_id), _p(np.ascontiguousarray(groups_grown)), _p(depth), _p(group))
    if flip_x:
        out = np.zeros_like(depth)
        lib().fo_flip_x(W, H, _p(group), _p(out))
    else:
        out = group.copy()
    lib().fo_convert_0s_to_maxuint(out.size, _p(out))
    return out


def flip_x(img):
    h, w = img.shape
    out = np.zeros_like(img)
    lib().fo_flip_x(w, h, _p(np.ascontiguousarray(img)), _p(out))
    return out


def make_rgba_from_labels(labels, colors, rgba):
    h, w = labels.shape
    lib().fo_make_rgba_from_labels(w, h, colors.shape[0], _p(labels), _p(np.ascontiguousarray(colors)), _p(rgba))


def make_depth_rgba(depth, d_min, d_max):
    h, w = depth.shape
    rgba = np.zeros((h, w, 4), dtype=np.uint8)
    lib().fo_make_depth_rgba(w, h, ctypes.c_uint16(d_min), ctypes.c_uint16(d_max), _p(np.ascontiguousarray(depth)), _p(rgba))
    return rgba


def fingertip_z(label_means, fingertip_idxes, labels_reduce, depth_raw, pp, fx, fy, plane):
    """src/3d_bz.py:503-522.  Returns float64[len(fingertip_idxes)], NaN where the reference calls reset_positions().
    rs2_deproject_pixel_to_point without distortion: fp32 x = (px - ppx) / fx, y = (py - ppy) / fy, point = (z*x, z*y, z)."""
    H, W = depth_raw.shape
    plane = np.asarray(plane, dtype=np.float32)
    out = np.full(len(fingertip_idxes), np.nan)
    for i, f_idx in enumerate(fingertip_idxes):
        with np.errstate(invalid='ignore'):
            m = np.asarray(label_means[f_idx - 1], dtype=np.float64)
            pxy = np.where(np.isfinite(m) & (np.abs(m) < 2147483648.0), np.trunc(m), -2147483648.0).astype(np.int64)
        px, py = int(pxy[0]) * labels_reduce, int(pxy[1]) * labels_reduce   # int64 product, as NumPy < 1.24 promotes on Linux
        if px < 0 or py < 0 or px >= W or py >= H:
            continue
        z = np.float32(depth_raw[py, px])
        x = (np.float32(px) - np.float32(pp[0])) / np.float32(fx)
        y = (np.float32(py) - np.float32(pp[1])) / np.float32(fy)
        pt = [float(z * x), float(z * y), float(z), 1.]
        r = plane[2].astype(np.float64)
        out[i] = -(((r[0] * pt[0] + r[1] * pt[1]) + r[2] * pt[2]) + r[3] * pt[3])
    return out
