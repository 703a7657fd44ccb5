"""The frame loop around the forest: mirror of the parts of the reference's `PointsOps` (src/cuda/points_ops.py) and
`CalibratedPlane.filter_points_by_plane` (src/calibrated_plane.py:27) that `App_3d_bz.tick` / `run_per_hand_pipeline` run before
and after `LayeredDecisionForest.run` (src/3d_bz.py:159-220, 252-259, 390-456, 503-522), over the fused entry points of
csrc/rdf_frame.cu.  The reference fetches one pycuda kernel per step (8 launches + 2 frame copies per frame, 5 + 2 per hand);
here a frame is conditioned in one launch, all hands are stencilled in one launch and the fingertip depths are read out in one.
"""
import ctypes

import numpy as np
import torch

from . import _capi
from .buffers import GPUArray, GpuBuffer, as_gpuarray


def gaussian_kernel(k_size, sigma):
    """src/cuda/points_ops.py:9-14; scipy.stats.norm.pdf(x, 0, sigma) written out (bit-identical, checked in tests)."""
    assert k_size % 2 == 1, 'kernel must be odd'
    l = k_size // 2
    x = (np.linspace(-l, l, k_size) - 0.) / sigma
    kern1d = np.exp(-x ** 2 / 2.0) / np.sqrt(2 * np.pi) / sigma
    kern2d = np.outer(kern1d, kern1d)
    return (kern2d / kern2d.sum()).astype(np.float32)


def _out_ptr(t):
    """device array or pinned host torch tensor (unified addressing makes pinned memory device-visible)"""
    if isinstance(t, torch.Tensor) and not t.is_cuda:
        assert t.is_pinned() and t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())
    return _capi.dptr(as_gpuarray(t))


class PointsOps():
    """Two call forms per kernel.  (1) This module's own: device arrays first, shapes taken from them.  (2) The reference's
    pycuda form, with the argument order of the `extern "C"` kernel and `grid=` / `block=` accepted and ignored (the launch
    geometry is the library's business), so that run_live.py:86-121, run_live_layered.py:87-135 and src/3d_bz.py:159-456 run
    unchanged against the drop-in (compat/rdf_dropin.py)."""
    MAX_FILTER_SIZE = 41          # src/cuda/points_ops.py:34

    def __init__(self):
        self._lib = _capi.load()
        self._gaussian_filter = None
        self._cached_filter_params = None

    def _filter(self, sigma, k_size):
        # cached on (sigma, k_size) like the reference (src/cuda/points_ops.py:85-88)
        assert k_size <= self.MAX_FILTER_SIZE
        if self._cached_filter_params != (sigma, k_size):
            if self._gaussian_filter is None:
                self._gaussian_filter = GPUArray((self.MAX_FILTER_SIZE * self.MAX_FILTER_SIZE,), dtype=np.float32)
            k = np.zeros(self.MAX_FILTER_SIZE * self.MAX_FILTER_SIZE, dtype=np.float32)
            k[:k_size * k_size] = gaussian_kernel(k_size, sigma).flatten()
            self._gaussian_filter.set(k)
            self._cached_filter_params = (sigma, k_size)
        return self._gaussian_filter

    def condition_depth(self, depth_in, depth_out, depth_mm, pp, focal, plane, plane_z_threshold, gauss_sigma=2.0, k_size=5,
                        mm_level=3):
        """src/3d_bz.py:159-220 in one launch.  depth_in / depth_out uint16[H,W] (device, distinct), depth_mm uint16[H>>l,W>>l]
        or None, plane = device float32[4,4] (CalibratedPlane.get_mat() uploaded once).  gauss_sigma <= 0.1 skips the filter as the
        reference does (src/3d_bz.py:205)."""
        depth_out, plane = as_gpuarray(depth_out), as_gpuarray(plane)
        pinned_in = isinstance(depth_in, torch.Tensor) and not depth_in.is_cuda     # zero-copy: the kernel reads the host frame itself
        if not pinned_in:
            depth_in = as_gpuarray(depth_in)
            assert depth_in.dtype == np.uint16
        assert depth_out.dtype == np.uint16 and plane.dtype == np.float32 and plane.size == 16
        H, W = depth_in.shape[-2:]
        assert depth_out.size == H * W
        mm = None
        if depth_mm is not None:
            mm = as_gpuarray(depth_mm)
            assert mm.dtype == np.uint16 and mm.size == (H >> mm_level) * (W >> mm_level)
        filt = self._filter(gauss_sigma, k_size) if gauss_sigma > 0.1 else None
        _capi.check(self._lib.rdf_condition_depth(_out_ptr(depth_in), W, H, float(pp[0]), float(pp[1]), float(focal), _capi.dptr(plane),
                                                  float(plane_z_threshold), _capi.dptr(filt), int(k_size), int(mm_level),
                                                  _capi.dptr(depth_out), _capi.dptr(mm), _capi.stream_ptr()))

    def gaussian_depth_filter(self, d_in, d_out, sigma, k_size=5):
        """src/cuda/points_ops.py:62-98 alone: the fused kernel with an identity plane and an infinite clip distance."""
        if not hasattr(self, '_identity'):
            self._identity = GPUArray((4, 4), dtype=np.float32)
            self._identity.set(np.eye(4, dtype=np.float32))
        d_in, d_out = as_gpuarray(d_in), as_gpuarray(d_out)
        assert d_in.shape == d_out.shape and d_in.dtype == np.uint16 and d_out.dtype == np.uint16
        H, W = d_in.shape[-2:]
        _capi.check(self._lib.rdf_condition_depth(_capi.dptr(d_in), W, H, 0., 0., 1., _capi.dptr(self._identity), float('-inf'),
                                                  _capi.dptr(self._filter(sigma, k_size)), int(k_size), 0, _capi.dptr(d_out), None,
                                                  _capi.stream_ptr()))

    def grow_groups(self, *args, grid=None, block=None):
        """grow_groups(g_in, g_out)  |  pycuda form (IMG_DIM int32[2] = (x, y), g_in, g_out, grid=, block=), src/3d_bz.py:254-259"""
        g_in, g_out = args[-2:]
        g_in, g_out = as_gpuarray(g_in), as_gpuarray(g_out)
        if len(args) == 3:
            g_in, g_out = g_in.reshape((int(args[0][1]), int(args[0][0]))), g_out.reshape((int(args[0][1]), int(args[0][0])))
        assert g_in.dtype == np.uint16 and g_out.dtype == np.uint16 and g_in.size == g_out.size
        h, w = g_in.shape[-2:]
        _capi.check(self._lib.rdf_grow_groups(_capi.dptr(g_in), w, h, _capi.dptr(g_out), _capi.stream_ptr()))

    def stencil_hands(self, depth, groups, mm_level, hands, out, grow=True):
        """src/3d_bz.py:390-420 for every (group id, flip_x) in `hands`, one launch; out uint16[len(hands),H,W].
        grow=True applies grow_groups (src/3d_bz.py:252-259) to `groups` on the fly."""
        depth, groups, out = as_gpuarray(depth), as_gpuarray(groups), as_gpuarray(out)
        assert depth.dtype == np.uint16 and groups.dtype == np.uint16 and out.dtype == np.uint16
        H, W = depth.shape[-2:]
        n = len(hands)
        assert out.size == n * H * W and groups.size == (H >> mm_level) * (W >> mm_level)
        ids = (ctypes.c_int * n)(*[int(g) for g, _ in hands])
        flips = (ctypes.c_int * n)(*[int(bool(f)) for _, f in hands])
        _capi.check(self._lib.rdf_stencil_hands(_capi.dptr(depth), W, H, _capi.dptr(groups), int(mm_level), int(bool(grow)), n, ids, flips,
                                                _capi.dptr(out), _capi.stream_ptr()))

    def flip_x(self, *args, grid=None, block=None):
        """flip_x(img_in, img_out)  |  pycuda form (IMG_DIM int32[2] = (x, y), in, out, grid=, block=), src/3d_bz.py:407-411,441-446"""
        img_in, img_out = args[-2:]
        img_in, img_out = as_gpuarray(img_in), as_gpuarray(img_out)
        if len(args) == 3:
            img_in, img_out = img_in.reshape((int(args[0][1]), int(args[0][0]))), img_out.reshape((int(args[0][1]), int(args[0][0])))
        assert img_in.dtype == np.uint16 and img_out.dtype == np.uint16 and img_in.size == img_out.size
        h, w = img_in.shape[-2:]
        _capi.check(self._lib.rdf_flip_x(_capi.dptr(img_in), w, h, _capi.dptr(img_out), _capi.stream_ptr()))

    def make_rgba_from_labels(self, *args, grid=None, block=None):
        """make_rgba_from_labels(labels, colors, rgba)  |  pycuda form (IMG_DIM_X, IMG_DIM_Y, NUM_COLORS, labels, colors, rgba,
        grid=, block=), src/run_live_layered.py:126-135, src/3d_bz.py:448-456"""
        labels, colors, rgba = (as_gpuarray(a) for a in args[-3:])
        assert labels.dtype == np.uint16 and colors.dtype == np.uint8 and rgba.dtype == np.uint8
        if len(args) == 6:
            w, h, num_colors = int(args[0]), int(args[1]), int(args[2])
            assert labels.size == w * h and colors.size >= 4 * num_colors
        else:
            (h, w), num_colors = labels.shape[-2:], colors.size // 4
        assert rgba.size == h * w * 4
        _capi.check(self._lib.rdf_labels_to_rgba(_capi.dptr(labels), w, h, _capi.dptr(colors), num_colors, _capi.dptr(rgba),
                                                 _capi.stream_ptr()))

    def make_depth_rgba(self, *args, grid=None, block=None):
        """make_depth_rgba(depth, d_min, d_max, rgba)  |  pycuda form (IMG_DIM int32[2] = (x, y), d_min, d_max, depth, rgba, grid=,
        block=), src/3d_bz.py:266-274"""
        if len(args) == 5:
            dims, d_min, d_max, depth, rgba = args
            depth = as_gpuarray(depth).reshape((int(dims[1]), int(dims[0])))
        else:
            depth, d_min, d_max, rgba = args
        depth, rgba = as_gpuarray(depth), as_gpuarray(rgba)
        assert depth.dtype == np.uint16 and rgba.dtype == np.uint8
        h, w = depth.shape[-2:]
        assert rgba.size == h * w * 4
        _capi.check(self._lib.rdf_depth_to_rgba(_capi.dptr(depth), w, h, int(d_min), int(d_max), _capi.dptr(rgba), _capi.stream_ptr()))

    def fingertip_z(self, means, fingertip_idxes, labels_reduce, raw_depth, pp, fx, fy, plane, z_out, means_copy=None):
        """src/3d_bz.py:503-522 in one launch: means float64[K,2] or [N,K,2] (N hands), z_out float64[(N,) len(fingertip_idxes)]
        (device array or pinned host tensor), NaN where the reference resets the fingertip; means_copy (optional, same kinds)
        receives a copy of `means`."""
        means, plane = as_gpuarray(means), as_gpuarray(plane)
        if not (isinstance(raw_depth, torch.Tensor) and not raw_depth.is_cuda):    # else: pinned host frame, read zero-copy
            raw_depth = as_gpuarray(raw_depth)
            assert raw_depth.dtype == np.uint16
        assert means.dtype == np.float64 and plane.dtype == np.float32
        H, W = raw_depth.shape[-2:]
        n = len(fingertip_idxes)
        idx = (ctypes.c_int * n)(*[int(i) for i in fingertip_idxes])
        num_images = means.shape[0] if len(means.shape) == 3 else 1
        _capi.check(self._lib.rdf_fingertip_z(_capi.dptr(means), num_images, means.shape[-2], idx, n, int(labels_reduce), _out_ptr(raw_depth), W, H,
                                              float(pp[0]), float(pp[1]), float(fx), float(fy), _capi.dptr(plane), _out_ptr(z_out),
                                              None if means_copy is None else _out_ptr(means_copy), _capi.stream_ptr()))

    # ---- the reference's remaining kernels, pycuda call form only (csrc/rdf_points.cu) --------------------------------------
    @staticmethod
    def _pts(pts, n):
        pts = as_gpuarray(pts)
        assert pts.dtype == np.float32 and pts.size >= 4 * n, 'points buffer: float32[n, 4]'
        return pts

    def deproject_points(self, imgs_dim, pp, f, imgs, pts, grid=None, block=None):
        """src/cuda/points_ops.cu:5-36; imgs_dim = int32[4] (num_images, dim_x, dim_y, -)  (src/run_live_layered.py:87-94)"""
        n, dim_x, dim_y = int(imgs_dim[0]), int(imgs_dim[1]), int(imgs_dim[2])
        imgs = as_gpuarray(imgs)
        assert imgs.dtype == np.uint16 and imgs.size == n * dim_x * dim_y
        _capi.check(self._lib.rdf_deproject_points(_capi.dptr(imgs), n, dim_x, dim_y, float(pp[0]), float(pp[1]), float(f),
                                                   _capi.dptr(self._pts(pts, n * dim_x * dim_y)), _capi.stream_ptr()))

    def transform_points(self, num_pts, pts, t, grid=None, block=None):
        """src/cuda/points_ops.cu:66-75; t = the row-major numpy float32[4,4] CalibratedPlane.get_mat() returns"""
        t = np.ascontiguousarray(t, dtype=np.float32)
        assert t.size == 16
        _capi.check(self._lib.rdf_transform_points(int(num_pts), _capi.dptr(self._pts(pts, int(num_pts))),
                                                   t.ctypes.data_as(ctypes.c_void_p), _capi.stream_ptr()))

    def filter_points_by_plane(self, num_pts, plane_z_threshold, pts, grid=None, block=None):
        """src/cuda/calibrated_plane.cu:30-45 (the reference fetches it on CalibratedPlane; see rdf_b200/calibrated_plane.py)"""
        _capi.check(self._lib.rdf_filter_points_by_plane(int(num_pts), float(plane_z_threshold), _capi.dptr(self._pts(pts, int(num_pts))),
                                                         _capi.stream_ptr()))

    def _depth_n(self, depth, n):
        depth = as_gpuarray(depth)
        assert depth.dtype == np.uint16 and depth.size >= n
        return depth

    def remove_missing_3d_points_from_depth_image(self, num_pixels, pts, depth, grid=None, block=None):
        """src/cuda/points_ops.cu:131-146"""
        n = int(num_pixels)
        _capi.check(self._lib.rdf_remove_missing_points(n, _capi.dptr(self._pts(pts, n)), _capi.dptr(self._depth_n(depth, n)), _capi.stream_ptr()))

    def setup_depth_image_for_forest(self, num_pixels, pts, depth, grid=None, block=None):
        """src/cuda/points_ops.cu:149-165  (src/run_live.py:116-121, src/run_live_layered.py:117-122)"""
        n = int(num_pixels)
        _capi.check(self._lib.rdf_setup_depth_for_forest(n, _capi.dptr(self._pts(pts, n)), _capi.dptr(self._depth_n(depth, n)), _capi.stream_ptr()))

    def convert_0s_to_maxuint(self, num_pixels, depth, grid=None, block=None):
        """src/cuda/points_ops.cu:118-127"""
        n = int(num_pixels)
        _capi.check(self._lib.rdf_zeros_to_no_pixel(n, _capi.dptr(self._depth_n(depth, n)), _capi.stream_ptr()))

    def shrink_image(self, img_dim_in, mipmap_level, d_in, d_out, grid=None, block=None):
        """src/cuda/points_ops.cu:375-404; img_dim_in = int32[2] (x, y)"""
        dim_x, dim_y, level = int(img_dim_in[0]), int(img_dim_in[1]), int(mipmap_level)
        d_in, d_out = self._depth_n(d_in, dim_x * dim_y), self._depth_n(d_out, (dim_x >> level) * (dim_y >> level))
        _capi.check(self._lib.rdf_shrink_image(_capi.dptr(d_in), dim_x, dim_y, level, _capi.dptr(d_out), _capi.stream_ptr()))

    def stencil_depth_image_by_group(self, img_dim, mipmap_level, group, g_in, d_in, d_out, grid=None, block=None):
        """src/cuda/points_ops.cu:441-463; img_dim = int32[2] (x, y)"""
        dim_x, dim_y, level = int(img_dim[0]), int(img_dim[1]), int(mipmap_level)
        g_in = self._depth_n(g_in, (dim_x >> level) * (dim_y >> level))
        d_in, d_out = self._depth_n(d_in, dim_x * dim_y), self._depth_n(d_out, dim_x * dim_y)
        _capi.check(self._lib.rdf_stencil_by_group(_capi.dptr(g_in), _capi.dptr(d_in), dim_x, dim_y, level, int(group), _capi.dptr(d_out),
                                                   _capi.stream_ptr()))

    def write_pixel_groups_to_stencil_image(self, coords, num_coords, stencil, stencil_dims, grid=None, block=None):
        """src/cuda/points_ops.cu:486-503; coords int32[>= num_coords, 3] = (row, col, group), stencil_dims = (rows, cols)"""
        coords, n = as_gpuarray(coords), int(num_coords)
        rows, cols = int(stencil_dims[0]), int(stencil_dims[1])
        assert coords.dtype == np.int32 and coords.size >= 3 * n
        _capi.check(self._lib.rdf_scatter_groups(_capi.dptr(coords), n, _capi.dptr(self._depth_n(stencil, rows * cols)), rows, cols,
                                                 _capi.stream_ptr()))
