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

    def grow_groups(self, g_in, g_out):
        g_in, g_out = as_gpuarray(g_in), as_gpuarray(g_out)
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

    def flip_x(self, img_in, img_out):
        img_in, img_out = as_gpuarray(img_in), as_gpuarray(img_out)
        assert img_in.dtype == np.uint16 and img_out.dtype == np.uint16 and img_in.size == img_out.size
        h, w = img_in.shape[-2:]
        _capi.check(self._lib.rdf_flip_x(_capi.dptr(img_in), w, h, _capi.dptr(img_out), _capi.stream_ptr()))

    def make_rgba_from_labels(self, labels, colors, rgba):
        labels, colors, rgba = as_gpuarray(labels), as_gpuarray(colors), as_gpuarray(rgba)
        assert labels.dtype == np.uint16 and colors.dtype == np.uint8 and rgba.dtype == np.uint8
        h, w = labels.shape[-2:]
        assert rgba.size == h * w * 4
        _capi.check(self._lib.rdf_labels_to_rgba(_capi.dptr(labels), w, h, _capi.dptr(colors), colors.size // 4, _capi.dptr(rgba),
                                                 _capi.stream_ptr()))

    def make_depth_rgba(self, depth, d_min, d_max, rgba):
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
