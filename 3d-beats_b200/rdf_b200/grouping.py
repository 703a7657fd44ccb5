"""Hand grouping on the device: mirror of the reference's `CppGrouping` (src/cpp_grouping/cpp_grouping.pyx:8-26) over
rdf_group_hands.  The reference copies the 1/8-resolution depth image to the host, flood-fills it in C++, copies a coordinate
list back and scatters it into a stencil image (src/3d_bz.py:222-250); `make_groups_cu` produces that stencil image and the
group statistics in one launch without leaving the GPU."""
import numpy as np

from . import _capi
from .buffers import GPUArray, as_gpuarray


class CppGrouping:
    def __init__(self):
        self._lib = _capi.load()
        self._stencil = None
        self._g_info = None

    def make_groups_cu(self, depth_mm, stencil_out, g_info_out, pct_thresh):
        """depth_mm uint16[h,w] (device), stencil_out uint16[h,w] (device; 1 = right group, 2 = left group, 0 elsewhere),
        g_info_out float32[2,3] (device) = (size, centroid x, centroid y) per group.  Asynchronous on the current stream."""
        depth_mm, stencil_out, g_info_out = as_gpuarray(depth_mm), as_gpuarray(stencil_out), as_gpuarray(g_info_out)
        assert depth_mm.dtype == np.uint16 and stencil_out.dtype == np.uint16 and g_info_out.dtype == np.float32
        h, w = depth_mm.shape[-2:]
        assert stencil_out.size == h * w and g_info_out.size == 6
        _capi.check(self._lib.rdf_group_hands(_capi.dptr(depth_mm), w, h, float(pct_thresh), _capi.dptr(stencil_out),
                                              _capi.dptr(g_info_out), _capi.stream_ptr()))

    def make_groups(self, in_arr, coords_arr, group_info_arr, pct_thresh):
        """The reference's host-array signature (cpp_grouping.pyx:15): in_arr uint16[h,w], coords_arr int32[>=n,3] receives
        (y, x, group) rows - right group first, then left, each in raster order (the reference lists them in BFS order; the
        consumer only scatters them into an image, src/3d_bz.py:243-250) - group_info_arr float32[2,3]."""
        in_arr = np.ascontiguousarray(in_arr, dtype=np.uint16)
        h, w = in_arr.shape
        if self._stencil is None or self._stencil.shape != (h, w):
            self._stencil = GPUArray((h, w), dtype=np.uint16)
            self._g_info = GPUArray((2, 3), dtype=np.float32)
        dev = GPUArray((h, w), dtype=np.uint16)
        dev.set(in_arr)
        self.make_groups_cu(dev, self._stencil, self._g_info, pct_thresh)
        stencil = self._stencil.get()
        group_info_arr[...] = self._g_info.get()
        n = 0
        for g in (1, 2):
            ys, xs = np.nonzero(stencil == g)
            coords_arr[n:n + len(ys), 0] = ys
            coords_arr[n:n + len(ys), 1] = xs
            coords_arr[n:n + len(ys), 2] = g
            n += len(ys)
        return n
