"""Mean-shift mode finding per label class: mirror of the reference's src/cuda/mean_shift.py over rdf_mean_shift.

All rounds run in one launch of a thread-block cluster (csrc/rdf_meanshift.cu); the only host transfer is the final
K x 2 float64 result, where the reference makes 2 blocking D2H + 1 H2D copies per round (mean_shift.py:50-55)."""
import ctypes

import numpy as np
import torch

from . import _capi
from .buffers import GPUArray, as_gpuarray


class MeanShift:
    def __init__(self):
        self._lib = _capi.load()
        self.means = None           # float64[num_labels, 2] on the device, as in the reference
        self._workspace = None
        self._variances = None

    def _ensure(self, num_labels, dim_x, dim_y, num_images=1):
        shape = (num_labels, 2) if num_images == 1 else (num_images, num_labels, 2)
        if self.means is None or self.means.shape != shape:
            self.means = GPUArray(shape, dtype=np.float64)
        need = ctypes.c_size_t()
        _capi.check(self._lib.rdf_mean_shift_workspace_bytes(dim_x, dim_y, num_labels, ctypes.byref(need)))
        if self._workspace is None or self._workspace.nbytes < need.value:
            self._workspace = GPUArray(((need.value + 3) // 4,), dtype=np.uint32)

    def run_fingertips_async(self, num_rounds, labels, num_labels, variances, fingertip_idxes, labels_reduce, raw_depth, pp, fx, fy, plane,
                             z_out, means_copy=None, batch=False):
        """Mean shift + the product's fingertip read-out (src/3d_bz.py:458-462 + 503-522) in one launch (rdf_mean_shift_fingertips):
        z_out float64[(N,) len(fingertip_idxes)] and means_copy (optional) may be device arrays or pinned host tensors; raw_depth is the
        raw camera frame (device array or pinned host tensor).  Returns the device means like run_async."""
        return self.run_async(num_rounds, labels, num_labels, variances, batch=batch,
                              _fingertips=(fingertip_idxes, labels_reduce, raw_depth, pp, fx, fy, plane, z_out, means_copy))

    def run_async(self, num_rounds, labels, num_labels, variances, means_out=None, batch=False, _fingertips=None):
        """Enqueue the whole mean shift on the current stream; returns the device array float64[num_labels,2].
        batch=True: labels is uint16[N,h,w], one independent mean shift per image in the same launch (the two hands of the
        product frame), result float64[N,num_labels,2].
        means_out: optional pinned-host torch tensor float64[num_labels,2]; the kernel then writes the centroids straight into
        host memory (zero-copy over PCIe, 16 bytes per class), which saves the D2H copy node of a per-frame pipeline."""
        labels = as_gpuarray(labels)
        assert labels.dtype == np.uint16
        dim_y, dim_x = labels.shape[-2:]
        if isinstance(variances, np.ndarray):
            if self._variances is None or self._variances.shape != variances.shape:
                self._variances = GPUArray(variances.shape, dtype=np.float32)
            self._variances.set(np.ascontiguousarray(variances, dtype=np.float32))
            variances = self._variances
        variances = as_gpuarray(variances)
        assert variances.dtype == np.float32 and variances.size >= num_labels
        num_images = int(labels.shape[0]) if batch else 1
        assert labels.size == num_images * dim_x * dim_y
        self._ensure(num_labels, dim_x, dim_y, num_images)
        if means_out is not None:
            assert means_out.is_pinned() and means_out.dtype == torch.float64 and means_out.numel() == 2 * num_labels * num_images
            out_ptr = ctypes.c_void_p(means_out.data_ptr())          # unified addressing: pinned host memory is device-visible
        else:
            out_ptr = _capi.dptr(self.means)
        if _fingertips is not None:
            from .points_ops import _out_ptr
            idxes, r, raw, pp, fx, fy, plane, z_out, means_copy = _fingertips
            assert means_out is None
            if not (isinstance(raw, torch.Tensor) and not raw.is_cuda):
                raw = as_gpuarray(raw)
                assert raw.dtype == np.uint16
            plane = as_gpuarray(plane)
            assert plane.dtype == np.float32 and plane.size == 16
            H, W = raw.shape[-2:]
            n = len(idxes)
            idx = (ctypes.c_int * n)(*[int(i) for i in idxes])
            _capi.check(self._lib.rdf_mean_shift_fingertips(
                _capi.dptr(labels), num_images, dim_x, dim_y, int(num_labels), _capi.dptr(variances), int(num_rounds), out_ptr,
                _capi.dptr(self._workspace), self._workspace.nbytes, idx, n, int(r), _out_ptr(raw), W, H, float(pp[0]), float(pp[1]),
                float(fx), float(fy), _capi.dptr(plane), _out_ptr(z_out), None if means_copy is None else _out_ptr(means_copy),
                _capi.stream_ptr()))
            return self.means
        _capi.check(self._lib.rdf_mean_shift_batch(_capi.dptr(labels), num_images, dim_x, dim_y, int(num_labels), _capi.dptr(variances),
                                                   int(num_rounds), out_ptr, _capi.dptr(self._workspace),
                                                   self._workspace.nbytes, _capi.stream_ptr()))
        return self.means if means_out is None else means_out

    def run(self, num_rounds, labels, num_labels, variances):
        """src/cuda/mean_shift.py:19-59: returns np.float64[num_labels, 2] = (x, y); NaN rows for classes without pixels."""
        return self.run_async(num_rounds, labels, num_labels, variances).get()
