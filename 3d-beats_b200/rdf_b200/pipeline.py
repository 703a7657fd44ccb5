"""Host-buffer front ends: the calls a user with frames in host memory makes.

HostBatchEvaluator  offline batches (test_on_saved_model.py shape, src/test_on_saved_model.py:46-58): depth frames in
                    pinned host memory -> label maps in pinned host memory, chunked and pipelined over three streams
                    (H2D | eval | D2H) so copies overlap compute.
LiveFramePipeline   the per-frame product path (src/3d_bz.py:437-465): H2D of one frame -> fused layered forest ->
                    cluster mean shift -> D2H of the K x 2 centroids, captured once as a CUDA graph and replayed.
"""
import numpy as np
import torch

from . import _capi
from .buffers import GPUArray, GpuBuffer
from .mean_shift import MeanShift


def pinned_like(shape, np_dtype):
    from .buffers import torch_dtype
    return torch.empty(tuple(shape), dtype=torch_dtype(np_dtype), pin_memory=True)


class HostBatchEvaluator:
    def __init__(self, evaluator, forest, frame_shape, chunk_frames=64, labels_reduce=1, scale_factor=1.):
        self.ev = evaluator
        self.forest = forest
        self.H, self.W = frame_shape
        self.r = labels_reduce
        self.scale = scale_factor
        self.chunk = int(chunk_frames)
        h, w = self.H // self.r, self.W // self.r
        self.depth_dev = [GPUArray((self.chunk, self.H, self.W), dtype=np.uint16) for _ in range(2)]
        self.labels_dev = [GPUArray((self.chunk, h, w), dtype=np.uint16) for _ in range(2)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))
        self.bytes_h2d = 0
        self.bytes_d2h = 0

    def run(self, depth_host, labels_host, prefill=65535):
        """depth_host: pinned torch uint16[N,H,W]; labels_host: pinned torch uint16[N,h,w] (fully overwritten:
        skipped pixels hold `prefill`, as after the reference's labels.fill(65535) + get_labels_forest + .get())."""
        N = depth_host.shape[0]
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        ev_in = [None, None]      # H2D of buffer b finished
        ev_run = [None, None]     # eval on buffer b finished
        ev_out = [None, None]     # D2H of buffer b finished (buffer reusable)
        self.bytes_h2d = self.bytes_d2h = 0
        for ci, n0 in enumerate(range(0, N, self.chunk)):
            b = ci & 1
            nb = min(self.chunk, N - n0)
            d_dev = self.depth_dev[b][:nb] if nb < self.chunk else self.depth_dev[b]
            l_dev = self.labels_dev[b][:nb] if nb < self.chunk else self.labels_dev[b]
            with torch.cuda.stream(self.s_in):
                if ev_run[b] is not None:
                    self.s_in.wait_event(ev_run[b])               # previous eval on this buffer consumed the depth
                d_dev.tensor.view(torch.int16).copy_(depth_host[n0:n0 + nb].view(torch.int16), non_blocking=True)
                ev_in[b] = torch.cuda.Event(); ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in[b])
                if ev_out[b] is not None:
                    self.s_run.wait_event(ev_out[b])              # previous D2H of this labels buffer finished
                l_dev.fill(prefill)
                self.ev.get_labels_forest(self.forest, d_dev, l_dev, labels_reduce=self.r, scale_factor=self.scale)
                ev_run[b] = torch.cuda.Event(); ev_run[b].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run[b])
                labels_host[n0:n0 + nb].view(torch.int16).copy_(l_dev.tensor.view(torch.int16), non_blocking=True)
                ev_out[b] = torch.cuda.Event(); ev_out[b].record(self.s_out)
            self.bytes_h2d += d_dev.nbytes
            self.bytes_d2h += l_dev.nbytes
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)
        return labels_host


class LiveFramePipeline:
    """One frame in, fingertip centroids out.  All device work of a frame is one CUDA-graph replay."""

    def __init__(self, layered_forest, num_rounds, variances, scale_factor=1., use_graph=True, upload=True, zero_copy_out=True):
        """upload: True = upload kernel reading the pinned frame (default), 'copy' = copy-engine H2D node, False = the frame is
        already in `depth_dev` (device-resident producer, e.g. a pre-processing kernel);
        zero_copy_out: the mean-shift kernel writes the K x 2 centroids straight into pinned host memory."""
        self.upload = upload
        self.zero_copy_out = zero_copy_out
        self.ldf = layered_forest
        self.rounds = int(num_rounds)
        self.scale = float(scale_factor)
        H, W = layered_forest.depth_dims
        h, w = layered_forest.labels_dims
        self.K = layered_forest.num_layered_classes
        self.depth_host = pinned_like((1, H, W), np.uint16)
        self.means_host = torch.empty((self.K, 2), dtype=torch.float64, pin_memory=True)
        self.depth_dev = GpuBuffer((1, H, W), np.uint16)
        self.labels_dev = GpuBuffer((1, h, w), np.uint16)
        self.ms = MeanShift()
        self.variances = GPUArray((len(variances),), dtype=np.float32)
        self.variances.set(np.ascontiguousarray(variances, dtype=np.float32))
        self.stream = torch.cuda.Stream()
        self.graph = None
        self.h2d_bytes = H * W * 2 if upload else 0
        self.d2h_bytes = self.K * 2 * 8
        # one eager pass: loads kernels, sets function attributes, allocates mean-shift scratch (nothing may allocate
        # during capture)
        with torch.cuda.stream(self.stream):
            self._enqueue()
        self.stream.synchronize()
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._enqueue()

    def _enqueue(self):
        if self.upload == 'copy':
            self.depth_dev.cu().tensor.view(torch.int16).copy_(self.depth_host.view(torch.int16), non_blocking=True)
        elif self.upload:
            # upload kernel (zero-copy read of the pinned frame): chains to the layered kernel by programmatic dependent launch
            import ctypes
            _capi.check(_capi.load().rdf_upload_frame(ctypes.c_void_p(self.depth_host.data_ptr()), _capi.dptr(self.depth_dev.cu()),
                                                      self.depth_host.numel() * 2, _capi.stream_ptr()))
        self.ldf.run(self.depth_dev, self.labels_dev, self.scale)
        if self.zero_copy_out:
            self.ms.run_async(self.rounds, self.labels_dev.cu(), self.K, self.variances, means_out=self.means_host)
        else:
            means = self.ms.run_async(self.rounds, self.labels_dev.cu(), self.K, self.variances)
            self.means_host.copy_(means.tensor, non_blocking=True)

    def submit(self):
        """Enqueue one frame (depth already written into self.depth_host)."""
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._enqueue()

    def run(self, depth_frame=None):
        """depth_frame: np.uint16[H,W] (or [1,H,W]); returns np.float64[K,2] centroids (x,y)."""
        if depth_frame is not None:
            self.depth_host.view(torch.int16).numpy()[...] = np.asarray(depth_frame, dtype=np.uint16).reshape(
                self.depth_host.shape).view(np.int16)
            if not self.upload:
                self.depth_dev.cu().tensor.view(torch.int16).copy_(self.depth_host.view(torch.int16))
        self.submit()
        self.stream.synchronize()
        return self.means_host.numpy().copy()
