"""Host-buffer front ends: the calls a user with frames in host memory makes.

HostBatchEvaluator  offline batches (test_on_saved_model.py shape, src/test_on_saved_model.py:46-58): depth frames in
                    pinned host memory -> label maps in pinned host memory, chunked and pipelined over three streams
                    (H2D | eval | D2H) so copies overlap compute.
LiveFramePipeline   the per-frame product path (src/3d_bz.py:437-465): H2D of one frame -> fused layered forest ->
                    cluster mean shift -> D2H of the K x 2 centroids, captured once as a CUDA graph and replayed.
HandsFramePipeline  the whole compute of one product frame (App_3d_bz.tick + run_per_hand_pipeline x 2, src/3d_bz.py:129-300,
                    387-522): raw camera frame in pinned host memory -> conditioning -> hand grouping -> per hand: stencil,
                    layered forest, mean shift, fingertip depths -> centroids + fingertip z of both hands in pinned host memory.
"""
import numpy as np
import torch

from . import _capi
from .buffers import GPUArray, GpuBuffer
from .mean_shift import MeanShift
from .grouping import CppGrouping
from .points_ops import PointsOps


def pinned_like(shape, np_dtype):
    from .buffers import torch_dtype
    return torch.empty(tuple(shape), dtype=torch_dtype(np_dtype), pin_memory=True)


class HostBatchEvaluator:
    def __init__(self, evaluator, forest, frame_shape, chunk_frames=128, labels_reduce=1, scale_factor=1., buffers=3, ramp=8):
        """buffers: device chunk buffers in rotation (3: the H2D of chunk i+1, the evaluation of chunk i and the D2H of chunk i-1
        each own one, so neither copy direction ever waits for the other's buffer).
        ramp: frames of the first chunk (0 / False: plain chunks).  The chunks double from there up to chunk_frames and shrink
        again at the end of the batch, so that the pipeline fills with the upload of `ramp` frames and drains with the download
        of `ramp` frames whatever the chunk size (the evaluation cannot start before the first upload ends, nor the last
        download before the last evaluation), while the bulk of the batch runs in large chunks with few stream hand-overs."""
        self.ramp = int(ramp) if ramp else 0
        self.ev = evaluator
        self.forest = forest
        self.H, self.W = frame_shape
        self.r = labels_reduce
        self.scale = scale_factor
        self.chunk = int(chunk_frames)
        self.nbuf = max(2, int(buffers))
        h, w = self.H // self.r, self.W // self.r
        self.depth_dev = [GPUArray((self.chunk, self.H, self.W), dtype=np.uint16) for _ in range(self.nbuf)]
        self.labels_dev = [GPUArray((self.chunk, h, w), dtype=np.uint16) for _ in range(self.nbuf)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))
        self.bytes_h2d = 0
        self.bytes_d2h = 0

    def chunk_sizes(self, N):
        """frames per chunk, in order, for a batch of N frames"""
        def doubling(c):
            steps, s = [], self.ramp
            while 0 < s < c:
                steps.append(s)
                s *= 2
            return steps
        c = self.chunk
        steps = doubling(c)
        while steps and 2 * sum(steps) + c > N:                          # a short batch ramps up to a smaller chunk
            c = max(self.ramp, c // 2)
            steps = doubling(c)
        if not steps:
            return [min(c, N - n0) for n0 in range(0, N, c)]
        full, rest = divmod(N - 2 * sum(steps), c)
        return steps + [c] * full + ([rest] if rest else []) + steps[::-1]

    def run(self, depth_host, labels_host, prefill=65535, copy_only=False):
        """depth_host: pinned torch uint16[N,H,W]; labels_host: pinned torch uint16[N,h,w] (fully overwritten:
        skipped pixels hold `prefill`, as after the reference's labels.fill(65535) + get_labels_forest + .get()).
        copy_only: the same chunked H2D / D2H traffic with NO kernel in between - the control experiment that separates what the
        host <-> device copies cost from what the evaluation costs (labels_host then receives whatever the buffers held)."""
        N = depth_host.shape[0]
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        nbuf = self.nbuf
        ev_in = [None] * nbuf     # H2D of buffer b finished
        ev_run = [None] * nbuf    # eval on buffer b finished
        ev_out = [None] * nbuf    # D2H of buffer b finished (buffer reusable)
        self.bytes_h2d = self.bytes_d2h = 0
        sizes = self.chunk_sizes(N)
        starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).tolist() if sizes else []
        for ci, (n0, nb) in enumerate(zip(starts, sizes)):
            b = ci % nbuf
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
                    self.s_run.wait_event(ev_out[b])              # previous D2H of this labels buffer (and its re-fill) finished
                elif not copy_only:
                    l_dev.fill(prefill)                           # first use of the buffer in this run
                if not copy_only:
                    self.ev.get_labels_forest(self.forest, d_dev, l_dev, labels_reduce=self.r, scale_factor=self.scale)
                ev_run[b] = torch.cuda.Event(); ev_run[b].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run[b])
                labels_host[n0:n0 + nb].view(torch.int16).copy_(l_dev.tensor.view(torch.int16), non_blocking=True)
                if not copy_only and ci + nbuf < len(sizes):
                    # pre-fill for the buffer's next chunk here, behind the download and beside the running evaluation, instead
                    # of in front of that chunk's evaluation (the kernels never write skipped pixels, tree_eval.cu:81-89)
                    nxt = sizes[ci + nbuf]
                    (self.labels_dev[b][:nxt] if nxt < self.chunk else self.labels_dev[b]).fill(prefill)
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


class HandsFramePipeline:
    """One raw camera frame in, both hands' fingertip centroids and plane-space depths out: the device work that
    `App_3d_bz.tick` (src/3d_bz.py:129-300) and its two `run_per_hand_pipeline` calls (:387-522) issue as ~45 launches, 8 frame-sized
    copies, one D2H + C++ flood fill + H2D round trip and 28 blocking mean-shift transfers is SIX launches here - upload,
    conditioning, grouping, stencil (both hands), layered forest (both hands), mean shift + read-out (both hands) -
    chained by programmatic dependent launch and captured once as a CUDA graph.  Defaults are the product's settings
    (src/3d_bz.py:49-113).

    batch_hands=False keeps one layered / mean-shift / read-out launch per hand (on two concurrent branches of the graph when
    concurrent_hands is set): the reference's structure, kept for comparison (measured: 106-114 us against the batched chain)."""

    def __init__(self, layered_forest, variances, pp, focal, plane, fx=None, fy=None, num_rounds=6, plane_z_threshold=40.,
                 gauss_sigma=2.0, k_size=5, mm_level=3, group_min_size=0.06, fingertip_idxes=(2, 3, 4, 5, 6), scale_factor=1.,
                 use_graph=True, batch_hands=True, concurrent_hands=True, upload='kernel', fused_readout=True):
        """upload: 'kernel' = upload kernel into depth_raw, then conditioning; 'fused' = the conditioning kernel (and the read-out)
        read the pinned host frame themselves (zero-copy over PCIe; measured 4x slower: 2-byte tile reads over PCIe)."""
        self.upload = upload
        self.batch_hands = bool(batch_hands)
        self.fused_readout = bool(fused_readout) and self.batch_hands     # read-out inside the mean-shift launch
        self.ldf = layered_forest
        H, W = layered_forest.depth_dims
        h, w = layered_forest.labels_dims
        self.H, self.W, self.h, self.w = H, W, h, w
        self.K = layered_forest.num_layered_classes
        self.rounds = int(num_rounds)
        self.pp = (float(pp[0]), float(pp[1]))
        self.focal = float(focal)
        self.fx = float(focal if fx is None else fx)
        self.fy = float(focal if fy is None else fy)
        self.thresh = float(plane_z_threshold)
        self.sigma, self.k_size, self.mm_level = float(gauss_sigma), int(k_size), int(mm_level)
        self.group_min_size = float(group_min_size)
        self.fingertips = tuple(int(i) for i in fingertip_idxes)
        self.scale = float(scale_factor)
        self.hands = ((1, False), (2, True))                      # (group id, flip_x): right hand, left hand (src/3d_bz.py:281-285)
        nh, nf = len(self.hands), len(self.fingertips)
        mh, mw = H >> self.mm_level, W >> self.mm_level
        self.ops = PointsOps()
        self.grouping = CppGrouping()
        # host side of the frame (pinned): what the camera thread writes and the consumer reads
        self.depth_host = pinned_like((H, W), np.uint16)
        self.means_host = torch.empty((nh, self.K, 2), dtype=torch.float64, pin_memory=True)
        self.z_host = torch.empty((nh, nf), dtype=torch.float64, pin_memory=True)
        # device buffers (names follow src/3d_bz.py:80-100)
        self.depth_raw = GpuBuffer((H, W), np.uint16)             # depth_image_cpu's device twin: the read-out looks z up here
        self.depth_image = GpuBuffer((H, W), np.uint16)
        self.depth_image_mm = GpuBuffer((mh, mw), np.uint16)
        self.depth_image_mm_groups_2 = GpuBuffer((mh, mw), np.uint16)   # stencil before grow_groups
        self.g_info = GpuBuffer((2, 3), np.float32)
        self.depth_image_hands = GpuBuffer((nh, H, W), np.uint16)       # depth_image_2 of each hand
        self.labels_images = GpuBuffer((nh, h, w), np.uint16)           # labels_image of each hand (un-mirrored)
        self.layer_images = [GpuBuffer((nh, h, w), np.uint16) for _ in range(layered_forest.num_models)]
        self.mean_shift = [MeanShift() for _ in range(nh)]
        self.plane = GPUArray((4, 4), dtype=np.float32)
        self.set_plane(plane)
        self.variances = GPUArray((len(variances),), dtype=np.float32)
        self.variances.set(np.ascontiguousarray(variances, dtype=np.float32))
        self.stream = torch.cuda.Stream()
        self.side = [torch.cuda.Stream() for _ in range(nh - 1)] if (concurrent_hands and not batch_hands) else []
        self.graph = None
        self.h2d_bytes = H * W * 2
        self.d2h_bytes = nh * (self.K * 2 + nf) * 8
        self.kernels_per_frame = (4 if upload == 'kernel' else 3) + ((2 if self.fused_readout else 3) if batch_hands else 3 * nh)
        with torch.cuda.stream(self.stream):
            self._enqueue()                                       # eager pass: function attributes, mean-shift scratch
        self.stream.synchronize()
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._enqueue()

    def set_plane(self, plane):
        """CalibratedPlane.get_mat() (src/calibrated_plane.py:33-35); lives in device memory, so a re-calibration does not
        invalidate the captured graph."""
        self.plane.set(np.ascontiguousarray(plane, dtype=np.float32).reshape(4, 4))

    def labels_image(self, i):
        """composite label image of hand i, uint16[h,w] (device view)"""
        return self.labels_images.cu()[i]

    def _raw(self):
        return self.depth_host if self.upload == 'fused' else self.depth_raw

    def _hand(self, i):
        g_id, flip = self.hands[i]
        self.ldf.run(self.depth_image_hands.cu()[i], self.labels_images.cu()[i], self.scale, composite_flip_x=[flip],
                     label_images=[b.cu()[i] for b in self.layer_images])
        means = self.mean_shift[i].run_async(self.rounds, self.labels_images.cu()[i], self.K, self.variances)
        self.ops.fingertip_z(means, self.fingertips, self.ldf.labels_reduce, self._raw(), self.pp, self.fx, self.fy, self.plane,
                             self.z_host[i], means_copy=self.means_host[i])

    def _enqueue(self):
        import ctypes
        if self.upload != 'fused':
            _capi.check(_capi.load().rdf_upload_frame(ctypes.c_void_p(self.depth_host.data_ptr()), _capi.dptr(self.depth_raw.cu()),
                                                      self.depth_host.numel() * 2, _capi.stream_ptr()))
        self.ops.condition_depth(self._raw(), self.depth_image, self.depth_image_mm, self.pp, self.focal, self.plane, self.thresh,
                                 self.sigma, self.k_size, self.mm_level)
        self.grouping.make_groups_cu(self.depth_image_mm, self.depth_image_mm_groups_2, self.g_info, self.group_min_size)
        self.ops.stencil_hands(self.depth_image, self.depth_image_mm_groups_2, self.mm_level, self.hands, self.depth_image_hands,
                               grow=True)
        if self.batch_hands:
            self.ldf.run(self.depth_image_hands, self.labels_images, self.scale, composite_flip_x=[f for _, f in self.hands],
                         label_images=self.layer_images)
            if self.fused_readout:
                self.mean_shift[0].run_fingertips_async(self.rounds, self.labels_images.cu(), self.K, self.variances, self.fingertips,
                                                        self.ldf.labels_reduce, self._raw(), self.pp, self.fx, self.fy, self.plane,
                                                        self.z_host, means_copy=self.means_host, batch=True)
                return
            means = self.mean_shift[0].run_async(self.rounds, self.labels_images.cu(), self.K, self.variances, batch=True)
            self.ops.fingertip_z(means, self.fingertips, self.ldf.labels_reduce, self._raw(), self.pp, self.fx, self.fy, self.plane,
                                 self.z_host, means_copy=self.means_host)
            return
        main = torch.cuda.current_stream()
        if self.side:
            fork = torch.cuda.Event()
            fork.record(main)
            for i, s in enumerate(self.side, start=1):
                s.wait_event(fork)
                with torch.cuda.stream(s):
                    self._hand(i)
            self._hand(0)
            for s in self.side:
                join = torch.cuda.Event()
                join.record(s)
                main.wait_event(join)
        else:
            for i in range(len(self.hands)):
                self._hand(i)

    def submit(self):
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._enqueue()

    def run(self, depth_frame=None):
        """depth_frame: np.uint16[H,W] raw camera frame (None = already written into self.depth_host).
        Returns (means float64[2,K,2], z float64[2,num_fingertips]); NaN z = the reference's reset_positions()."""
        if depth_frame is not None:
            self.depth_host.view(torch.int16).numpy()[...] = np.asarray(depth_frame, dtype=np.uint16).reshape(self.H, self.W).view(np.int16)
        self.submit()
        self.stream.synchronize()
        return self.means_host.numpy().copy(), self.z_host.numpy().copy()
