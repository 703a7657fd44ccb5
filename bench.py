#!/usr/bin/env python
"""bench.py - forest-eval throughput (BASELINE.json metric) on N GPUs of one node, one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path on the host cores

Workload (default `cfg3`, BASELINE.json configs[2], the configuration the Mpixels/s metric is quoted on):
4096 synthetic 848x480 uint16 depth frames (dense-smooth), random-init 4-tree depth-20 forest, 4 classes,
sharded by frame over the ranks with no data-path collective (strong scaling: the 4096 frames are split).
A step = one pass of get_labels_forest over the rank's frames, inputs resident in HBM.  Extra objects on the line:
  e2e          same metric through the host-buffer API (pinned host frames -> label maps in pinned host memory),
               H2D and D2H inside the timed region
  roofline     algorithmic bytes (4 + T*(32*D + 4*C) B per pixel) / kernel time vs the measured HBM peak
  latency      BASELINE.json configs[1]: one 848x480 frame -> 2-layer stacked forest + 6-round mean shift, p50/p95/p99
  ref_gpu      the reference's own kernels (compiled unchanged for sm_100a) on the same GPU, same inputs (sub-batch)
  cpu_baseline the C oracle on the host cores over a bounded sample (rank 0, N=1 only)
  train_cfg4   BASELINE.json configs[3]: one level of the training split search (bucket + histogram + pick-best)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (frames, W, H, T, D, C, frame kind)
    'cfg1': (1, 848, 480, 3, 16, 4, 'dense-smooth'),
    'cfg3': (4096, 848, 480, 4, 20, 4, 'dense-smooth'),
    'cfg3-noise': (4096, 848, 480, 4, 20, 4, 'dense-noise'),
    'cfg5': (1, 1280, 720, 8, 24, 4, 'dense-smooth'),
    'cfg5-noise': (1, 1280, 720, 8, 24, 4, 'dense-noise'),
}
KIND_ID = {'dense-smooth': 0, 'dense-noise': 1, 'live-mask': 2}


def b_alg_per_pixel(T, D, C):
    """2 B centre depth + 2 B label + per tree D x (28 B header + 2 x 2 B probes) + 4C B leaf pdf."""
    return 4 + T * (32 * D + 4 * C)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def host_cores():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so OpenMP's default is not it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(T, D, C, W, H, kind, budget_s=12.0, max_frames=64, seed=1234):
    """C oracle (oracle/rdf_oracle.c) on all host threads over a bounded sample of the workload's frames."""
    from rdf_b200 import synth
    from oracle import c_oracle as co
    forest = synth.hash_forest(T, D, C, seed=seed)
    cores = host_cores()
    depth1 = synth.depth_frames(kind, 1, H, W, seed=seed)
    lab1 = np.full((1, H, W), 65535, np.uint16)
    t0 = time.perf_counter()
    co.eval_forest(forest, depth1, lab1, nthreads=cores)
    t1 = time.perf_counter() - t0
    frames = int(max(1, min(max_frames, budget_s / max(t1, 1e-6))))
    depth = synth.depth_frames(kind, frames, H, W, seed=seed)
    labels = np.full((frames, H, W), 65535, np.uint16)
    t0 = time.perf_counter()
    co.eval_forest(forest, depth, labels, nthreads=cores)
    dt = time.perf_counter() - t0
    return {'value': frames * H * W / dt / 1e6, 'unit': 'Mpixels/s', 'cores': cores, 'kind': 'port',
            'sample': f'{frames} of the workload\'s {W}x{H} frames, same forest, C oracle with OpenMP on {cores} threads, {dt:.1f} s'}, forest


def run_reference_arm(args):
    """--impl reference: the reference has no CPU implementation of this path (its implementation IS CUDA kernels), so
    this arm times the C restatement (oracle/rdf_oracle.c, kind 'port') on all host threads.  The reference's own kernels
    on the B200 are timed inside the default arm as `ref_gpu`."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from rdf_b200 import synth
    from oracle import c_oracle as co
    frames, W, H, T, D, C, kind = WORKLOADS[args.workload]
    forest = synth.hash_forest(T, D, C, seed=args.seed)
    cores = host_cores()
    sample = max(1, min(frames, args.ref_frames))
    depth = synth.depth_frames(kind, sample, H, W, seed=args.seed)
    labels = np.full((sample, H, W), 65535, np.uint16)
    for _ in range(args.warmup):
        co.eval_forest(forest, depth[:1], labels[:1], nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.eval_forest(forest, depth, labels, nthreads=cores)
    dt = time.perf_counter() - t0
    value = sample * H * W * args.steps / dt / 1e6
    line = {
        'impl': 'reference', 'metric': 'forest_eval_mpixels_per_s', 'value': value, 'unit': 'Mpixels/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_desc(args.workload), 'step': f'bounded sample: {sample} frames per step on the host cores'},
        'cpu_baseline': {'value': value, 'unit': 'Mpixels/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} frames x {args.steps} steps, C oracle (OpenMP, {cores} threads)'},
        'e2e': {'value': value, 'unit': 'Mpixels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_desc(name):
    frames, W, H, T, D, C, kind = WORKLOADS[name]
    return f'{name}: {frames} synthetic {W}x{H} uint16 {kind} frames, random-init {T}-tree depth-{D} forest, {C} classes'


def latency_cfg2(iters=1000, warm=100):
    """BASELINE configs[1]: one 848x480 live-mask frame through the 2-layer stacked forest (hand/background -> 10 finger
    parts, labels_reduce 2) + 6-round mean shift.  Two variants, each ONE CUDA-graph replay per frame, timed as host wall
    time (perf_counter around replay + stream sync) and as device time (CUDA events on the pipeline's stream):
      e2e       frame in pinned host memory -> upload kernel (zero-copy read over PCIe) -> layered kernel -> mean-shift kernel
                writing the centroids straight into pinned host memory, the three kernels chained by programmatic dependent
                launch (the call a user of run_live_layered.py / 3d_bz.py makes per frame); also measured with a copy-engine
                H2D node instead of the upload kernel
      resident  frame already in HBM (what the product has after its own pre-processing kernels, src/3d_bz.py:394-420):
                layered kernel -> mean-shift kernel -> centroids in pinned host memory."""
    import torch
    from rdf_b200 import synth
    from rdf_b200 import decision_tree as dt
    from rdf_b200.pipeline import LiveFramePipeline
    from oracle import numpy_oracle as no
    H, W, r = 480, 848, 2
    forests, cfg, variances = synth.layered_cfg2()
    for layer, f in zip(cfg['layers'], forests):
        m = dt.DecisionForest(f.shape[0], 16, (f.shape[2] - 7) // 2)
        m.forest_cu.set(f)
        layer['model'] = m
    cfg['root'] = '.'
    ldf = dt.LayeredDecisionForest(cfg, (H, W), r)
    depth = synth.depth_frames('live-mask', 1, H, W)
    valid_px = int(((depth[0, ::r, ::r] != 65535) & (depth[0, ::r, ::r] != 0)).sum())

    def measure(pipe):
        means = pipe.run(depth)
        wall, devt = [], []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warm + iters):
            t0 = time.perf_counter()
            with torch.cuda.stream(pipe.stream):
                e0.record()
            pipe.submit()
            with torch.cuda.stream(pipe.stream):
                e1.record()
            pipe.stream.synchronize()
            t1 = time.perf_counter()
            if i >= warm:
                wall.append((t1 - t0) * 1e6)
                devt.append(e0.elapsed_time(e1) * 1e3)
        # wall time without the two event records (what a caller that does not time on the device pays)
        plain = []
        for i in range(warm + iters):
            t0 = time.perf_counter()
            pipe.submit()
            pipe.stream.synchronize()
            t1 = time.perf_counter()
            if i >= warm:
                plain.append((t1 - t0) * 1e6)
        wall, devt, plain = np.array(wall), np.array(devt), np.array(plain)
        pc = lambda a, q: float(np.percentile(a, q))
        return means, {'p50_us': pc(plain, 50), 'p95_us': pc(plain, 95), 'p99_us': pc(plain, 99),
                       'p50_us_with_event_records': pc(wall, 50), 'device_p50_us': pc(devt, 50), 'device_p99_us': pc(devt, 99),
                       'h2d_bytes': pipe.h2d_bytes, 'd2h_bytes': pipe.d2h_bytes}

    pipe = LiveFramePipeline(ldf, 6, variances, scale_factor=1.0)              # upload kernel -> layered -> mean shift (PDL chain)
    means, e2e = measure(pipe)
    pipe_copy = LiveFramePipeline(ldf, 6, variances, scale_factor=1.0, upload='copy')   # copy-engine H2D node instead
    means_copy, e2e_copy = measure(pipe_copy)
    pipe_res = LiveFramePipeline(ldf, 6, variances, scale_factor=1.0, upload=False)
    means_res, resident = measure(pipe_res)
    # parity of exactly what was timed: centroids vs the NumPy oracle on the oracle's own composite map
    exp_comp, _ = no.layered_run(forests, [(None, None), (0, 1)], cfg['conditions'], depth[0], r, 1.0)
    exp_means = no.mean_shift(exp_comp, ldf.num_layered_classes, variances, 6)
    comp_ok = bool(np.array_equal(pipe.labels_dev.cu().get()[0], exp_comp))
    means_ok = bool(np.array_equal(np.isnan(means), np.isnan(exp_means)) and np.nanmax(np.abs(means - exp_means)) <= 1e-5 and
                    np.array_equal(np.nan_to_num(means), np.nan_to_num(means_res)) and
                    np.array_equal(np.nan_to_num(means), np.nan_to_num(means_copy)))

    def _upload(pp):
        import ctypes
        from rdf_b200 import _capi
        _capi.check(_capi.load().rdf_upload_frame(ctypes.c_void_p(pp.depth_host.data_ptr()), _capi.dptr(pp.depth_dev.cu()),
                                                  pp.depth_host.numel() * 2, _capi.stream_ptr()))

    # per-stage device time: each stage captured alone as its own CUDA graph and replayed back to back
    def stage(fn, n=300):
        with torch.cuda.stream(pipe.stream):
            fn()
        pipe.stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=pipe.stream):
            fn()
        with torch.cuda.stream(pipe.stream):
            for _ in range(20):
                g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                g.replay()
            b.record()
        pipe.stream.synchronize()
        return a.elapsed_time(b) * 1e3 / n
    stages = {
        'h2d_copy_engine_us': stage(lambda: pipe.depth_dev.cu().tensor.view(torch.int16).copy_(pipe.depth_host.view(torch.int16), non_blocking=True)),
        'h2d_upload_kernel_us': stage(lambda: _upload(pipe)),
        'layered_kernel_us': stage(lambda: pipe.ldf.run(pipe.depth_dev, pipe.labels_dev, pipe.scale)),
        'mean_shift_kernel_us': stage(lambda: pipe.ms.run_async(pipe.rounds, pipe.labels_dev.cu(), pipe.K, pipe.variances, means_out=pipe.means_host)),
        'note': 'each stage replayed alone as a 1-node CUDA graph, back to back; includes per-graph launch latency',
    }
    return {
        'workload': 'cfg2: one 848x480 live-mask frame, L1 (T3 D16 C3) -> L2 (T3 D16 C11, gated by L1==1), labels_reduce 2, '
                    'mean shift 6 rounds over 11 classes; one CUDA-graph replay per frame',
        'p50_us': e2e['p50_us'], 'p95_us': e2e['p95_us'], 'p99_us': e2e['p99_us'], 'device_p50_us': e2e['device_p50_us'],
        'e2e_host_frame': e2e, 'e2e_host_frame_copy_engine': e2e_copy, 'resident_frame': resident, 'iters': iters,
        'evaluated_pixels': valid_px,
        'labelled_classes': int(np.isfinite(means[:, 0]).sum()), 'kernels_per_frame': 2,
        'parity': {'composite_bit_exact_vs_oracle': comp_ok, 'centroids_within_1e-5': means_ok}, 'stages': stages,
    }


def train_cfg4(level=8, iters=2):
    """BASELINE configs[3] on this rank's GPU: one level of the split search over 42 dense-smooth 848x480 frames (17.1 M labelled
    pixels), 2000 features x 64 thresholds, C = 4, 2^level active nodes: bucket the pixels by node + histogram + pick-best.
    A feature sub-block is checked against the C oracle on two frames."""
    import ctypes
    import torch
    from rdf_b200 import _capi, synth
    from oracle import c_oracle as co
    lib = _capi.load()
    N, H, W, C, F, NT, D = 42, 480, 848, 4, 2000, 64, 16
    S = 1 << level
    depth_np = synth.depth_frames('dense-smooth', N, H, W)
    labels_np = synth.train_labels(N, H, W)
    nodes_np = synth.random_node_assignment(labels_np, level)
    off_np, th_np = synth.random_proposals(F, NT)
    depth = torch.from_numpy(depth_np.view(np.int16)).cuda()
    labels = torch.from_numpy(labels_np.view(np.int16)).cuda()
    nodes = torch.from_numpy(nodes_np).cuda()
    offsets, thresholds = torch.from_numpy(off_np).cuda(), torch.from_numpy(th_np).cuda()
    slot = torch.arange(S, dtype=torch.int32, device='cuda')
    parent = torch.zeros((1 << D, C), dtype=torch.int64, device='cuda')
    parent.view(-1).index_add_(0, nodes.view(-1).long() * C + labels.view(-1).long(), torch.ones(N * H * W, dtype=torch.int64, device='cuda'))
    next_counts = torch.zeros_like(parent)
    best_gain = torch.full((1 << D,), -1.0, dtype=torch.float32, device='cuda')
    tree = torch.zeros(((1 << D) - 1, 7 + 2 * C), dtype=torch.float32, device='cuda')
    hist = torch.zeros((S, F, NT + 1, C), dtype=torch.int32, device='cuda')
    need = ctypes.c_size_t()
    _capi.check(lib.rdf_train_bucket_workspace_bytes(N * H * W, S, ctypes.byref(need)))
    ws = torch.zeros(((need.value + 3) // 4,), dtype=torch.int32, device='cuda')
    st = _capi.stream_ptr

    def one_level():
        _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes), N * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
        hist.zero_()
        _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(ws), S, _capi.dptr(offsets),
                                                _capi.dptr(thresholds), F, NT, C, _capi.dptr(hist), st()))
        _capi.check(lib.rdf_train_pick_best(S, _capi.dptr(slot), _capi.dptr(slot), _capi.dptr(parent), _capi.dptr(hist), S,
                                            _capi.dptr(offsets), _capi.dptr(thresholds), F, NT, C, level, D, _capi.dptr(tree),
                                            _capi.dptr(next_counts), _capi.dptr(best_gain), st()))
    one_level()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        one_level()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nf = 6                                                       # parity: 6 features x 64 thresholds on 2 frames vs the C oracle
    h2 = torch.zeros((S, nf, NT + 1, C), dtype=torch.int32, device='cuda')
    nodes2 = nodes[:2].contiguous()
    _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes2), 2 * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
    _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth[:2]), _capi.dptr(labels[:2]), 2, W, H, _capi.dptr(ws), S, _capi.dptr(offsets[:nf]),
                                            _capi.dptr(thresholds[:nf]), nf, NT, C, _capi.dptr(h2), st()))
    torch.cuda.synchronize()
    exp = co.train_hist(depth_np[:2], labels_np[:2], nodes_np[:2], np.arange(S, dtype=np.int32), S, off_np[:nf], th_np[:nf], C)
    px = N * H * W
    return {'workload': f'cfg4: {px} labelled pixels (42 frames 848x480), {F} features x {NT} thresholds, C={C}, level {level} ({S} nodes)',
            'ms_per_level': ms, 'g_feature_evals_per_s': px * F / ms / 1e6, 'algorithmic_GBps': (px * 8 + px * F * 8) / ms / 1e6,
            'histogram_bit_exact_vs_c_oracle': bool(np.array_equal(h2.cpu().numpy().view(np.uint32), exp)),
            'what': 'rdf_train_bucket + rdf_train_hist_bucketed + rdf_train_pick_best, inputs resident'}


def hands_frame(iters=500):
    """SURVEY 8(f) ranks 1, 2, 4 around configs[1]: one WHOLE product frame (raw camera frame in pinned host memory -> plane clip +
    zero-aware gaussian + 1/8 image -> hand grouping -> both hands: stencil, 2-layer forest, mean shift, fingertip depths -> pinned
    host memory), seven launches in one CUDA-graph replay, checked against the oracles; beside it the reference's own kernels and
    host sequence on the same GPU (tools/bench_hands_frame.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_hands_frame', os.path.join(ROOT, 'tools', 'bench_hands_frame.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.run(iters=iters)


def ref_gpu_rate(forest_canon, depth_dev, frames, H, W, steps=2):
    """The reference's evaluate_image_using_forest (compiled unchanged, its own launch geometry) on a sub-batch."""
    import torch
    from oracle import ref_kernels as rk
    if not rk.available():
        return None
    sub = depth_dev.tensor[:frames]
    labels = torch.full((frames, H, W), -1, dtype=torch.int16, device=sub.device).view(torch.uint16)
    rk.eval_forest(forest_canon.tensor, sub, labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rk.eval_forest(forest_canon.tensor, sub, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': frames * H * W / ms / 1e3, 'unit': 'Mpixels/s', 'frames': frames, 'ms_per_pass': ms,
            'what': 'reference src/cuda/tree_eval.cu:evaluate_image_using_forest compiled unchanged for sm_100a, reference launch '
                    'geometry (block (1024//T, T)), same frames and forest, inputs resident', 'labels': labels}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=0, help='override the number of frames (debug)')
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--ref-frames', type=int, default=4, help='frames per step of the --impl reference CPU arm')
    ap.add_argument('--no-extras', action='store_true', help='skip latency / ref_gpu / cpu_baseline / e2e legs')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--latency-only', action='store_true', help='only the cfg2 frame-latency leg (profiling aid)')
    ap.add_argument('--latency-iters', type=int, default=1000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    import torch
    from rdf_b200 import _capi, dist as rdist
    from rdf_b200 import decision_tree as dt
    from rdf_b200.pipeline import HostBatchEvaluator, pinned_like

    rank, world, local = rdist.init_from_env()
    host_cpus_bound = 0
    if world > 1 and not os.environ.get('RDF_NO_NUMA_BIND'):
        # before any pinned host buffer is allocated (first touch decides the NUMA node); single-rank runs keep every core for the
        # CPU baseline leg
        host_cpus_bound = rdist.bind_host_to_gpu(local)
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: the product path has no CPU fallback'
    torch.cuda.set_device(local)
    if args.latency_only:
        if rank == 0:
            print(json.dumps({'latency': latency_cfg2(iters=args.latency_iters, warm=min(100, args.latency_iters))}), flush=True)
        return
    lib = _capi.load()

    frames, W, H, T, D, C, kind = WORKLOADS[args.workload]
    if args.frames:
        frames = args.frames
    f0, f1 = rdist.shard_range(frames, rank, world)
    if frames == 1:                                # single-frame workloads do not shard: every rank runs a replica
        f0, f1 = 0, 1
    my_frames = f1 - f0

    # ---- inputs, generated on the device (bit-exact twins of rdf_b200/synth.py) ----
    forest = dt.DecisionForest(T, D, C)
    _capi.check(lib.rdf_synth_forest(_capi.dptr(forest.forest_cu), T, D, C, args.seed, _capi.stream_ptr()))
    # Single-frame workloads (cfg1, cfg5: replicas only) would re-walk the same tree paths out of L2 every step, so each step
    # gets its own frame (frame index = step) and L2 is flushed (256 MB written) before it, outside the per-step events.
    single = frames == 1
    ring = (args.steps + args.warmup) if single else my_frames
    depth = dt.cu_array.GPUArray((ring, H, W), dtype=np.uint16)
    _capi.check(lib.rdf_synth_depth(_capi.dptr(depth), KIND_ID[kind], ring, W, H, args.seed, f0, _capi.stream_ptr()))
    labels = dt.cu_array.GPUArray((ring, H, W), dtype=np.uint16).fill(65535)
    ev = dt.DecisionTreeEvaluator()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda') if single else None
    torch.cuda.synchronize()

    def step(k=0):
        if single:
            ev.get_labels_forest(forest, depth[k:k + 1], labels[k:k + 1])
        else:
            ev.get_labels_forest(forest, depth, labels)

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    rdist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if single:
        elapsed = 0.0
        for k in range(args.steps):
            flush.zero_()
            e0.record()
            step(args.warmup + k)
            e1.record()
            torch.cuda.synchronize()
            elapsed += e0.elapsed_time(e1)
    else:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        elapsed = e0.elapsed_time(e1)
    rdist.barrier()
    elapsed_ms = rdist.max_over_ranks(elapsed)
    clocks = sampler.stop() if rank == 0 else None
    total_px = frames * H * W * (world if single else 1)
    value = total_px * args.steps / elapsed_ms / 1e3                 # Mpixels/s, whole job
    launches_per_step = (my_frames + 65534) // 65535

    # ---- parity spot check inside the bench: first frame of this rank vs the C oracle (outside the timed region) ----
    parity = None
    if rank == 0:
        from rdf_b200 import synth
        from oracle import c_oracle as co
        canon = forest.forest_cu.get()
        d0 = synth.depth_frames(kind, 1, H, W, seed=args.seed, first_frame=f0)
        assert np.array_equal(depth[0:1].get(), d0), 'device frame generator differs from synth.py'
        exp = np.full((1, H, W), 65535, np.uint16)
        co.eval_forest(canon, d0, exp)
        parity = bool(np.array_equal(labels[0:1].get(), exp))
        assert parity, 'label map of frame 0 differs from the oracle'

    # ---- e2e: host buffers through the public host API ----
    e2e = None
    if not args.no_extras:
        depth_host = pinned_like((my_frames, H, W), np.uint16)
        labels_host = pinned_like((my_frames, H, W), np.uint16)
        depth_host.view(torch.int16).copy_(depth.tensor[:my_frames].view(torch.int16))
        hb = HostBatchEvaluator(ev, forest, (H, W), chunk_frames=min(64, my_frames))
        hb.run(depth_host, labels_host)
        torch.cuda.synchronize()
        rdist.barrier()
        e0.record()
        for _ in range(args.e2e_steps):
            hb.run(depth_host, labels_host)
        e1.record()
        torch.cuda.synchronize()
        rdist.barrier()
        e2e_ms = rdist.max_over_ranks(e0.elapsed_time(e1))
        if rank == 0:
            got = labels_host[0:1].view(torch.int16).numpy().view(np.uint16)
            assert np.array_equal(got, labels[0:1].get()), 'host-API label map differs from the resident run'
        e2e = {'value': total_px * args.e2e_steps / e2e_ms / 1e3, 'unit': 'Mpixels/s',
               'h2d_bytes_per_step': int(rdist.sum_over_ranks(hb.bytes_h2d)), 'd2h_bytes_per_step': int(rdist.sum_over_ranks(hb.bytes_d2h)),
               'steps': args.e2e_steps, 'api': 'rdf_b200.pipeline.HostBatchEvaluator.run (pinned host frames -> pinned host label maps, 64-frame chunks on 3 streams)'}
        del depth_host, labels_host, hb

    if rank != 0:
        return

    peak, peak_src = measured_peaks()
    b_alg = b_alg_per_pixel(T, D, C)
    kernel_ms = elapsed_ms / args.steps / max(1, launches_per_step)   # max over ranks; one launch per step and rank
    px_per_launch = my_frames * H * W / max(1, launches_per_step)
    achieved = b_alg * px_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'eval_traffic.json')
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get('workload') == args.workload:
                traffic = tj['dram_bytes_per_pixel'] * px_per_launch
        except Exception:
            pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'kernel': f'rdf_eval_packed_kernel<{T},*,true,false>', 'algorithmic_bytes_per_pixel': b_alg, 'peak_source': peak_src,
                'compulsory_hbm_frac': 4.0 * px_per_launch / (kernel_ms * 1e-3) / 1e9 / peak,
                'node_steps_per_s': T * D * px_per_launch / (kernel_ms * 1e-3),
                'note': 'logical-traffic roofline (SURVEY 8d): exceeds 1.0 because the forest is cache-resident; the binding '
                        'resource is the L1 data pipe (93 % of peak under ncu, profiles/r01_ncu_eval_v3.md)'}

    line = {
        'metric': 'forest_eval_mpixels_per_s', 'value': value, 'unit': 'Mpixels/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_desc(args.workload), 'frames_total': frames, 'frames_per_rank': my_frames,
                   'parallelism': (f'replicas only: the same frame sequence on each of {world} rank(s)' if single else
                                   f'frames sharded over {world} rank(s), no collective'),
                   'l2': ('a different frame every step, L2 flushed (256 MB written) before each step, per-step CUDA events' if single else
                          'inputs larger than L2 (depth + labels = %.1f GB per rank per step)' % (2 * my_frames * H * W * 2 / 1e9))},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': args.steps * launches_per_step, 'roofline': roofline,
        'host_cpus_bound_per_rank': host_cpus_bound,
        'parity_checked': parity,
    }

    if not args.no_extras:
        # the reference's own kernel on the same GPU and inputs (sub-batch), with a bit-exact cross-check
        sub_frames = min(my_frames, 256)
        rg = ref_gpu_rate(forest.forest_cu, depth, sub_frames, H, W)
        if rg is not None:
            ref_labels = rg.pop('labels')
            same = bool(torch.equal(ref_labels.view(torch.int16), labels.tensor[:sub_frames].view(torch.int16)))
            rg['labels_bit_exact_vs_ours'] = same
            rg['speedup_ours_per_gpu'] = (value / world) / rg['value']
            line['ref_gpu'] = rg
            del ref_labels
        if world == 1:
            cb, _ = cpu_oracle_rate(T, D, C, W, H, kind, seed=args.seed)
            line['cpu_baseline'] = cb
        del depth, labels
        torch.cuda.empty_cache()
        try:
            line['latency'] = latency_cfg2()
        except Exception as e:                                        # never lose the headline number to an extra
            line['latency'] = {'error': repr(e)}
        try:
            line['train_cfg4'] = train_cfg4()
        except Exception as e:
            line['train_cfg4'] = {'error': repr(e)}
        try:
            line['hands_frame'] = hands_frame()
        except Exception as e:
            line['hands_frame'] = {'error': repr(e)}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
    try:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
    except Exception:
        pass
