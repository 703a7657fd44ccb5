#!/usr/bin/env python
"""bench.py - forest-eval throughput (BASELINE.json metric) on N GPUs of one node, one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path on the host cores

Workload (default `cfg3`, BASELINE.json configs[2], the configuration the Mpixels/s metric is quoted on):
4096 synthetic 848x480 uint16 depth frames (dense-smooth), random-init 4-tree depth-20 forest, 4 classes,
sharded by frame over the ranks with no data-path collective (strong scaling: the 4096 frames are split).
A step = one pass of get_labels_forest over the rank's frames, inputs resident in HBM.  Objects on the (compact) line, in
this order so that a truncated tail keeps the headline facts:
  latency      BASELINE.json configs[1]: one 848x480 frame -> 2-layer stacked forest + 6-round mean shift, p50/p99 (N=1)
  configs      the other BASELINE configs, each with mpix_s + parity: cfg1, cfg3_noise, cfg3_survey_forest (the SURVEY 8d
               generator uploaded from NumPy instead of the on-device hash forest), cfg5, cfg5_noise (N=1)
  e2e          same metric through the host-buffer API (pinned host frames -> label maps in pinned host memory), H2D and D2H
               inside the timed region; copy_only_* = the same chunked copies with no kernel (the host-copy ceiling)
  roofline     algorithmic bytes (4 + T*(32*D + 4*C) B per pixel) / kernel time vs the measured HBM peak, plus
               `binding` = the physically bounded fraction (L1 data-pipe wavefronts, from the ncu capture in profiles/)
  ref_gpu      the reference's own kernels (compiled unchanged for sm_100a) on the same GPU, same inputs (sub-batch)
  cpu_baseline the C oracle on the host cores over a bounded sample; cpu_baseline_numpy = the NumPy oracle, one process and a
               pool of all host cores (N=1)
  train_cfg4   BASELINE.json configs[3]: the training split search at levels 0/4/8/12 (bucket + histogram + pick-best); under
               torchrun (N>1) images are sharded and both exchange modes are timed (NVLink reduce-scatter fused into the
               histogram kernel, NCCL allreduce), with a single-GPU run of the same level beside them and a parity check of a
               whole sharded training
  hands_frame  one whole product frame (N=1)
Verbose per-leg details go to stderr as `# details: {...}`.
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


def cpu_oracle_rate(T, D, C, W, H, kind, budget_s=12.0, max_frames=512, seed=1234):
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


def train_cfg4(level=8, iters=2, check=True):
    """BASELINE configs[3] on this rank's GPU: one level of the split search over 42 dense-smooth 848x480 frames (17.1 M labelled
    pixels, ~13 M of them still active at the deeper levels), 2000 features x 64 thresholds, C = 4, 2^level active nodes: bucket
    the pixels by node + histogram + pick-best.  A feature sub-block is checked against the C oracle on two frames."""
    import ctypes
    import hashlib
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
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def one_level(timed=False):
        best_gain.fill_(-1.0)
        _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes), N * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
        hist.zero_()
        if timed:
            ev[0].record()
        _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(ws), S, _capi.dptr(offsets),
                                                _capi.dptr(thresholds), F, NT, C, _capi.dptr(hist), st()))
        if timed:
            ev[1].record()
        _capi.check(lib.rdf_train_pick_best(S, _capi.dptr(slot), _capi.dptr(slot), _capi.dptr(parent), _capi.dptr(hist), S,
                                            _capi.dptr(offsets), _capi.dptr(thresholds), F, NT, C, level, D, _capi.dptr(tree),
                                            _capi.dptr(next_counts), _capi.dptr(best_gain), st()))
        if timed:
            ev[2].record()
    one_level()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        one_level()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    one_level(timed=True)                                        # per-kernel split of one more level
    torch.cuda.synchronize()
    hist_ms, pick_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    digest = hashlib.md5(tree.cpu().numpy().tobytes() + next_counts.cpu().numpy().tobytes()).hexdigest()
    active_px = int((nodes_np >= 0).sum())
    out = {'level': level, 'nodes': S, 'ms_per_level': ms, 'hist_ms': hist_ms, 'pick_best_ms': pick_ms,
           'pick_best_GBps': S * F * (NT + 1) * C * 4 / 1e6 / pick_ms, 'active_px': active_px,
           'g_feature_evals_per_s': active_px * F / ms / 1e6, 'algorithmic_GBps': (active_px * 8 + active_px * F * 8) / ms / 1e6,
           'records_md5': digest}
    if check:
        nf = 6                                                   # parity: 6 features x 64 thresholds on 2 frames vs the C oracle
        h2 = torch.zeros((S, nf, NT + 1, C), dtype=torch.int32, device='cuda')
        nodes2 = nodes[:2].contiguous()
        _capi.check(lib.rdf_train_bucket(_capi.dptr(nodes2), 2 * H * W, _capi.dptr(slot), S, _capi.dptr(ws), need.value, st()))
        _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth[:2]), _capi.dptr(labels[:2]), 2, W, H, _capi.dptr(ws), S, _capi.dptr(offsets[:nf]),
                                                _capi.dptr(thresholds[:nf]), nf, NT, C, _capi.dptr(h2), st()))
        torch.cuda.synchronize()
        exp = co.train_hist(depth_np[:2], labels_np[:2], nodes_np[:2], np.arange(S, dtype=np.int32), S, off_np[:nf], th_np[:nf], C)
        out['histogram_bit_exact_vs_c_oracle'] = bool(np.array_equal(h2.cpu().numpy().view(np.uint32), exp))
    del hist, ws, depth, labels, nodes
    torch.cuda.empty_cache()
    return out


def train_cfg4_sweep(levels=(0, 4, 8, 12)):
    """one full level sweep at 1 / 16 / 256 / 4096 active nodes."""
    res = {'workload': 'cfg4: 42 frames 848x480 (17.1 M labelled px), 2000 features x 64 thresholds, C=4; rdf_train_bucket + '
                       'rdf_train_hist_bucketed + rdf_train_pick_best, inputs resident', 'levels': {}}
    for lv in levels:
        r = train_cfg4(lv, check=(lv == 8))
        res['levels'][str(lv)] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k not in ('level',)}
    return res


def train_cfg4_multi_gpu(levels=(0, 8, 12)):
    """The path's one collective, under torchrun (every rank calls this): images sharded over the ranks, both exchange modes
    (tools/bench_train_mgpu.py), a whole sharded training checked against single-GPU training (tools/mgpu_check.py), and - on rank 0
    alone, afterwards - the same levels on one GPU for the efficiency."""
    import importlib.util
    import torch
    from rdf_b200 import dist as rdist

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'tools', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    rank, world, _ = rdist.env_rank_world()
    per_level = load('bench_train_mgpu').run(levels=list(levels), iters=3)
    parity = load('mgpu_check').run()
    out = None
    if rank == 0:
        out = {'n_gpus': world, 'exchange': 'images sharded: p2p = reduce-scatter fused into the histogram kernel over NVLink + all-gather of '
                                            'per-node winners; allreduce = NCCL sum-allreduce of the whole histogram.  dataset replicated: '
                                            'features = every rank builds the complete histograms of its own feature slice, only per-node '
                                            'winners are all-gathered',
               'levels': {}, 'sharded_training_matches_single_gpu': parity['sharded_training_matches_single_gpu'],
               'eval_shards_match_single_gpu': parity['eval_shards_match_single_gpu'],
               'exchange_modes_tested': parity['exchange_modes_tested']}
        for rec in per_level:
            one = train_cfg4(rec['cfg4_level'], check=False)
            t1 = one['ms_per_level']
            out['levels'][str(rec['cfg4_level'])] = {
                'nodes': rec['active_nodes'], 'ms_per_level_p2p': round(rec['ms_per_level_p2p'], 3),
                'ms_per_level_allreduce': round(rec['ms_per_level_allreduce'], 3),
                'ms_per_level_features': round(rec['ms_per_level_features'], 3), 'ms_per_level_1gpu': round(t1, 3),
                'efficiency_features': round(t1 / (world * rec['ms_per_level_features']), 3),
                'efficiency_p2p': round(t1 / (world * rec['ms_per_level_p2p']), 3),
                'efficiency_allreduce': round(t1 / (world * rec['ms_per_level_allreduce']), 3),
                'node_records_identical': bool(rec['node_records_identical'] and rec['records_md5'] == one['records_md5']),
                'records_md5': rec['records_md5']}
    rdist.barrier()
    return out


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


def make_forest(T, D, C, seed, kind='hash'):
    """kind 'hash': rdf_synth_forest on the device (csrc/rdf_synth.cu: every node from a hash of its index, each of ux, uy, vx, vy
    independently log2-uniform in magnitude - a 7.5 GiB forest is generated in place); 'survey': the SURVEY 8d generator
    (synth.random_forest: direction U(0, 2 pi), magnitude e^U(0,14)) drawn with NumPy and uploaded."""
    from rdf_b200 import _capi, synth
    from rdf_b200 import decision_tree as dt
    forest = dt.DecisionForest(T, D, C)
    if kind == 'survey':
        forest.forest_cu.set(synth.random_forest(T, D, C, seed=seed))
    else:
        _capi.check(_capi.load().rdf_synth_forest(_capi.dptr(forest.forest_cu), T, D, C, seed, _capi.stream_ptr()))
    return forest


FOREST_DESC = {'hash': 'hash_forest (on-device, per-component log2-uniform offsets)', 'survey': 'SURVEY 8d generator (NumPy, uploaded)'}


def time_eval(forest, kind, frames, W, H, steps, warmup, seed, first_frame=0, single=None):
    """Device time of `steps` passes of get_labels_forest, inputs resident.  single (one-frame workloads): a different frame every
    step and a 256 MB write to flush L2 before it, per-step events.  Returns (ms_per_step, depth, labels) - the device arrays of the
    LAST step stay available for parity checks."""
    import torch
    from rdf_b200 import _capi
    from rdf_b200 import decision_tree as dt
    lib = _capi.load()
    single = (frames == 1) if single is None else single
    ring = (steps + warmup) if single else frames
    depth = dt.cu_array.GPUArray((ring, H, W), dtype=np.uint16)
    _capi.check(lib.rdf_synth_depth(_capi.dptr(depth), KIND_ID[kind], ring, W, H, seed, first_frame, _capi.stream_ptr()))
    labels = dt.cu_array.GPUArray((ring, H, W), dtype=np.uint16).fill(65535)
    ev = dt.DecisionTreeEvaluator()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if single:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        for k in range(warmup):
            ev.get_labels_forest(forest, depth[k:k + 1], labels[k:k + 1])
        torch.cuda.synchronize()
        elapsed = 0.0
        for k in range(steps):
            flush.zero_()
            e0.record()
            ev.get_labels_forest(forest, depth[warmup + k:warmup + k + 1], labels[warmup + k:warmup + k + 1])
            e1.record()
            torch.cuda.synchronize()
            elapsed += e0.elapsed_time(e1)
        del flush
    else:
        for _ in range(warmup):
            ev.get_labels_forest(forest, depth, labels)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            ev.get_labels_forest(forest, depth, labels)
        e1.record()
        torch.cuda.synchronize()
        elapsed = e0.elapsed_time(e1)
    return elapsed / steps, depth, labels


def parity_vs_c_oracle(forest, kind, W, H, seed, frame_index, labels_frame):
    """One frame of what was timed against the C oracle (host threads)."""
    from rdf_b200 import synth
    from oracle import c_oracle as co
    d0 = synth.depth_frames(kind, 1, H, W, seed=seed, first_frame=frame_index)
    exp = np.full((1, H, W), 65535, np.uint16)
    co.eval_forest(forest.forest_cu.get(), d0, exp, nthreads=host_cores())
    return bool(np.array_equal(labels_frame.get(), exp))


def parity_vs_reference_kernel(forest, depth_frame, labels_frame):
    """One frame of what was timed against the reference's own kernel on the device (no host copy of the forest: cfg5's is 8 GB)."""
    import torch
    from oracle import ref_kernels as rk
    if not rk.available():
        return None
    ref = torch.full(tuple(labels_frame.shape), -1, dtype=torch.int16, device='cuda').view(torch.uint16)
    rk.eval_forest(forest.forest_cu.tensor, depth_frame.tensor, ref)
    torch.cuda.synchronize()
    return bool(torch.equal(ref.view(torch.int16), labels_frame.tensor.view(torch.int16)))


def other_configs(seed):
    """BASELINE.json's remaining configurations as compact {mpix_s, parity} entries (N = 1, rank 0)."""
    import torch
    out = {}

    def entry(name, T, D, C, W, H, kind, frames, steps, warmup, forest, checker, note=None):
        ms, depth, labels = time_eval(forest, kind, frames, W, H, steps, warmup, seed)
        last = (steps + warmup - 1) if frames == 1 else 0
        if checker == 'ref_kernel':
            par = parity_vs_reference_kernel(forest, depth[last:last + 1], labels[last:last + 1])
        else:
            par = parity_vs_c_oracle(forest, kind, W, H, seed, last, labels[last:last + 1])
        e = {'mpix_s': round(frames * H * W / ms / 1e3, 1), 'ms_per_step': round(ms, 4), 'parity': par,
             'parity_vs': 'reference kernel on the device, whole frame' if checker == 'ref_kernel' else 'C oracle, one frame',
             'frac_logical_roofline': round(b_alg_per_pixel(T, D, C) * frames * H * W / (ms * 1e-3) / 1e9 / measured_peaks()[0], 3)}
        if note:
            e['note'] = note
        out[name] = e
        del depth, labels
        torch.cuda.empty_cache()

    f1 = make_forest(3, 16, 4, seed)
    entry('cfg1', 3, 16, 4, 848, 480, 'dense-smooth', 1, 20, 5, f1, 'c_oracle', 'one frame per step, new frame + L2 flush each step')
    del f1
    f3 = make_forest(4, 20, 4, seed)
    entry('cfg3_noise', 4, 20, 4, 848, 480, 'dense-noise', 512, 3, 3, f3, 'c_oracle', '512 of the 4096 frames per step')
    del f3
    f3s = make_forest(4, 20, 4, seed, 'survey')
    entry('cfg3_survey_forest', 4, 20, 4, 848, 480, 'dense-smooth', 512, 3, 3, f3s, 'c_oracle',
          '512 frames per step, forest = ' + FOREST_DESC['survey'])
    del f3s
    torch.cuda.empty_cache()
    f5 = make_forest(8, 24, 4, seed)
    entry('cfg5', 8, 24, 4, 1280, 720, 'dense-smooth', 1, 10, 3, f5, 'ref_kernel', 'one frame per step, new frame + L2 flush each step')
    entry('cfg5_noise', 8, 24, 4, 1280, 720, 'dense-noise', 1, 10, 3, f5, 'ref_kernel', 'one frame per step, new frame + L2 flush each step')
    del f5
    torch.cuda.empty_cache()
    return out


def numpy_baseline_main(args):
    """--numpy-baseline (a subprocess of the default arm, started before it touches CUDA): the NumPy oracle (SURVEY 8d's CPU
    baseline) on a bounded sample of the workload - one process on a band of rows, then a multiprocessing pool of all host cores
    over the row bands of one whole frame.  Prints one JSON object."""
    import multiprocessing as mp
    from rdf_b200 import synth
    from oracle import numpy_oracle as no
    frames, W, H, T, D, C, kind = WORKLOADS[args.workload]
    forest = synth.hash_forest(T, D, C, seed=args.seed)
    NF = 4                                                   # frames of the pooled sample
    depth = synth.depth_frames(kind, NF, H, W, seed=args.seed)
    cores = host_cores()

    def band(n, y0, y1):
        filt = np.zeros((1, H, W), np.uint16)
        filt[0, y0:y1] = 1
        lab = np.full((1, H, W), 65535, np.uint16)
        no.eval_forest(forest, depth[n:n + 1], lab, 1, filt, 1)
        return lab[0, y0:y1]
    rows1 = H // 2
    t0 = time.perf_counter()
    band(0, 0, rows1)
    t1 = time.perf_counter() - t0
    single = rows1 * W / t1 / 1e6
    global _NB_BAND
    _NB_BAND = band
    bands = [(n, i * H // cores, (i + 1) * H // cores) for n in range(NF) for i in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(cores) as pool:
        parts = pool.starmap(_nb_call, bands)
    t2 = time.perf_counter() - t0
    print(json.dumps({'unit': 'Mpixels/s', 'single_process': single, 'pool': NF * H * W / t2 / 1e6, 'cores': cores, 'kind': 'port (NumPy oracle)',
                      'sample': f'single: {rows1} rows of one {W}x{H} frame ({t1:.1f} s); pool: {NF} whole frames in {cores} row bands each '
                                f'({t2:.1f} s); forest and frames of the workload', 'labelled_px_pool': int(sum((p != 65535).sum() for p in parts))}),
          flush=True)


_NB_BAND = None


def _nb_call(n, y0, y1):
    return _NB_BAND(n, y0, y1)


def numpy_baseline(args):
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--numpy-baseline', '--workload', args.workload, '--seed', str(args.seed)],
                             capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {'error': repr(e)[:200]}


def binding_from_profile(workload):
    """The physically bounded fraction of the dominant kernel, from the committed ncu capture of this workload's step."""
    path = os.path.join(ROOT, 'profiles', 'eval_traffic.json')
    try:
        tj = json.load(open(path))
        if tj.get('workload') == workload and 'binding' in tj:
            return tj['binding'], tj.get('dram_bytes_per_pixel')
        return None, (tj.get('dram_bytes_per_pixel') if tj.get('workload') == workload else None)
    except Exception:
        return None, None


def r(x, n=3):
    return round(float(x), n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--forest', default='hash', choices=['hash', 'survey'], help='forest generator of the headline workload')
    ap.add_argument('--frames', type=int, default=0, help='override the number of frames (debug)')
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--ref-frames', type=int, default=4, help='frames per step of the --impl reference CPU arm')
    ap.add_argument('--no-extras', action='store_true', help='skip every leg but the headline number')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--e2e-chunk', type=int, default=128, help='frames per chunk of the host-buffer pipeline')
    ap.add_argument('--e2e-no-ramp', action='store_true', help='plain chunks (no short first / last chunks): A/B aid')
    ap.add_argument('--latency-only', action='store_true', help='only the cfg2 frame-latency leg (profiling aid)')
    ap.add_argument('--latency-iters', type=int, default=1000)
    ap.add_argument('--numpy-baseline', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.numpy_baseline:
        numpy_baseline_main(args)
        return
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    rank_env = int(os.environ.get('RANK', '0'))
    world_env = int(os.environ.get('WORLD_SIZE', '1'))
    np_base = None
    if rank_env == 0 and world_env == 1 and not args.no_extras and not args.latency_only:
        np_base = numpy_baseline(args)              # before this process initialises CUDA (the pool forks)

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
    if os.environ.get('RDF_L2_FETCH'):               # experiment knob: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes)
        import ctypes
        rt = ctypes.CDLL('libcudart.so.12')
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ['RDF_L2_FETCH'])))
        val = ctypes.c_size_t()
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print(f'# cudaLimitMaxL2FetchGranularity -> rc {rc}, now {val.value}', file=sys.stderr)
    if args.latency_only:
        if rank == 0:
            print(json.dumps({'latency': latency_cfg2(iters=args.latency_iters, warm=min(100, args.latency_iters))}), flush=True)
        return
    _capi.load()

    frames, W, H, T, D, C, kind = WORKLOADS[args.workload]
    if args.frames:
        frames = args.frames
    f0, f1 = rdist.shard_range(frames, rank, world)
    single = frames == 1
    if single:                                     # single-frame workloads do not shard: every rank runs a replica
        f0, f1 = 0, 1
    my_frames = f1 - f0

    # ---- inputs, generated on the device (bit-exact twins of rdf_b200/synth.py); headline: inputs resident ----
    forest = make_forest(T, D, C, args.seed, args.forest)
    ev = dt.DecisionTreeEvaluator()
    torch.cuda.synchronize()
    rdist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, depth, labels = time_eval(forest, kind, my_frames, W, H, args.steps, args.warmup, args.seed, first_frame=f0, single=single)
    rdist.barrier()
    ms_step = rdist.max_over_ranks(ms_step)
    clocks = sampler.stop() if rank == 0 else None
    total_px = frames * H * W * (world if single else 1)
    value = total_px / ms_step / 1e3                                 # Mpixels/s, whole job
    launches_per_step = (my_frames + 65534) // 65535

    # ---- parity spot check inside the bench: one frame of this rank vs the C oracle (outside the timed region) ----
    parity = None
    if rank == 0:
        last = (args.steps + args.warmup - 1) if single else 0
        if T * (1 << D) * (7 + 2 * C) * 4 <= (2 << 30):
            parity = parity_vs_c_oracle(forest, kind, W, H, args.seed, f0 + last, labels[last:last + 1])
        else:                                                        # cfg5: the forest stays on the device
            parity = parity_vs_reference_kernel(forest, depth[last:last + 1], labels[last:last + 1])
        assert parity, 'label map differs from the checker'

    # ---- e2e: host buffers through the public host API, and the same copies with no kernel ----
    e2e = None
    if not args.no_extras:
        depth_host = pinned_like((my_frames, H, W), np.uint16)
        labels_host = pinned_like((my_frames, H, W), np.uint16)
        depth_host.view(torch.int16).copy_(depth.tensor[:my_frames].view(torch.int16))
        hb = HostBatchEvaluator(ev, forest, (H, W), chunk_frames=min(args.e2e_chunk, my_frames), ramp=0 if args.e2e_no_ramp else 8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(n, **kw):
            hb.run(depth_host, labels_host, **kw)
            torch.cuda.synchronize()
            rdist.barrier()
            e0.record()
            for _ in range(n):
                hb.run(depth_host, labels_host, **kw)
            e1.record()
            torch.cuda.synchronize()
            rdist.barrier()
            return rdist.max_over_ranks(e0.elapsed_time(e1)) / n
        copy_ms = timed(2, copy_only=True)
        e2e_ms = timed(args.e2e_steps)
        if rank == 0 and not single:
            got = labels_host[0:1].view(torch.int16).numpy().view(np.uint16)
            assert np.array_equal(got, labels[0:1].get()), 'host-API label map differs from the resident run'
        h2d, d2h = int(rdist.sum_over_ranks(hb.bytes_h2d)), int(rdist.sum_over_ranks(hb.bytes_d2h))
        e2e = {'value': r(total_px / e2e_ms / 1e3, 1), 'unit': 'Mpixels/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'steps': args.e2e_steps, 'ms_per_step': r(e2e_ms),
               'copy_only_mpix_s': r(total_px / copy_ms / 1e3, 1), 'host_copy_ceiling_gbs': r((h2d + d2h) / copy_ms / 1e6, 1),
               'achieved_copy_gbs': r((h2d + d2h) / e2e_ms / 1e6, 1),
               'chunks': len(hb.chunk_sizes(my_frames)),
               'api': 'rdf_b200.pipeline.HostBatchEvaluator.run (pinned host frames -> pinned host label maps, %d-frame chunks%s, 3 device '
                      'buffers on 3 streams); copy_only_* / host_copy_ceiling_gbs = the same chunked H2D + D2H with no kernel'
                      % (hb.chunk, ' reached by doubling from %d frames at both ends' % hb.ramp if hb.ramp else '')}
        del depth_host, labels_host, hb

    # ---- the reference's own kernel on the same GPU and inputs (sub-batch), with a bit-exact cross-check ----
    rg = None
    if rank == 0 and not args.no_extras and not single:
        sub_frames = min(my_frames, 256)
        rg = ref_gpu_rate(forest.forest_cu, depth, sub_frames, H, W)
        if rg is not None:
            ref_labels = rg.pop('labels')
            rg['labels_bit_exact_vs_ours'] = bool(torch.equal(ref_labels.view(torch.int16), labels.tensor[:sub_frames].view(torch.int16)))
            rg['speedup_ours_per_gpu'] = r((value / world) / rg['value'], 2)
            rg['value'], rg['ms_per_pass'] = r(rg['value'], 1), r(rg['ms_per_pass'])
            rg.pop('what')
            del ref_labels
    del depth, labels, forest
    torch.cuda.empty_cache()

    # ---- the path's one collective: split search sharded over the ranks (all ranks take part) ----
    train_mgpu = None
    if world > 1 and not args.no_extras:
        try:
            train_mgpu = train_cfg4_multi_gpu()
        except Exception as e:                                        # never lose the headline number to an extra
            train_mgpu = {'error': repr(e)[:300]}
    if rank != 0:
        return

    peak, peak_src = measured_peaks()
    b_alg = b_alg_per_pixel(T, D, C)
    kernel_ms = ms_step / max(1, launches_per_step)                  # max over ranks; one launch per step and rank
    px_per_launch = my_frames * H * W / max(1, launches_per_step)
    achieved = b_alg * px_per_launch / (kernel_ms * 1e-3) / 1e9
    binding, dram_bpp = binding_from_profile(args.workload)
    roofline = {'bound': 'hbm', 'achieved': r(achieved, 1), 'peak': peak, 'unit': 'GB/s', 'frac': r(achieved / peak, 4),
                'traffic': (dram_bpp * px_per_launch if dram_bpp else None),
                'kernel': f'rdf_eval_packed_kernel<{T},16,true,*>(rdf_eval_launch<{T}>)', 'algorithmic_bytes_per_pixel': b_alg, 'peak_source': peak_src,
                'compulsory_hbm_frac': r(4.0 * px_per_launch / (kernel_ms * 1e-3) / 1e9 / peak, 5),
                'node_steps_per_s': T * D * px_per_launch / (kernel_ms * 1e-3), 'binding': binding,
                'note': 'logical-traffic roofline (SURVEY 8d): exceeds 1.0 because the forest is cache-resident; `binding` is the '
                        'physically bounded fraction (ncu, profiles/)'}

    line = {
        'metric': 'forest_eval_mpixels_per_s', 'value': r(value, 1), 'unit': 'Mpixels/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': r(ms_step), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    }
    details = {}
    if not args.no_extras and world == 1:
        try:
            lat = latency_cfg2()
            details['latency'] = lat
            line['latency'] = {'p50_us': r(lat['p50_us'], 1), 'p99_us': r(lat['p99_us'], 1), 'device_p50_us': r(lat['device_p50_us'], 1),
                               'resident_p50_us': r(lat['resident_frame']['p50_us'], 1), 'parity': lat['parity'],
                               'what': 'cfg2 e2e: host frame -> upload -> 2-layer forest -> mean shift -> centroids in pinned host memory'}
        except Exception as e:
            line['latency'] = {'error': repr(e)[:200]}
        try:
            line['configs'] = other_configs(args.seed)
        except Exception as e:
            line['configs'] = {'error': repr(e)[:300]}
    line.update({
        'config': {'workload': workload_desc(args.workload), 'forest': FOREST_DESC[args.forest], 'frames_total': frames,
                   'frames_per_rank': my_frames,
                   'parallelism': (f'replicas only: the same frame sequence on each of {world} rank(s)' if single else
                                   f'frames sharded over {world} rank(s), no collective'),
                   'l2': ('a different frame every step, L2 flushed (256 MB written) before each step, per-step CUDA events' if single else
                          'inputs larger than L2 (depth + labels = %.1f GB per rank per step)' % (2 * my_frames * H * W * 2 / 1e9))},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': args.steps * launches_per_step, 'roofline': roofline,
        'host_cpus_bound_per_rank': host_cpus_bound, 'parity_checked': parity,
    })
    if rg is not None:
        line['ref_gpu'] = rg
    if train_mgpu is not None:
        line['train_cfg4'] = train_mgpu
    if not args.no_extras and world == 1:
        cb, _ = cpu_oracle_rate(T, D, C, W, H, kind, seed=args.seed)
        cb['value'] = r(cb['value'], 3)
        line['cpu_baseline'] = cb
        line['cpu_baseline_numpy'] = np_base
        try:
            line['train_cfg4'] = train_cfg4_sweep()
        except Exception as e:
            line['train_cfg4'] = {'error': repr(e)[:300]}
        try:
            hf = hands_frame()
            details['hands_frame'] = hf
            line['hands_frame'] = compact_hands(hf)
        except Exception as e:
            line['hands_frame'] = {'error': repr(e)[:200]}
    print(json.dumps(line), flush=True)
    if details:
        print('# details: ' + json.dumps(details), file=sys.stderr, flush=True)


def compact_hands(hf):
    """Numbers and parity flags of tools/bench_hands_frame.run's result, without its prose."""
    def walk(o):
        if isinstance(o, dict):
            return {k: walk(v) for k, v in o.items() if not (isinstance(v, str) and len(v) > 60)}
        if isinstance(o, float):
            return round(o, 2)
        return o
    return walk(hf)


if __name__ == '__main__':
    main()
    try:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
    except Exception:
        pass
