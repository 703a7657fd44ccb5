#!/bin/bash
# All BASELINE.json configs in one go (one GPU): run after every kernel change - a tweak for one tree count or level shape can
# silently cost another (the 6-8 tree instantiations once spilled after a launch-bound change made for 4 trees).
#   bash tools/bench_all.sh > gpurun_out/bench_all.log
cd "$(dirname "$0")/.."
one() { python bench.py --no-extras "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-12s %9.1f Mpx/s  %9.4f ms/step  %6.0f G node-steps/s  parity=%s' % (sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['node_steps_per_s'] / 1e9, d['parity_checked']))" "$2"; }
one --workload cfg1 --steps 30
one --workload cfg3 --frames 1024
one --workload cfg3-noise --frames 512
one --workload cfg5 --steps 10
one --workload cfg5-noise --steps 10
python bench.py --latency-only --latency-iters 500 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])['latency']
print('cfg2 e2e p50 %.1f us p99 %.1f | copy-engine p50 %.1f | resident p50 %.1f | stages %s | parity %s' % (d['p50_us'], d['p99_us'], d['e2e_host_frame_copy_engine']['p50_us'], d['resident_frame']['p50_us'], {k: round(v, 1) for k, v in d['stages'].items() if k != 'note'}, d['parity']))"
python tools/bench_train.py --levels 0,8,12 --check 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    if 'cfg4_level' in d: print('cfg4 level %2d  %7.1f ms/level  %6.1f G feature evals/s' % (d['cfg4_level'], d['ms_per_level'], d['g_feature_evals_per_s']))
    else: print(d)"
python tools/bench_train_tree.py 2>/dev/null | tail -1
python tools/bench_grouping.py 2>/dev/null | tail -1
python tools/bench_hands_frame.py --iters 500 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])['hands_frame']
r = d.get('reference_kernels_same_gpu', {})
print('product frame e2e p50 %.1f us p99 %.1f (device p50 %.1f, %d launches) | reference kernels + host sequence p50 %s us | parity %s' % (
    d['e2e_host_frame']['p50_us'], d['e2e_host_frame']['p99_us'], d['e2e_host_frame']['device_p50_us'], d['kernels_per_frame'],
    r.get('p50_us'), d['parity']))"
