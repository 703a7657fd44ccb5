#!/bin/bash
# Round-2 ncu captures (one per hot kernel, --set full, -lineinfo sources): run under gpurun, reports land in gpurun_out/.
# Each command first runs WITHOUT ncu and must exit 0 (B200_PROFILING.md).
set -u
N="ncu --set full --clock-control none --import-source on"
run() {   # name, kernel regex, skip, count, command...
  local name="$1" re="$2" skip="$3" cnt="$4"; shift 4
  echo "== $name"
  "$@" > gpurun_out/ncu_${name}_plain.log 2>&1 || { echo "plain run failed: $name"; tail -5 gpurun_out/ncu_${name}_plain.log; return; }
  $N -k "regex:$re" -s "$skip" -c "$cnt" -f -o gpurun_out/r02_ncu_$name "$@" > gpurun_out/ncu_${name}.log 2>&1
  tail -2 gpurun_out/ncu_${name}.log
}
run eval_cfg3_full 'rdf_eval_packed' 3 1 python bench.py --no-extras --steps 1
run eval_cfg5_noise 'rdf_eval_packed' 3 1 python bench.py --no-extras --steps 2 --workload cfg5-noise
run layered_meanshift 'rdf_layered_walks_kernel|rdf_mean_shift_v3_kernel' 0 2 python bench.py --latency-only --latency-iters 20
run hands_frame 'rdf_group_hands|rdf_condition_kernel|rdf_stencil_hands' 0 3 python tools/bench_hands_frame.py --iters 20 --no-ref
run train_l12 'hist_bucketed|pick_best' 0 2 python tools/bench_train_phases.py --levels 12 --features 500
ls -la gpurun_out/r02_ncu_*.ncu-rep
