#!/bin/bash
# Re-capture of the two kernels that changed after tools/ncu_r02.sh ran (forest eval instances, histogram kernel): same commands.
set -u
N="ncu --set full --clock-control none --import-source on"
run() {
  local name="$1" re="$2" skip="$3" cnt="$4"; shift 4
  echo "== $name"
  "$@" > gpurun_out/ncu_${name}_plain.log 2>&1 || { echo "plain run failed: $name"; tail -5 gpurun_out/ncu_${name}_plain.log; return; }
  $N -k "regex:$re" -s "$skip" -c "$cnt" -f -o gpurun_out/r02_ncu_$name "$@" > gpurun_out/ncu_${name}.log 2>&1
  tail -2 gpurun_out/ncu_${name}.log
}
run eval_cfg3_full_final 'rdf_eval_packed' 3 1 python bench.py --no-extras --steps 1
run train_l12_final 'hist_bucketed|pick_best' 0 2 python tools/bench_train_phases.py --levels 12 --features 500
