#!/bin/bash
# Re-capture of the forest kernel after the staged-level change (same command as tools/ncu_r02_final.sh).
set -u
python bench.py --no-extras --steps 1 > gpurun_out/ncu_eval_ks5_plain.log 2>&1 || { echo plain run failed; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:rdf_eval_packed" -s 3 -c 1 -f -o gpurun_out/r02_ncu_eval_cfg3_full_ks5 \
    python bench.py --no-extras --steps 1 > gpurun_out/ncu_eval_ks5.log 2>&1
tail -2 gpurun_out/ncu_eval_ks5.log
