#!/bin/bash
# Evidence pass for the eval kernel whose upper levels travel as kernel parameters: GPU tests, default bench + reference arm,
# ncu --set full of the kernel on the full 4096-frame step.  Outputs under gpurun_out/final3/.
set -u
O=gpurun_out/final3; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
T0=$SECONDS; python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log; echo "pytest wall $((SECONDS-T0)) s"
T0=$SECONDS; python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $? wall $((SECONDS-T0)) s"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc $?"
python bench.py --no-extras --steps 1 > $O/ncu_eval_plain.log 2>&1 || { echo plain run failed; exit 1; }
T0=$SECONDS; ncu --set full --clock-control none --import-source on -k "regex:rdf_eval_packed" -s 3 -c 1 -f -o $O/r02_ncu_eval_cfg3_const_top \
    python bench.py --no-extras --steps 1 > $O/ncu_eval.log 2>&1; echo "ncu rc $? wall $((SECONDS-T0)) s"
head -c 400 $O/bench_default.json; echo; ls -la $O
