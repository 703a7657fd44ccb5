#!/bin/bash
# Second-session evidence pass on one B200: smoke, GPU tests, default bench (wall-clock timed) and the reference arm.
set -u
O=gpurun_out/final2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
T0=$SECONDS; python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log; echo "pytest wall $((SECONDS-T0)) s"
T0=$SECONDS; python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $? wall $((SECONDS-T0)) s"
T0=$SECONDS; python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc $? wall $((SECONDS-T0)) s"
head -c 600 $O/bench_default.json; echo; cat $O/bench_reference.json | head -c 600
