#!/bin/bash
# Round-end evidence pass on one B200: smoke, GPU tests, default bench (+ reference arm), launch list of the same bench command,
# whole-tree training runs, split-search phases.  Outputs under gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc $?"
python tools/bench_train_phases.py --levels 12 > $O/train_phases.log 2>&1; tail -3 $O/train_phases.log
python tools/bench_train_tree.py --frames 42 --depth 16 --proposals 2000 --blocks 1 --thresholds 64 > $O/train_tree_cfg4.json 2>&1
python tools/bench_train_tree.py --frames 8 --depth 12 --proposals 64 --blocks 4 > $O/train_tree_small.json 2>&1
tail -1 $O/train_tree_cfg4.json; tail -1 $O/train_tree_small.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default.csv \
    python bench.py --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1; echo "ncu rc $?"
