#!/usr/bin/env python
"""Run a few eager (no CUDA graph) product frames so that ncu sees every kernel of HandsFramePipeline as its own launch:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hands.csv python tools/profile_hands_frame.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bench_hands_frame as b   # noqa: E402

pipe, scene, forests, cfg, variances = b.build(use_graph=False)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    pipe.run(scene['depth_raw'])
print('ok')
