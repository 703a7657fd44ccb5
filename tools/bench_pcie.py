#!/usr/bin/env python
"""Host <-> device copy ceiling of a box, with NO kernels: every rank moves the e2e leg's traffic (pinned uint16 frames in,
pinned uint16 label maps out) (a) as the same chunk schedule on the same three streams HostBatchEvaluator uses and (b) as one
large H2D and one large D2H issued together.  Run it alone or under torchrun at N = 1, 2, 4, 8:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_pcie.py
Prints one JSON line on rank 0: aggregate GB/s (both directions summed over all ranks, max-over-ranks time)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-beats_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=4096, help='frames of the whole job (sharded over the ranks like bench.py)')
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--no-bind', action='store_true')
    args = ap.parse_args()
    from rdf_b200 import dist as rdist
    from rdf_b200 import decision_tree as dt
    from rdf_b200.pipeline import HostBatchEvaluator, pinned_like
    rank, world, local = rdist.init_from_env()
    bound = 0 if (world == 1 or args.no_bind) else rdist.bind_host_to_gpu(local)
    torch.cuda.set_device(local)
    H, W = 480, 848
    f0, f1 = rdist.shard_range(args.frames, rank, world)
    n = f1 - f0
    depth_host, labels_host = pinned_like((n, H, W), np.uint16), pinned_like((n, H, W), np.uint16)
    depth_host.view(torch.int16).zero_()
    hb = HostBatchEvaluator(None, None, (H, W), chunk_frames=min(128, n))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        rdist.barrier()
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        rdist.barrier()
        return rdist.max_over_ranks(e0.elapsed_time(e1)) / args.iters
    chunked_ms = timed(lambda: hb.run(depth_host, labels_host, copy_only=True))
    big_in = torch.empty((n, H, W), dtype=torch.int16, device='cuda')
    big_out = torch.zeros((n, H, W), dtype=torch.int16, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            big_in.copy_(depth_host.view(torch.int16), non_blocking=True)
        with torch.cuda.stream(s2):
            labels_host.view(torch.int16).copy_(big_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    big_ms = timed(both)
    h2d_ms = timed(lambda: big_in.copy_(depth_host.view(torch.int16), non_blocking=True))
    d2h_ms = timed(lambda: labels_host.view(torch.int16).copy_(big_out, non_blocking=True))
    total_bytes = rdist.sum_over_ranks(2 * n * H * W * 2)
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'frames': args.frames, 'bytes_each_way_total': int(total_bytes // 2),
                          'chunked_both_ways_gbs': round(total_bytes / chunked_ms / 1e6, 1),
                          'single_copy_both_ways_gbs': round(total_bytes / big_ms / 1e6, 1),
                          'h2d_alone_gbs': round(total_bytes / 2 / h2d_ms / 1e6, 1), 'd2h_alone_gbs': round(total_bytes / 2 / d2h_ms / 1e6, 1),
                          'mpix_s_if_kernels_were_free': round(args.frames * H * W / chunked_ms / 1e3, 1),
                          'host_cpus_bound_per_rank': bound, 'host_cores': len(os.sched_getaffinity(0))}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
