"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic:
  * forest eval shards by frame with no collective: the ranks' label maps concatenate to the single-rank result;
  * the training split search shards images and sum-allreduces the integer histograms: the reduced histogram is identical to
    the single-rank one (exact integers, order-independent), so every rank picks the same split.
The per-rank compute stands in with the C oracle here (no GPU in this container); on a GPU box the same rdf_b200.dist helpers
wrap the CUDA kernels (bench.py, DecisionTreeTrainer)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, '3d-beats_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from rdf_b200 import dist as rdist, synth
    from oracle import c_oracle as co
    r, w, _ = rdist.init_from_env(backend='gloo')
    assert (r, w) == (rank, world) and dist.get_world_size() == world

    # ---- frame-sharded eval, no collective ----
    N, H, W = 5, 40, 56                                    # 5 frames over 2 ranks: uneven shards (3 + 2)
    forest = synth.random_forest(3, 7, 4, seed=3, ragged=True)
    f0, f1 = rdist.shard_range(N, rank, world)
    depth = synth.depth_frames('dense-smooth', f1 - f0, H, W, seed=8, first_frame=f0)
    labels = np.full((f1 - f0, H, W), 65535, np.uint16)
    co.eval_forest(forest, depth, labels, nthreads=1)
    np.save(os.path.join(out_dir, f'labels_{rank}.npy'), labels)

    # ---- image-sharded split histograms + sum-allreduce ----
    Nt, C, F, NT, level = 4, 4, 6, 8, 2
    i0, i1 = rdist.shard_range(Nt, rank, world)
    tdepth = synth.depth_frames('dense-smooth', i1 - i0, H, W, seed=5, first_frame=i0)
    tlabels = synth.train_labels(i1 - i0, H, W, first_frame=i0)
    nodes_all = synth.random_node_assignment(synth.train_labels(Nt, H, W), level, seed=2)
    nodes = nodes_all[i0:i1]
    off, th = synth.random_proposals(F, NT, seed=4)
    slot = np.arange(1 << level, dtype=np.int32)
    hist = co.train_hist(tdepth, tlabels, nodes, slot, 1 << level, off, th, C, nthreads=1)
    t = torch.from_numpy(hist.view(np.int32).copy())
    dist.all_reduce(t)                                     # the path's only exchange step
    np.save(os.path.join(out_dir, f'hist_{rank}.npy'), t.numpy().view(np.uint32))

    # ---- timing helpers ----
    assert rdist.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    assert rdist.sum_over_ranks(1.0) == float(world)
    rdist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from rdf_b200 import synth
    from oracle import c_oracle as co
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

    N, H, W = 5, 40, 56
    forest = synth.random_forest(3, 7, 4, seed=3, ragged=True)
    depth = synth.depth_frames('dense-smooth', N, H, W, seed=8)
    want = np.full((N, H, W), 65535, np.uint16)
    co.eval_forest(forest, depth, want)
    got = np.concatenate([np.load(tmp_path / f'labels_{r}.npy') for r in range(world)])
    assert np.array_equal(got, want)

    Nt, C, F, NT, level = 4, 4, 6, 8, 2
    tdepth = synth.depth_frames('dense-smooth', Nt, H, W, seed=5)
    tlabels = synth.train_labels(Nt, H, W)
    nodes = synth.random_node_assignment(tlabels, level, seed=2)
    off, th = synth.random_proposals(F, NT, seed=4)
    full = co.train_hist(tdepth, tlabels, nodes, np.arange(1 << level, dtype=np.int32), 1 << level, off, th, C)
    h0, h1 = np.load(tmp_path / 'hist_0.npy'), np.load(tmp_path / 'hist_1.npy')
    assert np.array_equal(h0, h1), 'ranks disagree after the allreduce'
    assert np.array_equal(h0, full), 'allreduced shard histograms differ from the single-rank histogram'


def _proposal_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, '3d-beats_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    from rdf_b200 import dist as rdist
    from rdf_b200 import decision_tree as dt
    rdist.init_from_env(backend='gloo')
    np.random.seed(100 + rank)                              # the ranks are NOT seeded identically
    out = {}
    for name, nt in (('nt1', 1), ('nt4', 4)):
        tr = dt.DecisionTreeTrainer(None, 6, thresholds_per_feature=nt)
        blocks = [tr._next_proposals(0, b) for b in range(3)]
        out[name + '.off'] = np.stack([b[0] for b in blocks])
        out[name + '.thr'] = np.stack([b[1] for b in blocks])
    # a proposal_fn that differs per rank is overridden by rank 0's as well; unsorted rows come back sorted
    tr = dt.DecisionTreeTrainer(None, 2, thresholds_per_feature=3,
                                proposal_fn=lambda lvl, blk: (np.full((2, 4), rank + 1.0), np.array([[3., 1., 2.], [9., -7., 8.]]) + rank))
    out['fn.off'], out['fn.thr'] = tr._next_proposals(0, 0)
    bad = dt.DecisionTreeTrainer(None, 1, thresholds_per_feature=2, proposal_fn=lambda lvl, blk: (np.zeros((1, 4)), np.array([[1., np.nan]])))
    try:
        bad._next_proposals(0, 0)
        out['nan_rejected'] = np.array(False)
    except ValueError:
        out['nan_rejected'] = np.array(True)
    np.savez(os.path.join(out_dir, f'prop_{rank}.npz'), **out)
    rdist.barrier()
    dist.destroy_process_group()


def test_default_proposal_stream_is_rank0s_on_every_rank(tmp_path):
    """ranks draw proposals from their own np.random state; rank 0's block must be what every rank scores."""
    import torch.multiprocessing as mp
    from rdf_b200 import decision_tree as dt
    world, port = 2, _free_port()
    mp.spawn(_proposal_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    z0, z1 = np.load(tmp_path / 'prop_0.npz'), np.load(tmp_path / 'prop_1.npz')
    for k in z0.files:
        assert np.array_equal(z0[k], z1[k]), f'ranks disagree on {k}'
    assert bool(z0['nan_rejected'])
    # ... and it IS rank 0's stream: replay seed 100 single-process
    np.random.seed(100)
    for name, nt in (('nt1', 1), ('nt4', 4)):
        tr = dt.DecisionTreeTrainer(None, 6, thresholds_per_feature=nt, process_group=False)
        blocks = [tr._next_proposals(0, b) for b in range(3)]
        assert np.array_equal(z0[name + '.off'], np.stack([b[0] for b in blocks]))
        assert np.array_equal(z0[name + '.thr'], np.stack([b[1] for b in blocks]))
        assert (np.diff(z0[name + '.thr'], axis=-1) >= 0).all()
    assert np.array_equal(z0['fn.off'], np.full((2, 4), 1.0, np.float32))
    assert np.array_equal(z0['fn.thr'], np.array([[1., 2., 3.], [-7., 8., 9.]], np.float32))


@pytest.mark.parametrize('total,world', [(4096, 1), (4096, 8), (5, 2), (7, 8), (0, 4), (42, 4)])
def test_shard_range_partitions(total, world):
    from rdf_b200.dist import shard_range
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` needs no GPU and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '3',
                          '--workload', 'cfg1', '--ref-frames', '1'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'Mpixels/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['e2e']['h2d_bytes_per_step'] == 0


def test_bind_host_to_gpu_is_harmless_without_nvml_devices():
    """Without a GPU (or NVML) the NUMA binding helper changes nothing and reports 0; it never leaves an empty CPU set."""
    import os
    from rdf_b200 import dist as rdist
    before = os.sched_getaffinity(0)
    n = rdist.bind_host_to_gpu(0)
    after = os.sched_getaffinity(0)
    assert isinstance(n, int) and n >= 0 and len(after) > 0
    if n == 0:
        assert after == before
    os.sched_setaffinity(0, before)
