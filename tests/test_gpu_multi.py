"""Multi-GPU parity as a test (runs when the box shows at least two GPUs; the driver's 1-GPU test box skips it and the same checks
run inside `bench.py --gpus N` for N > 1, key train_cfg4): torchrun with two ranks over NCCL -
  * frame-sharded forest eval equals the single-GPU label maps;
  * image-sharded training through DecisionTreeTrainer, with the reduce-scatter fused into the histogram kernel over peer-mapped
    buffers (symmetric-memory barriers, all-gather of per-node winners) AND with the NCCL allreduce, builds the bit-identical tree a
    single GPU trains on the whole dataset (tools/mgpu_check.py);
  * the cfg4 level step gives identical node records in both exchange modes (tools/bench_train_mgpu.py, small size)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, *extra):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (the same checks run in bench.py --gpus N, key train_cfg4)')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', str(_port()), os.path.join(ROOT, 'tools', script), *extra],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{')]


def test_sharded_eval_and_training_match_single_gpu():
    res = _torchrun('mgpu_check.py')[-1]
    assert res['world'] == 2 and res['eval_shards_match_single_gpu'] and res['sharded_training_matches_single_gpu']
    assert any('p2p' in v for v in res['exchange_modes_tested'].values()), res['exchange_modes_tested']


def test_level_step_records_identical_in_both_exchange_modes():
    recs = _torchrun('bench_train_mgpu.py', '--frames', '4', '--features', '96', '--levels', '0,5', '--iters', '1')
    assert len(recs) == 2 and all(r['node_records_identical'] for r in recs)
