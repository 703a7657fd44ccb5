"""Host logic of HostBatchEvaluator that needs no GPU: the chunk schedule (ramp up, bulk, ramp down) covers every frame once, in
order, within the device buffers' capacity."""
import pytest

from rdf_b200.pipeline import HostBatchEvaluator


def _schedule(chunk, ramp, N):
    hb = HostBatchEvaluator.__new__(HostBatchEvaluator)          # no device buffers: chunk_sizes only reads these two
    hb.chunk, hb.ramp = chunk, ramp
    return hb.chunk_sizes(N)


@pytest.mark.parametrize('chunk,ramp', [(128, 8), (64, 8), (16, 2), (8, 1), (32, 0), (4, 8), (1, 8)])
def test_schedule_covers_every_frame_once(chunk, ramp):
    for N in list(range(0, 80)) + [100, 183, 512, 1000, 4096, 4100]:
        sizes = _schedule(chunk, ramp, N)
        assert sum(sizes) == N
        assert all(1 <= s <= chunk for s in sizes), (N, sizes)


def test_schedule_shape_of_the_bench_workload():
    sizes = _schedule(128, 8, 4096)
    assert sizes[:5] == [8, 16, 32, 64, 128] and sizes[-4:] == [64, 32, 16, 8]      # fills and drains with 8-frame copies
    assert sizes.count(128) == (4096 - 2 * 120) // 128 and len(sizes) <= 40
    assert _schedule(128, 0, 300) == [128, 128, 44]                                  # ramp off: plain chunks
    assert _schedule(128, 8, 512 // 1) [:4] == [8, 16, 32, 64]                       # a rank's share at 8 GPUs still ramps
    assert _schedule(128, 8, 5) == [5] and _schedule(128, 8, 0) == []
