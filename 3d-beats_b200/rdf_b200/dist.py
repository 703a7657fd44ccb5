"""Multi-GPU plumbing for the RDF path: one process per GPU under torchrun.

Forest evaluation shards by frame with NO data-path collective (probes never cross an image, src/cuda/cu_utils.hpp:79-86);
training shards whole images across ranks and sum-allreduces the integer split histograms (exact, order-independent), which
DecisionTreeTrainer does when a process group with world_size > 1 is initialised.
"""
import os

import torch


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns (rank, world_size, local_rank).  No-op for world_size 1."""
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def bind_host_to_gpu(local_rank):
    """Pin the calling process to the CPU cores next to its GPU (NVML's ideal CPU affinity), so that the pinned host buffers it
    allocates afterwards are first-touched on the GPU's own NUMA node and its PCIe copies do not cross the socket link.  With eight
    ranks moving 6.7 GB per pass through host memory this is what bounds the host-buffer path.  Returns the number of CPUs the
    process is bound to, or 0 when NVML is not available (nothing changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = None
        try:                                                         # NVML and CUDA may order the devices differently
            pr = torch.cuda.get_device_properties(int(local_rank))
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(('%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def shard_range(total, rank, world):
    """Contiguous block partition of `total` units (frames / images) over `world` ranks; sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (timing is reported as the slowest rank's device time)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ('cuda' if dist.get_backend() == 'nccl' else 'cpu'))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ('cuda' if dist.get_backend() == 'nccl' else 'cpu'))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def broadcast_proposals(offsets, thresholds, group=None, src=0):
    """Rank `src`'s proposal block (offsets float32[P,4], thresholds float32[P,NT], NumPy) on every rank.  The trainer draws
    proposals from the process-local np.random state; without this, ranks that were not seeded identically would sum histograms
    of DIFFERENT features.  One small broadcast per proposal block (P x (4 + NT) floats), on the backend's own device."""
    import numpy as np
    import torch.distributed as dist
    P = offsets.shape[0]
    packed = np.concatenate([np.ascontiguousarray(offsets, dtype=np.float32).reshape(P, 4),
                             np.ascontiguousarray(thresholds, dtype=np.float32).reshape(P, -1)], axis=1)
    t = torch.from_numpy(packed)
    on_gpu = dist.get_backend(group) == 'nccl'
    if on_gpu:
        t = t.cuda()
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    packed = t.cpu().numpy()
    return np.ascontiguousarray(packed[:, :4]), np.ascontiguousarray(packed[:, 4:])


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
