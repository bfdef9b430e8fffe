"""Multi-GPU plumbing: one process per GPU, envs sharded, no data-path collective.

The only exchange of the path is the per-rollout statistics reduction: one
all-reduce of MRS_STATS_SLOTS int64 counters -- NCCL over NVLink/NVSwitch on GPUs, gloo in the
CPU tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """RANK / WORLD_SIZE / MASTER_* come from torchrun.  Returns (rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            local = int(os.environ.get('LOCAL_RANK', rank))
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def max_over_ranks(value: float, device) -> float:
    """Timing reduction of the bench contract: max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa(gpu_index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to the GPU, so that pinned host
    buffers are first-touched on the GPU's NUMA node (matters for the host-buffer paths when 8 ranks
    share one box).  Best effort: returns False when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return True
    except Exception:
        pass
    return False
