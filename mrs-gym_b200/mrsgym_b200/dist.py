"""Multi-GPU plumbing: one process per GPU, envs sharded, no data-path collective.

The only exchange of the path is the per-rollout statistics reduction: one
all-reduce of MRS_STATS_SLOTS int64 counters.  On GPUs it is the library's own peer-memory kernel
(mrs_stats_allreduce over a PeerComm: NVLink stores into the peers' mailboxes, no host involvement,
graph-capturable); torch.distributed (NCCL / gloo) is the plumbing that exchanges the mailbox
handles, aligns host-side barriers and takes the max of the timings -- and the fallback reduction of
the CPU (gloo) tests."""
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


class PeerComm:
    """The library's peer-memory communicator over the GPUs of one node (include/mrs_b200.h, mrs_comm_*).
    One process per GPU: the 64-byte mailbox handles travel through torch.distributed once, at construction."""

    def __init__(self, rank=None, world=None, group=None):
        import ctypes as C
        from . import _abi
        self._abi, self._C = _abi, C
        self.lib = _abi.lib()
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        if self.world > _abi.COMM_MAX_WORLD:
            raise _abi.MrsError('PeerComm spans at most %d GPUs of one node' % _abi.COMM_MAX_WORLD)
        h = C.c_void_p()
        _abi.check(self.lib.mrs_comm_create(self.rank, self.world, C.byref(h)), 'mrs_comm_create')
        self.handle = h
        if self.world > 1:
            buf = C.create_string_buffer(_abi.COMM_HANDLE_BYTES)
            _abi.check(self.lib.mrs_comm_handle(self.handle, buf), 'mrs_comm_handle')
            mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8)
            on_gpu = dist.get_backend(group) == 'nccl'
            if on_gpu:
                mine = mine.cuda()
            allh = torch.empty(self.world * _abi.COMM_HANDLE_BYTES, dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            raw = bytes(allh.cpu().numpy().tobytes())
            _abi.check(self.lib.mrs_comm_connect(self.handle, raw), 'mrs_comm_connect')
            dist.barrier(group)        # nobody uses a mailbox before every rank has mapped all of them

    @classmethod
    def local_group(cls, devices):
        """One process driving several GPUs: a communicator per device, connected through plain peer pointers
        (mrs_comm_connect_ptrs).  Returns the list of PeerComm objects in `devices` order."""
        import ctypes as C
        from . import _abi
        lib = _abi.lib()
        world = len(devices)
        comms = []
        for r, d in enumerate(devices):
            with torch.cuda.device(d):
                c = cls.__new__(cls)
                c._abi, c._C, c.lib, c.rank, c.world = _abi, C, lib, r, world
                h = C.c_void_p()
                _abi.check(lib.mrs_comm_create(r, world, C.byref(h)), 'mrs_comm_create')
                c.handle = h
                comms.append(c)
        boxes = (C.c_void_p * world)(*[lib.mrs_comm_mailbox(c.handle) for c in comms])
        devs = (C.c_int * world)(*[int(d) for d in devices])
        for c, d in zip(comms, devices):
            with torch.cuda.device(d):
                _abi.check(lib.mrs_comm_connect_ptrs(c.handle, boxes, devs), 'mrs_comm_connect_ptrs')
        return comms

    def barrier(self, status=None, stream=None):
        """Device-side barrier on the current stream (no host synchronisation)."""
        C = self._C
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        self._abi.check(self.lib.mrs_comm_barrier(self.handle, C.c_void_p(status.data_ptr()) if status is not None else None, st),
                        'mrs_comm_barrier')

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.mrs_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """torch.distributed reduction of the counters (gloo in the CPU tests; NCCL when no PeerComm is used)."""
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
