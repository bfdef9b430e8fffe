#!/usr/bin/env python
"""Multi-GPU check of the peer-memory communicator (mrs_comm_* / mrs_stats_allreduce), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py

Every rank fills its counters with rank-dependent values, reduces them 200 times through the library's kernel
(eagerly and as CUDA-graph replays), compares every result with the closed form and with NCCL, and times one
reduction and one device barrier with CUDA events.  Rank 0 prints one JSON line."""
import json
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (_REPO, os.path.join(_REPO, 'mrs-gym_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def main():
    import torch
    import torch.distributed as dist
    import mrsgym_b200 as M
    from mrsgym_b200 import dist as D
    rank, world = D.init_from_env()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    sw = M.Swarm(32, 8, 0, 'set_speeds', M._abi.X_POS_VEL, 2.0)
    comm = D.PeerComm()
    ok = True
    for it in range(200):
        sw.stats.copy_(torch.arange(8, dtype=torch.int64, device=dev) * (rank + 1) + it)
        out = sw.allreduce_stats(comm)
        want = torch.arange(8, dtype=torch.int64, device=dev) * (world * (world + 1) // 2) + it * world
        ok = ok and bool((out == want).all())
    ref = sw.stats.clone()
    dist.all_reduce(ref)
    ok = ok and bool((sw.allreduce_stats(comm) == ref).all())
    # graph replays
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        sw.allreduce_stats(comm)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(4):
                sw.allreduce_stats(comm)
    torch.cuda.current_stream().wait_stream(side)
    for it in range(50):
        g.replay()
    torch.cuda.synchronize()
    ok = ok and bool((sw._stats_sum == ref).all())
    # latency: device-aligned start, then n reductions / n barriers back to back
    def timed(fn, n=100):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda._sleep(int(1e6))
        comm.barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return D.max_over_ranks(e0.elapsed_time(e1) / n * 1e3, dev)
    us_reduce = timed(lambda: sw.allreduce_stats(comm))
    us_barrier = timed(lambda: comm.barrier())
    tiny = torch.zeros(8, dtype=torch.int64, device=dev)
    us_nccl = timed(lambda: dist.all_reduce(tiny))
    status = sw.read_status()
    oks = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({'world': world, 'ok': bool(oks.item()), 'status_word': status, 'us_per_stats_allreduce': us_reduce,
                          'us_per_barrier': us_barrier, 'us_per_nccl_allreduce_64B': us_nccl}))
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
