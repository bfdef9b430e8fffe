"""GPU checks of the peer-memory statistics reduction (mrs_comm_* / mrs_stats_allreduce, include/mrs_b200.h).
The world-size-2 case runs in ONE process over two devices (mrs_comm_connect_ptrs) and is skipped on a
single-GPU box; the one-process-per-GPU case (cudaIpc handles) is tools/comm_check.py under torchrun."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_stats_allreduce_world_1_is_a_copy():
    import mrsgym_b200 as M
    from mrsgym_b200 import dist as D
    sw = M.Swarm(4, 8, 0, 'set_speeds', M._abi.X_POS_VEL, 2.0)
    comm = D.PeerComm(rank=0, world=1)
    sw.stats.copy_(torch.arange(8, dtype=torch.int64, device='cuda') + 5)
    out = sw.allreduce_stats(comm)
    comm.barrier()
    torch.cuda.synchronize()
    assert out.tolist() == list(range(5, 13))
    assert sw.stats.tolist() == list(range(5, 13))          # the local counters are left alone
    assert sw.read_status() == 0
    comm.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_stats_allreduce_two_devices_one_process():
    import mrsgym_b200 as M
    from mrsgym_b200 import dist as D
    comms = D.PeerComm.local_group([0, 1])
    sws = []
    for d in (0, 1):
        with torch.cuda.device(d):
            sws.append(M.Swarm(4, 8, 0, 'set_speeds', M._abi.X_POS_VEL, 2.0, device='cuda:%d' % d))
    for it in range(20):
        outs = []
        for d in (0, 1):
            with torch.cuda.device(d):
                sws[d].stats.copy_(torch.arange(8, dtype=torch.int64, device='cuda:%d' % d) * (d + 1) + it)
                outs.append(sws[d].allreduce_stats(comms[d]))
        for d in (0, 1):
            torch.cuda.synchronize(d)
        want = [3 * i + 2 * it for i in range(8)]
        assert outs[0].tolist() == want and outs[1].tolist() == want
    for d in (0, 1):
        with torch.cuda.device(d):
            assert sws[d].read_status() == 0
            comms[d].close()
