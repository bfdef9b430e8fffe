"""Profiling target: C5 shape, a few single-step launches followed by one multi-step launch.

    ncu --set full --cache-control none --clock-control none --import-source on \
        -k regex:step_group -s 3 -c 4 -o gpurun_out/prof python tools/prof_c5.py [workload]
"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'mrs-gym_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import mrsgym_b200 as M, helpers as H
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c5']
E, N, K = w['E'], w['N'], w['K']
T = 24
st, act = bench.make_inputs(w, E, T, 1)
sw = M.Swarm(E, N, K, w['mode'], M._abi.X_POS_VEL, w['R'], tape_slots=T + 2 * K + 2, want_A=w.get('A', True))
H.upload_state(sw, st)
actions = torch.from_numpy(act).cuda()
for t in range(6):
    sw.step(actions[t])
torch.cuda.synchronize()
sw.step_many(actions[6:6 + 16].contiguous(), 16)
torch.cuda.synchronize()
print('status', sw.read_status())
