"""Wall-clock latency of MRS.step (closed loop: one Python call and one launch per step, nothing overlapped by a
graph) for the README example (C1: one env, 3 agents), small batches and the C5 shape.  The agents start on a 1 m grid
2 m above the ground and hold their place (free flight: no agent on the contact path); the last line repeats the
32-agent case with every agent sent to the origin, i.e. a permanent heap, to show what the contact solver costs."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'mrs-gym_b200'))
import mrsgym_b200 as mrsgym
for E, N, mode, heap in ((1, 3, 'set_target_vel', False), (1, 32, 'set_target_pos', False), (256, 32, 'set_target_pos', False),
                         (4096, 16, 'set_control', False), (65536, 8, 'set_speeds', False), (256, 32, 'set_target_pos', True)):
    idx = torch.arange(N)
    start = torch.stack([(idx % 4).float(), ((idx // 4) % 4).float(), 2.0 + (idx // 16).float()], dim=1)
    env = mrsgym.make('mrs-v0', N_ENVS=E, N_AGENTS=N, K_HOPS=3 if N in (8, 32) else 0, COMM_RANGE=2.0, ACTION_TYPE=mode,
                      START_POS=start, START_ORI=torch.zeros(N, 3))
    adim = env.swarm.action_dim
    a_host = torch.zeros(E, N, adim) if E > 1 else torch.zeros(N, adim)
    if mode == 'set_target_pos' and not heap:
        a_host[...] = start
    if mode == 'set_control':
        a_host[..., 0] = 9.81
    if mode == 'set_speeds':
        a_host[...] = 14475.8
    a_dev = a_host.cuda()
    for name, a in (('host actions', a_host), ('device actions', a_dev)):
        for _ in range(20):
            env.step(a)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 300
        for _ in range(n):
            X, r, d, info = env.step(a)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print('E=%d N=%d %s%s, %s: %.1f us per env.step (%.3g agent-steps/s)' % (E, N, mode, ' (heap at the origin)' if heap else '',
                                                                               name, dt * 1e6, E * N / dt))
