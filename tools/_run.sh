python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/latency_c1.py 2>&1 | tail -8
python - <<"PY"
import os, sys, time, torch
sys.path.insert(0, 'mrs-gym_b200')
import mrsgym_b200 as mrsgym
env = mrsgym.make('mrs-v0', N_ENVS=65536, N_AGENTS=8, K_HOPS=3, COMM_RANGE=2.0, ACTION_TYPE='set_speeds')
env.reset()
a = torch.full((65536, 8, 4), 14475.8, device='cuda')
for _ in range(50): env.step(a)
torch.cuda.synchronize(); t0 = time.perf_counter(); n = 2000
for _ in range(n): X, r, d, info = env.step(a)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print('closed-loop MRS.step at C5, device actions: %.1f us per env.step (%.3g agent-steps/s)' % (dt * 1e6, 65536 * 8 / dt))
PY
