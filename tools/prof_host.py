"""cProfile of the Python side of MRS.step for the README example (one env, three agents): where the per-call
host time goes.  python tools/prof_host.py (on a GPU box)."""
import cProfile, pstats, os, sys, io
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'mrs-gym_b200'))
import mrsgym_b200 as mrsgym
env = mrsgym.make('mrs-v0', N_AGENTS=3, K_HOPS=0, ACTION_TYPE='set_target_vel')
env.reset()
act = torch.tensor([[0.5, 0, 0]] * 3, device='cuda')
for _ in range(200): env.step(act)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(3000): env.step(act)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
