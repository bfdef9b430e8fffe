"""Soak: two independent swarms run the same long graph rollout (chained single-step launches: dynamic chunk
hand-out inside a CTA, range hand-over between launches) and one runs the same steps as plain stream-ordered
mrs_step launches; all three must end bit-identical, finite, with unit quaternions and a symmetric, zero-diagonal
adjacency.  A fourth runs multi-step launches (separately compiled kernel: float32 rounding level, checked over the
first 10 steps -- a tumbling swarm is chaotic, rounding differences grow e-fold every few steps).
python tools/soak.py [workload] [steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'mrs-gym_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import mrsgym_b200 as M, helpers as H
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c5']
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
E, N, K, T = min(w['E'], 16384 + 3), w['N'], w['K'], 100          # +3: ragged last warp-chunk for N = 8
st, act = bench.make_inputs(w, E, T, 5)
actions = torch.from_numpy(act).cuda()
sws = []
for i in range(4):
    sw = M.Swarm(E, N, K, w['mode'], M._abi.X_POS_VEL, w['R'], tape_slots=T, ring=True)
    H.upload_state(sw, st)
    sws.append(sw)
rolls = [sws[0].capture_rollout(actions, T), sws[1].capture_rollout(actions, T)]
sws[3].step_many(actions[:10], 10)
for t in range(10):
    sws[2].step(actions[t])
torch.cuda.synchronize()
dev = float((sws[3].state - sws[2].state).abs().max())
assert dev < 1e-4, 'step_many drifts from the single-step kernels: %g' % dev
H.upload_state(sws[2], st)
sws[2].ctrl.copy_(sws[0].ctrl)
for r in range(steps // T):
    for roll in rolls:
        roll.replay()
    for t in range(T):
        sws[2].step(actions[t])
torch.cuda.synchronize()
a, b, c, _ = sws
assert torch.equal(a.state, b.state), 'two graph rollouts differ'
assert torch.equal(a.state, c.state), 'graph rollout and plain steps differ'
assert torch.equal(a.X_tape, b.X_tape) and torch.equal(a.A_tape, b.A_tape)
assert bool(torch.isfinite(a.state).all()) and a.read_status() == 0, a.read_status()
qn = (a.state[3:7] ** 2).sum(0)
assert float((qn - 1).abs().max()) < 1e-5, float((qn - 1).abs().max())
A = a.A_window()[0]
assert torch.equal(A, A.transpose(-1, -2)) and float(A.diagonal(dim1=-2, dim2=-1).abs().sum()) == 0.0
print('soak ok: %s, %d envs x %d agents, %d steps; stats %s; mean z %.3f' % (w['mode'], E, N, steps // T * T, a.read_stats(), float(a.state[2].mean())))
