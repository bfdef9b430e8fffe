"""Debug: per-warp timeline of step_group_kernel at C5 (needs a -DMRS_TRACE build via MRS_B200_LIB)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'mrs-gym_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import mrsgym_b200 as M, helpers as H
import bench
w = bench.WORKLOADS['c5']
E, N, K = w['E'], w['N'], w['K']
T = 8
st, act = bench.make_inputs(w, E, T, 1)
sw = M.Swarm(E, N, K, w['mode'], M._abi.X_POS_VEL, w['R'], tape_slots=T, ring=True)
sw.scratch = torch.zeros(max(7, sw.lib.mrs_scratch_planes(sw.E, sw.N)), sw.S, device='cuda')
sw._bind()
H.upload_state(sw, st)
actions = torch.from_numpy(act).cuda()
roll = sw.capture_rollout(actions, T)
for _ in range(3):
    roll.replay()
torch.cuda.synchronize()
# the trace holds the LAST step of the last replay; to see consecutive kernels, snapshot after single steps
traces = []
for t in range(4):
    sw.step(actions[t])
torch.cuda.synchronize()
def grab():
    raw = sw.scratch.view(-1)[:4096 * 16].clone().cpu().numpy().view(np.uint64).reshape(-1, 8)
    return raw[raw[:, 0] > 0]
tr = grab().astype(np.int64)
t0 = tr[:, 0].min()
rel = (tr - t0) / 1e3
names = ['start', 'after_wait', 'chunk0_end', 'chunk1_end', 'chunk2_end', 'chunk3_end', 'c4', 'c5']
print('warps', len(tr))
for i, n in enumerate(names[:6]):
    c = rel[:, i]
    print('%-11s min %7.2f  p10 %7.2f  med %7.2f  p90 %7.2f  max %7.2f us' % (n, c.min(), np.percentile(c, 10), np.median(c), np.percentile(c, 90), c.max()))
d = np.diff(rel[:, 1:6], axis=1)
print('chunk durations (us): med', np.median(d, axis=0).round(2), 'p90', np.percentile(d, 90, axis=0).round(2))
# per-step aggregates over one graph replay
L = sw.L
aggv = sw.scratch.view(-1)[8192 * 16: 8192 * 16 + L * 8]
def reset_agg():
    a = np.zeros((L, 4), np.uint64); a[:, 0] = np.iinfo(np.uint64).max; a[:, 3] = np.iinfo(np.uint64).max
    aggv.copy_(torch.from_numpy(a.view(np.float32).reshape(-1)).cuda())
reset_agg(); torch.cuda.synchronize()
roll.replay(); torch.cuda.synchronize()
a = aggv.clone().cpu().numpy().view(np.uint64).reshape(L, 4).astype(np.int64)
rows = [(sl, a[sl]) for sl in range(L) if a[sl, 1] > 0]
rows.sort(key=lambda r: r[1][1])
base = rows[0][1][1]
print('step: first_start  wait_release  first_end  last_end   | duration  gap_to_next_release (us, relative)')
for i, (sl, r) in enumerate(rows):
    nxt = rows[i + 1][1][1] if i + 1 < len(rows) else None
    print('slot %3d: %8.2f %8.2f %8.2f %8.2f | dur %6.2f  period %s' % (sl, (r[0] - base) / 1e3, (r[1] - base) / 1e3, (r[3] - base) / 1e3, (r[2] - base) / 1e3, (r[2] - r[1]) / 1e3, ('%6.2f' % ((nxt - r[1]) / 1e3)) if nxt else '-'))
# graph replay: timeline of the last kernel of a replay relative to kernel start spread
roll.replay(); torch.cuda.synchronize()
tr = grab().astype(np.int64); t0 = tr[:, 0].min(); rel = (tr - t0) / 1e3
print('in-graph last step: start spread max %.2f us, after_wait med %.2f, end med %.2f max %.2f' % (rel[:, 0].max(), np.median(rel[:, 1]), np.median(rel[:, 5]), rel[:, 5].max()))
