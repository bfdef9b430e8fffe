"""CPU checks of the contact model of the oracle (oracle/bullet_model.solve_contacts), the behaviour the CUDA
kernels mirror: tournament schedule, resting on the ground, a tilted landing righting itself, a heap of
overlapping agents spreading out and coming to rest.  (The GPU counterparts are in tests/test_gpu_parity.py.)"""
import numpy as np
from scipy.spatial.transform import Rotation as R

from oracle import bullet_model as bm


def _round_formula(i, j, N):
    """csrc/mrs_device.cuh tour_round: the round in which i < j meet."""
    M = N + (N & 1)
    m1 = M - 1
    return i if j == m1 else ((i + j) * (M // 2)) % m1


def _partner_formula(r, i, N):
    """csrc/mrs_device.cuh tour_partner."""
    M = N + (N & 1)
    m1 = M - 1
    if i == m1:
        j = r
    else:
        j = 2 * r - i
        if j < 0:
            j += m1
        if j >= m1:
            j -= m1
        if j == i:
            j = m1
    return i if j >= N else j


def test_tournament_schedule_meets_every_pair_once():
    for N in (2, 3, 4, 5, 8, 16, 31, 32, 40, 129):
        part = bm.tournament_partner(N)
        seen = set()
        for r in range(part.shape[0]):
            for i in range(N):
                j = int(part[r, i])
                assert j == _partner_formula(r, i, N)
                assert int(part[r, j]) == i                       # symmetric within a round
                if i < j:
                    assert (i, j) not in seen
                    seen.add((i, j))
                    assert _round_formula(i, j, N) == r
        assert len(seen) == N * (N - 1) // 2


def _run(pos, quat, steps, P=None, hook=None):
    P = P or bm.PhysicsParams()
    v, w = np.zeros_like(pos), np.zeros_like(pos)
    F, T = np.zeros_like(pos), np.zeros_like(pos)
    for t in range(steps):
        pos, quat, v, w = bm.bullet_step(pos, quat, v, w, F, T, P)
        if hook:
            hook(t, pos, quat, v, w)
    return pos, quat, v, w


def test_flat_drop_rests_on_the_rim_points():
    P = bm.PhysicsParams()
    pos = np.array([[[0.0, 0.0, 0.6]]])
    quat = np.array([[[0.0, 0.0, 0.0, 1.0]]])
    pos, quat, v, w = _run(pos, quat, 300, P)
    rest = P.ground_z + P.col_halfheight + P.col_margin - P.slop
    assert abs(pos[0, 0, 2] - rest) < 1e-6
    assert np.abs(v).max() < 1e-9 and np.abs(w).max() < 1e-9


def test_tilted_landing_rights_itself():
    quat = R.from_euler('xyz', [0.4, 0.2, 0.1]).as_quat().reshape(1, 1, 4)
    pos = np.array([[[0.0, 0.0, 0.62]]])
    pos, quat, v, w = _run(pos, quat, 150)
    roll, pitch, _ = R.from_quat(quat.reshape(4)).as_euler('xyz')
    assert abs(roll) < 1e-4 and abs(pitch) < 1e-4
    assert abs(pos[0, 0, 2] - 0.51349) < 1e-5
    # and upside down it comes to rest on its top face
    quat = R.from_euler('xyz', [np.pi - 0.3, 0.1, 0.0]).as_quat().reshape(1, 1, 4)
    pos, quat, v, w = _run(np.array([[[0.0, 0.0, 0.62]]]), quat, 150)
    assert bm.quat_to_mat(quat)[0, 0, 2, 2] < -0.9999 and abs(pos[0, 0, 2] - 0.51349) < 1e-5


def test_heap_spreads_and_rests():
    """16 agents dropped 0.5 m apart (contact spheres of 2 x 0.3 m overlap): pushed apart to 0.6 m, then at rest"""
    N = 16
    rng = np.random.default_rng(0)
    g = np.array([[i, j] for i in range(4) for j in range(4)], float)
    pos = np.zeros((1, N, 3))
    pos[0, :, :2] = g * 0.5 + rng.uniform(-0.02, 0.02, (N, 2))
    pos[0, :, 2] = 0.56 + rng.uniform(0, 0.05, N)
    quat = R.from_euler('xyz', np.concatenate([rng.uniform(-0.1, 0.1, (N, 2)), rng.uniform(-1, 1, (N, 1))], 1)
                        ).as_quat().reshape(1, N, 4)
    snap = {}

    def hook(t, p, q, v, w):
        if t == 219:
            snap['p'] = p.copy()
    pos, quat, v, w = _run(pos, quat, 340, hook=hook)
    d = np.linalg.norm(pos[0, :, None, :] - pos[0, None, :, :], axis=-1) + 9 * np.eye(N)
    assert d.min() > 0.6 - 1e-3
    assert np.abs(pos - snap['p']).max() < 1e-3          # 120 steps after settling: within a millimetre
    assert np.abs(v).max() < 1e-4


def _levels(pairs, rounds, n_agents):
    """csrc/mrs_contact_env.cuh: the relaxation that turns tournament rounds into dependency levels -- per-agent pair
    lists sorted by round, level[k] >= 1 + level[previous pair of either agent], iterated to the fixed point."""
    inc = [[] for _ in range(n_agents)]
    for k, (i, j) in enumerate(pairs):
        inc[i].append(k)
        inc[j].append(k)
    for lst in inc:
        lst.sort(key=lambda k: rounds[k])
    level = [1] * len(pairs)
    changed = True
    while changed:
        changed = False
        for lst in inc:
            run = 0
            for k in lst:
                lv = max(level[k], run + 1)
                if lv > level[k]:
                    level[k] = lv
                    changed = True
                run = lv
    return level


def test_level_schedule_of_the_per_env_solver_equals_the_round_order():
    """The per-env contact kernel (N > 32) walks dependency levels instead of tournament rounds.  Pairs of one level
    must touch disjoint agents, pairs that share an agent must keep their round order, and a non-commutative update
    applied level by level must give exactly what the round-by-round walk gives."""
    rng = np.random.default_rng(12)
    for N, p in ((40, 0.08), (64, 0.05), (129, 0.02), (200, 0.03), (33, 0.5)):
        pairs = [(i, j) for i in range(N) for j in range(i + 1, N) if rng.random() < p]
        rounds = [_round_formula(i, j, N) for i, j in pairs]
        level = _levels(pairs, rounds, N)
        for lv in set(level):
            touched = [a for k, pr in enumerate(pairs) if level[k] == lv for a in pr]
            assert len(touched) == len(set(touched))                   # a level is a matching
        for a in range(N):
            mine = sorted((rounds[k], level[k]) for k, pr in enumerate(pairs) if a in pr)
            assert all(x[1] < y[1] for x, y in zip(mine, mine[1:]))     # round order kept along every agent
        assert max(level, default=0) <= max(len(set(rounds)), 1)       # never more barriers than occupied rounds

        def walk(order):
            v = np.arange(1.0, N + 1.0)
            for k in order:
                i, j = pairs[k]
                d = 0.37 * (v[i] - v[j]) + 0.001 * v[i] * v[j]          # order-sensitive, like a Gauss-Seidel row
                v[i] -= d
                v[j] += 0.5 * d
            return v
        by_round = sorted(range(len(pairs)), key=lambda k: (rounds[k], k))
        by_level = sorted(range(len(pairs)), key=lambda k: (level[k], -k))
        assert np.array_equal(walk(by_round), walk(by_level))
