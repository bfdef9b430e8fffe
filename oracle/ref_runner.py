"""TEST INFRASTRUCTURE -- runs the reference's own Python VERBATIM on the fake backend.

``load_reference()`` puts the gym/ray stubs and ``oracle.fake_pybullet`` (as module name
``pybullet``) in front of ``sys.path``/``sys.modules`` and imports ``mrsgym`` from
``baseline/_ref`` (if a driver put an install there) or ``/root/reference``.  Nothing of
the reference is copied; it is imported where it lies.  Used only by
``oracle/make_golden.py`` (in the build container) and by the CPU-baseline leg of
``bench.py`` when a reference tree is present -- never by the product path.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def reference_root():
    for cand in (os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if os.path.isfile(os.path.join(cand, 'mrsgym', 'MRS.py')):
            return cand
    return None


def load_reference():
    """Returns (mrsgym module, fake pybullet module).  Raises if no reference tree."""
    root = reference_root()
    if root is None:
        raise RuntimeError('no reference tree (baseline/_ref or /root/reference)')
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    stubs = os.path.join(_HERE, 'stubs')
    for p in (root, stubs):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    fake = importlib.import_module('oracle.fake_pybullet')
    sys.modules['pybullet'] = fake
    mrsgym = importlib.import_module('mrsgym')
    assert os.path.dirname(os.path.dirname(mrsgym.__file__)) == root, mrsgym.__file__
    return mrsgym, fake


def state_fn_pos_vel(quad):
    """The ubiquitous state_fn of the reference README (README.md:28-29)."""
    import torch
    return torch.cat([quad.get_pos(), quad.get_vel()])


def read_state64(env, fake):
    """Exact float64 backend state of the N agents of a reference MRS env."""
    w = fake._WORLDS[env.sim.id]
    uids = [a.uid for a in env.env.agents]
    return dict(pos=np.stack([w.pos[u] for u in uids]), quat=np.stack([w.quat[u] for u in uids]),
                vel=np.stack([w.vel[u] for u in uids]), angvel=np.stack([w.angvel[u] for u in uids]))


def write_state64(env, fake, pos, quat, vel, angvel):
    """Upload an exact float64 state, bypassing Object.set_state's euler round trip."""
    w = fake._WORLDS[env.sim.id]
    for i, a in enumerate(env.env.agents):
        w.pos[a.uid] = np.array(pos[i], dtype=np.float64)
        w.quat[a.uid] = np.array(quat[i], dtype=np.float64)
        w.vel[a.uid] = np.array(vel[i], dtype=np.float64)
        w.angvel[a.uid] = np.array(angvel[i], dtype=np.float64)


def make_env(mrsgym, fake, N, action_type, K=0, comm_range=float('inf'), agent_radius=0.3,
             start_pos=None, start_ori=None, dt=0.01, state_fn=state_fn_pos_vel, **kw):
    import torch
    fake.PHYSICS.agent_radius = agent_radius
    if start_pos is None:
        # deterministic, collision-free default so the reference's rejection sampler
        # (mrsgym/MRS.py:127-154) is not exercised by construction
        g = int(np.ceil(N ** (1 / 3)))
        grid = np.array([[i, j, k] for k in range(g) for j in range(g) for i in range(g)][:N], dtype=np.float32)
        start_pos = torch.tensor(grid) + torch.tensor([0.0, 0.0, 2.0])
    if start_ori is None:
        start_ori = torch.zeros(N, 3)
    env = mrsgym.MRS(state_fn=state_fn, N_AGENTS=N, K_HOPS=K, COMM_RANGE=comm_range,
                     ACTION_TYPE=action_type, HEADLESS=True, AGENT_RADIUS=agent_radius,
                     START_POS=start_pos, START_ORI=start_ori, DT=dt, **kw)
    return env
