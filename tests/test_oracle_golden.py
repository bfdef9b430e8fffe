"""Pins the oracle port (oracle/spec.py) against the golden vectors that the reference's
own Python produced on the fake backend (oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import spec
from oracle import bullet_model as bm

from conftest import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, 'ref_*.npz')))


def run_spec(g, T=None):
    N, K = int(g['N']), int(g['K'])
    T = T or int(g['T'])
    phys = bm.PhysicsParams(agent_radius=float(g['agent_radius']))
    env = spec.SpecEnv(1, N, str(g['mode']), K=K, comm_range=float(g['comm_range']), dt=float(g['dt']),
                       gravity=float(g['gravity']), phys=phys)
    none_steps = set(int(t) for t in g['none_steps'])
    env.set_state(pos=g['start_pos'], quat=g['start_quat'], vel=g['start_vel'], angvel=g['start_angvel'])
    X0 = env.reset_rings()
    out = dict(X0=X0[0], pos=[], quat=[], vel=[], angvel=[], rpm=[], force=[], torque=[], X=[], A=[])
    for t in range(T):
        X, A = env.step(None if t in none_steps else g['actions'][t][None])
        for k in ('pos', 'quat', 'vel', 'angvel'):
            out[k].append(getattr(env, k)[0].copy())
        for k in ('rpm', 'force', 'torque'):
            out[k].append(env.last[k][0].copy())
        out['X'].append(X[0])
        out['A'].append(A[0])
    return {k: (np.stack(v) if isinstance(v, list) else v) for k, v in out.items()}


@pytest.mark.parametrize('path', FILES, ids=[os.path.basename(f)[4:-4] for f in FILES])
def test_spec_matches_reference_verbatim(path):
    g = np.load(path)
    o = run_spec(g)
    np.testing.assert_array_equal(o['X0'], g['X0'])
    # rpm: same float64 expressions, same float32 getter roundings
    # The contact solver stops sweeping below solver_tol: a last-bit difference in its input can move the exit by
    # one sweep, i.e. change the velocities by up to that threshold (an iterative solver is not continuous at its
    # exit test; Bullet's leastSquaresResidualThreshold behaves the same).  Contact files get that band.
    contact = 'contact' in os.path.basename(path)
    np.testing.assert_allclose(o['rpm'], g['rpm'], rtol=1e-5 if contact else 1e-9, atol=0)
    np.testing.assert_allclose(o['force'], g['force'], rtol=1e-4 if contact else 1e-7, atol=1e-7 if contact else 1e-12)
    np.testing.assert_allclose(o['torque'], g['torque'], rtol=1e-4 if contact else 1e-7, atol=1e-9 if contact else 1e-13)
    for k in ('pos', 'quat', 'vel', 'angvel'):
        tol = {'pos': 1e-6, 'quat': 1e-5, 'vel': 5e-6, 'angvel': 5e-5}[k] if contact else 2e-8
        np.testing.assert_allclose(o[k], g[k], rtol=0, atol=tol, err_msg=k)
    # observation windows: float32 views of the state => equal up to a float32 ulp
    np.testing.assert_allclose(o['X'], g['X'], rtol=2e-7, atol=5e-6 if contact else 1e-7)
    assert o['A'].shape == g['A'].shape
    mism = np.sum(o['A'] != g['A'])
    assert mism == 0, 'adjacency differs in %d entries' % mism


def test_goldens_exist():
    assert len(FILES) >= 17
