"""TEST INFRASTRUCTURE -- generates tests/golden/ref_*.npz by running the reference's
own Python (/root/reference/mrsgym, imported where it lies) verbatim on
oracle/fake_pybullet.  Run in the build container only (the GPU box has no
/root/reference):   python -m oracle.make_golden

Each file: start state (float64), float32 actions [T,N,A] and, per step, the exact
float64 backend state, rotor rpm, applied world wrench, and the X / A windows the
reference returned.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _REPO)

from oracle import ref_runner  # noqa: E402

OUT = os.path.join(_REPO, 'tests', 'golden')

ANCHOR_STATE = dict(pos=[[1.0, 2.0, 3.0]], rpy=[0.1, -0.2, 0.3], vel=[[0.3, -0.1, 0.2]],
                    angvel=[[0.5, -0.4, 0.2]])


ONLY = None          # --only a,b: write just these files (the random stream of main() is consumed all the same)


def rollout(mrsgym, fake, name, N, mode, T, K, comm_range, start, actions, agent_radius=0.3, dt=0.01, gravity=9.81,
            none_steps=()):
    import torch
    if ONLY is not None and name not in ONLY:
        return
    env = ref_runner.make_env(mrsgym, fake, N, mode, K=K, comm_range=comm_range,
                              agent_radius=agent_radius, dt=dt, GRAVITY=gravity)
    ref_runner.write_state64(env, fake, **start)
    X0 = env.set()          # MRS.set with no args: keeps state, clears rings (MRS.py:196-205)
    # Object.set_state round-trips the state through float32 getters / euler: re-upload exact
    ref_runner.write_state64(env, fake, **start)
    env.X.clear()
    X0 = env.calc_Xk()
    rec = {k: [] for k in ('pos', 'quat', 'vel', 'angvel', 'rpm', 'force', 'torque', 'X', 'A')}
    for t in range(T):
        fake.RECORD = []
        X, reward, done, info = env.step(None if t in none_steps else torch.tensor(actions[t]))
        st = ref_runner.read_state64(env, fake)
        for k in ('pos', 'quat', 'vel', 'angvel'):
            rec[k].append(st[k])
        rec['rpm'].append(np.stack([np.asarray(a.speeds, np.float64) for a in env.env.agents]))
        rec['force'].append(fake.RECORD[0]['force'])
        rec['torque'].append(fake.RECORD[0]['torque'])
        rec['X'].append(X.numpy())
        rec['A'].append(info['A'].numpy())
    fake.RECORD = None
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(start_pos=np.asarray(start['pos'], np.float64), start_quat=np.asarray(start['quat'], np.float64),
               start_vel=np.asarray(start['vel'], np.float64), start_angvel=np.asarray(start['angvel'], np.float64),
               X0=X0.numpy(), actions=np.asarray(actions, np.float32), mode=mode, N=N, K=K, T=T,
               comm_range=comm_range, agent_radius=agent_radius, dt=dt, gravity=gravity,
               none_steps=np.asarray(sorted(none_steps), np.int64))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'ref_%s.npz' % name), **out)
    print('wrote ref_%s.npz  N=%d T=%d mode=%s' % (name, N, T, mode))
    env.close()


def rand_start(rng, N, spacing=1.0, z0=2.0, jitter=0.1, tilt=0.15, vel=0.3, angvel=0.5):
    from scipy.spatial.transform import Rotation as R
    g = int(np.ceil(N ** (1 / 3)))
    grid = np.array([[i, j, k] for k in range(g) for j in range(g) for i in range(g)][:N], dtype=np.float64)
    pos = grid * spacing + np.array([0, 0, z0]) + rng.uniform(-jitter, jitter, (N, 3))
    rpy = np.concatenate([rng.uniform(-tilt, tilt, (N, 2)), rng.uniform(-np.pi / 2, np.pi / 2, (N, 1))], axis=1)
    quat = R.from_euler('xyz', rpy).as_quat()
    return dict(pos=pos, quat=quat, vel=rng.uniform(-vel, vel, (N, 3)), angvel=rng.uniform(-angvel, angvel, (N, 3)))


def main():
    from scipy.spatial.transform import Rotation as R
    mrsgym, fake = ref_runner.load_reference()
    hover = 14475.809

    # --- single-agent anchors, first controller call
    q = R.from_euler('xyz', ANCHOR_STATE['rpy']).as_quat()[None]
    st = dict(pos=np.array(ANCHOR_STATE['pos']), quat=q, vel=np.array(ANCHOR_STATE['vel']),
              angvel=np.array(ANCHOR_STATE['angvel']))
    anchors = [('vel', 'set_target_vel', [0.5, 0, 0]), ('pos', 'set_target_pos', [1.5, 2, 3.5]),
               ('accel', 'set_target_accel', [1, 0, 0.5]), ('ori', 'set_target_ori', [0, 0.1, 0]),
               ('control', 'set_control', [9.81, 0.1, -0.1, 0.05]),
               ('control_nnls', 'set_control', [3, 40, -30, 5]),
               ('speeds', 'set_speeds', [hover * 1.02, hover * 0.97, hover, hover * 1.05])]
    for nm, mode, act in anchors:
        rollout(mrsgym, fake, 'anchor_' + nm, 1, mode, 3, 0, float('inf'), st,
                np.tile(np.array(act, np.float32), (3, 1, 1)))

    # --- C1: README example, N=3 set_target_vel K=0 (BASELINE.json configs[0])
    st = dict(pos=np.array([[0, 0, 1.5], [1, 0, 2], [-1, 0, 2.5]]), quat=np.tile([0, 0, 0, 1.0], (3, 1)),
              vel=np.zeros((3, 3)), angvel=np.zeros((3, 3)))
    rollout(mrsgym, fake, 'c1_vel', 3, 'set_target_vel', 100, 0, float('inf'), st,
            np.tile(np.array([0.5, 0, 0], np.float32), (100, 3, 1)))

    # --- multi-agent trajectories per mode, K=2/3, finite COMM_RANGE, downwash active
    rng = np.random.default_rng(20261018)
    T = 100
    N = 8
    for nm, mode in [('vel', 'set_target_vel'), ('pos', 'set_target_pos'), ('accel', 'set_target_accel'),
                     ('ori', 'set_target_ori'), ('control', 'set_control'), ('speeds', 'set_speeds')]:
        st = rand_start(rng, N, spacing=0.9)
        if mode == 'set_target_vel':
            a = rng.normal(0, 0.5, (1, N, 3)).repeat(T, 0) + rng.normal(0, 0.05, (T, N, 3))
        elif mode == 'set_target_pos':
            a = (st['pos'] + rng.normal(0, 0.5, (N, 3)))[None].repeat(T, 0)
        elif mode == 'set_target_accel':
            a = rng.normal(0, 1.0, (T, N, 3))
        elif mode == 'set_target_ori':
            a = rng.uniform(-0.2, 0.2, (1, N, 3)).repeat(T, 0)
        elif mode == 'set_control':
            a = np.stack([9.81 + rng.uniform(-1, 1, (T, N)), rng.uniform(-1, 1, (T, N)),
                          rng.uniform(-1, 1, (T, N)), rng.uniform(-1, 1, (T, N))], axis=-1)
            a[::7, :, 1:] *= 60.0          # drive some rows into the NNLS branch
        else:
            a = hover * (1 + 0.05 * rng.normal(0, 1, (T, N, 4)))
        rollout(mrsgym, fake, 'traj_' + nm, N, mode, T, 3 if mode != 'set_target_vel' else 2, 1.5, st,
                a.astype(np.float32))

    # --- sim constants other than the defaults (BulletSim DT / GRAVITY kwargs, BulletSim.py:11-24): the
    # controller keeps DefaultSim's 0.01 / 9.81 (QuadControl.py:10) while the integrator follows the kwargs
    st = rand_start(rng, 6, spacing=1.1)
    a = rng.normal(0, 0.4, (1, 6, 3)).repeat(60, 0).astype(np.float32)
    rollout(mrsgym, fake, 'traj_vel_dt005_g371', 6, 'set_target_vel', 60, 1, 2.0, st, a, dt=0.005, gravity=3.71)
    # --- MRS.step(None) interleaved with actions: no forces at all on those steps (MRS.py:243,252)
    st = rand_start(rng, 5, spacing=1.0)
    a = (hover * (1 + 0.03 * rng.normal(0, 1, (40, 5, 4)))).astype(np.float32)
    rollout(mrsgym, fake, 'traj_speeds_none_steps', 5, 'set_speeds', 40, 2, float('inf'), st, a,
            none_steps=(0, 7, 8, 9, 25))

    # --- contact: ground landing + sphere-sphere (AGENT_RADIUS 0.3), set_control free-fall-ish
    N = 4
    st = rand_start(rng, N, spacing=0.55, z0=0.62, jitter=0.02, tilt=0.05, vel=0.2, angvel=0.2)
    st['pos'][:, 2] = np.array([0.56, 0.60, 0.7, 0.9])
    st['vel'][:, 2] = np.array([-0.5, -1.0, 0.0, -2.0])
    a = np.stack([rng.uniform(4, 8, (T, N)), rng.uniform(-1, 1, (T, N)), rng.uniform(-1, 1, (T, N)),
                  rng.uniform(-1, 1, (T, N))], axis=-1)
    rollout(mrsgym, fake, 'contact_control', N, 'set_control', T, 1, 2.0, st, a.astype(np.float32))

    # --- larger swarms: N = 16 (two envs per warp in the group kernel, half-round pair schedule) and N = 40
    # (first size of the wide path), K_HOPS 3 / 1, finite COMM_RANGE
    st = rand_start(rng, 16, spacing=0.9)
    a = (rng.normal(0, 0.5, (1, 16, 3)).repeat(40, 0) + rng.normal(0, 0.05, (40, 16, 3))).astype(np.float32)
    rollout(mrsgym, fake, 'traj_vel_n16', 16, 'set_target_vel', 40, 3, 1.5, st, a)
    st = rand_start(rng, 40, spacing=1.0)
    a = (st['pos'] + rng.normal(0, 0.4, (40, 3)))[None].repeat(30, 0).astype(np.float32)
    rollout(mrsgym, fake, 'traj_pos_n40', 40, 'set_target_pos', 30, 1, 1.8, st, a)
    # --- the shapes of BASELINE configs[1] / configs[2] as single envs: 32 agents set_target_pos K_HOPS 3
    # COMM_RANGE 2.0 (one env = one warp of the group kernel), 16 agents set_control in the contact regime
    st = rand_start(rng, 32, spacing=1.0, z0=3.0)
    a = (st['pos'] + rng.normal(0, 0.5, (32, 3)))[None].repeat(25, 0).astype(np.float32)
    rollout(mrsgym, fake, 'traj_pos_n32_c2', 32, 'set_target_pos', 25, 3, 2.0, st, a)
    st = rand_start(rng, 16, spacing=0.7, z0=0.95, jitter=0.03, tilt=0.05, vel=0.2, angvel=0.2)
    st['vel'][:, 2] -= 0.8
    a = np.stack([9.81 + rng.uniform(-1, 1, (40, 16)), rng.uniform(-1, 1, (40, 16)), rng.uniform(-1, 1, (40, 16)),
                  rng.uniform(-1, 1, (40, 16))], axis=-1).astype(np.float32)
    rollout(mrsgym, fake, 'contact_control_n16_c3', 16, 'set_control', 40, 0, float('inf'), st, a)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--only':
        ONLY = set(sys.argv[2].split(','))
    main()
