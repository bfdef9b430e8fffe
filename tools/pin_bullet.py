#!/usr/bin/env python
"""Pins the constants of oracle/bullet_model.PhysicsParams (and MrsPhysicsParams) against a REAL pybullet, the
script SURVEY.md 8(a-P) "pin these first" describes.  The integrator / contact arithmetic of the path
(/root/reference/mrsgym/BulletSim.py:46-47, p.stepSimulation) lives in the pybullet wheel, which is not installable
in this image: until it is, the oracle restates Bullet from memory and says PARITY UNPINNED.  The day a wheel is
importable, run

    python tools/pin_bullet.py --out tests/golden/bullet_pinned.json [--models /path/to/mrsgym/models]

and feed the JSON to bm.PhysicsParams.from_json / regenerate the goldens.  Without pybullet the script runs on
oracle/fake_pybullet (--backend fake; tests/test_cpu_host.py does, so the script is known to work) and then merely
reads the oracle's own constants back -- it says so in the JSON ("backend": "fake").

Probes (each through the 13 backend calls of SURVEY.md 8b seam 2, on the scene the reference builds: plane.urdf at
the origin + cf2x.urdf bodies, gravity (0, 0, -9.81), timestep 0.01, EnvCreator.py:7-13,60):
  dynamics info   getDynamicsInfo(quad, -1): mass, lateral friction, LOCAL INERTIA DIAGONAL, restitution, margin
  engine params   getPhysicsEngineParameters(): numSolverIterations, numSubSteps, contactERP, frictionERP, contactSlop
  drop            zero force, 5 steps: gravity, and the linear damping law a = -v (k + k |v|) from v(t)
  spin-down       resetBaseVelocity(w only): angular damping, gyroscopic coupling (w x Iw) with an off-axis spin
  link force      applyExternalForce on link 0 (LINK_FRAME, pos 0), one step: moment arm of the prop link CoM
  velocity clamp  a huge force, one step: max coordinate velocity
  rest            drop onto the ground, 400 steps: resting height (ground top + half height + margin - slop), ERP
  tilted landing  lands tilted by 0.4 rad: does it right itself (contact impulses act at the hull points)
  push-out        two quads 0.1 m apart at rest: at which centre distance does quad-quad contact stop pushing
                  (the reference collides the 0.06 m hull cylinders, the north star uses AGENT_RADIUS spheres)
  one-step goldens  seeded one-step (pos, quat, vel, angvel) after an external wrench for 8 random states
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def backend(name):
    if name in ('auto', 'pybullet'):
        try:
            import pybullet as p          # noqa: F401
            return p, 'pybullet'
        except Exception:
            if name == 'pybullet':
                raise
    from oracle import fake_pybullet as p
    return p, 'fake'


def models_dir(arg):
    if arg:
        return arg
    from oracle import ref_runner
    root = ref_runner.reference_root()
    return os.path.join(root, 'mrsgym', 'models') if root else ''


class Scene:
    def __init__(self, p, models, n_quads=1, dt=0.01, gravity=9.81):
        self.p = p
        self.cid = p.connect(p.DIRECT)
        p.setGravity(0, 0, -gravity, physicsClientId=self.cid)
        p.setTimeStep(dt, physicsClientId=self.cid)
        p.setRealTimeSimulation(0, physicsClientId=self.cid)
        self.dt = dt
        self.quads = [p.loadURDF(fileName=os.path.join(models, 'cf2x.urdf'), basePosition=[2.0 * i, 0, 5.0],
                                 baseOrientation=[0, 0, 0, 1], physicsClientId=self.cid) for i in range(n_quads)]
        self.plane = p.loadURDF(fileName=os.path.join(models, 'plane.urdf'), basePosition=[0, 0, 0],
                                baseOrientation=[0, 0, 0, 1], physicsClientId=self.cid)

    def set(self, i, pos, quat=(0, 0, 0, 1), vel=(0, 0, 0), angvel=(0, 0, 0)):
        self.p.resetBasePositionAndOrientation(self.quads[i], list(pos), list(quat), physicsClientId=self.cid)
        self.p.resetBaseVelocity(self.quads[i], list(vel), list(angvel), physicsClientId=self.cid)

    def get(self, i):
        pos, quat = self.p.getBasePositionAndOrientation(self.quads[i], physicsClientId=self.cid)
        vel, ang = self.p.getBaseVelocity(self.quads[i], physicsClientId=self.cid)
        return np.array(pos), np.array(quat), np.array(vel), np.array(ang)

    def step(self, n=1):
        for _ in range(n):
            self.p.stepSimulation(physicsClientId=self.cid)

    def close(self):
        self.p.disconnect(physicsClientId=self.cid)


def quat_from_euler(r, pt, y):
    cr, sr, cp, sp, cy, sy = math.cos(r / 2), math.sin(r / 2), math.cos(pt / 2), math.sin(pt / 2), math.cos(y / 2), math.sin(y / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def probe(p, models, seed=0):
    out = {}
    s = Scene(p, models)
    q = s.quads[0]
    di = p.getDynamicsInfo(q, -1, physicsClientId=s.cid)
    out['dynamics_info'] = dict(mass=di[0], lateral_friction=di[1], local_inertia_diag=list(di[2]), restitution=di[5],
                                collision_margin=(di[11] if len(di) > 11 else None))
    dp = p.getDynamicsInfo(s.plane, -1, physicsClientId=s.cid)
    out['plane_lateral_friction'] = dp[1]
    try:
        out['engine'] = {k: v for k, v in p.getPhysicsEngineParameters(physicsClientId=s.cid).items()
                         if isinstance(v, (int, float))}
    except Exception as exc:                     # very old wheels
        out['engine'] = {'error': str(exc)}
    dt = s.dt
    # drop: v_z after one step from rest = -g dt; damping from a fast body
    s.set(0, (0, 0, 5))
    s.step()
    out['gravity'] = float(-s.get(0)[2][2] / dt)
    s.set(0, (0, 0, 5), vel=(3.0, 0, 0))
    s.step()
    v1 = s.get(0)[2][0]
    out['lin_damping_accel_at_3mps'] = float((3.0 - v1) / dt)              # = 3 (k + 3 k) for the law a = -v (k + k |v|)
    out['lin_damping_k'] = float((3.0 - v1) / dt / (3.0 * (1.0 + 3.0)))
    # spin-down about body z, and an off-axis spin for the gyroscopic term
    s.set(0, (0, 0, 5), angvel=(0, 0, 2.0))
    s.step()
    w1 = s.get(0)[3][2]
    out['ang_damping_k'] = float((2.0 - w1) / dt / (2.0 * (1.0 + 2.0)))
    s.set(0, (0, 0, 5), angvel=(3.0, 0, 4.0))
    s.step()
    out['gyro_step_angvel'] = s.get(0)[3].tolist()                          # w_y != 0 iff the w x Iw term is applied
    # force on link 0 in its own frame at its CoM: torque arm
    s.set(0, (0, 0, 5))
    p.applyExternalForce(q, 0, [0, 0, 0.01], [0, 0, 0], p.LINK_FRAME, physicsClientId=s.cid)
    s.step()
    _, _, v, w = s.get(0)
    out['link0_force_step'] = dict(vel=v.tolist(), angvel=w.tolist())
    I = out['dynamics_info']['local_inertia_diag']
    if I[0] > 0:
        out['link0_arm_y'] = float(w[0] * I[0] / (0.01 * dt))              # tau_x = y f
        out['link0_arm_x'] = float(-w[1] * I[1] / (0.01 * dt))
    # clamp
    s.set(0, (0, 0, 5))
    p.applyExternalForce(q, -1, [1e4, 0, 0], [0, 0, 0], p.LINK_FRAME, physicsClientId=s.cid)
    s.step()
    out['max_coord_vel'] = float(s.get(0)[2][0])
    # rest on the ground
    s.set(0, (0, 0, 0.6))
    s.step(400)
    pos, quat, v, w = s.get(0)
    out['rest'] = dict(height=float(pos[2]), speed=float(np.abs(v).max()), spin=float(np.abs(w).max()))
    # tilted landing
    s.set(0, (0, 0, 0.62), quat=quat_from_euler(0.4, 0.2, 0.1))
    s.step(300)
    pos, quat, v, w = s.get(0)
    R22 = 1 - 2 * (quat[0] ** 2 + quat[1] ** 2)
    out['tilted_landing'] = dict(height=float(pos[2]), body_z_dot_world_z=float(R22))
    s.close()
    # quad-quad push-out distance
    s = Scene(p, models, n_quads=2)
    s.set(0, (0.0, 0, 0.5135))
    s.set(1, (0.1, 0, 0.5135))
    s.step(600)
    out['pair_rest_distance'] = float(np.linalg.norm(s.get(0)[0] - s.get(1)[0]))
    s.close()
    # seeded one-step goldens under a body wrench
    rng = np.random.default_rng(seed)
    s = Scene(p, models)
    q = s.quads[0]
    gold = []
    for _ in range(8):
        pos = rng.uniform(-1, 1, 3) + np.array([0, 0, 3.0])
        quat = quat_from_euler(*rng.uniform(-0.3, 0.3, 3))
        vel, ang = rng.uniform(-1, 1, 3), rng.uniform(-2, 2, 3)
        F, T = rng.uniform(-0.1, 0.1, 3) + np.array([0, 0, 0.26]), rng.uniform(-1e-4, 1e-4, 3)
        s.set(0, pos, quat, vel, ang)
        p.applyExternalForce(q, 4, F.tolist(), [0, 0, 0], p.LINK_FRAME, physicsClientId=s.cid)
        p.applyExternalTorque(q, 4, T.tolist(), p.LINK_FRAME, physicsClientId=s.cid)
        s.step()
        p1, q1, v1, w1 = s.get(0)
        gold.append(dict(pos=pos.tolist(), quat=list(quat), vel=vel.tolist(), angvel=ang.tolist(), force_body=F.tolist(),
                         torque_body=T.tolist(), pos1=p1.tolist(), quat1=q1.tolist(), vel1=v1.tolist(), angvel1=w1.tolist()))
    s.close()
    out['one_step_goldens'] = gold
    return out


def physics_params(probes):
    """PhysicsParams fields the probes determine (the rest keep the oracle's defaults)."""
    from oracle import bullet_model as bm
    P = bm.PhysicsParams()
    d = {}
    di = probes['dynamics_info']
    d['mass'] = di['mass']
    d['lin_damping'] = round(probes['lin_damping_k'], 6)
    d['ang_damping'] = round(probes['ang_damping_k'], 6)
    d['max_coord_vel'] = round(probes['max_coord_vel'], 3)
    d['mu_ground'] = probes['plane_lateral_friction'] * di['lateral_friction']
    d['mu_agent'] = di['lateral_friction'] ** 2
    if di.get('collision_margin'):
        d['col_margin'] = di['collision_margin']
    eng = probes.get('engine', {})
    if 'numSolverIterations' in eng:
        d['solver_iters'] = int(eng['numSolverIterations'])
    if 'contactERP' in eng:
        d['erp2'] = eng['contactERP']
    if 'contactSlop' in eng:
        d['slop'] = eng['contactSlop']
    d['gyro'] = abs(probes['gyro_step_angvel'][1]) > 1e-9
    d['ground_z'] = probes['rest']['height'] - (P.col_halfheight + d.get('col_margin', P.col_margin)) + d.get('slop', P.slop)
    d['measured_inertia_diag'] = di['local_inertia_diag']           # compare with PhysicsParams.inertia_diag()
    d['contact_radius_equivalent'] = 0.5 * probes['pair_rest_distance']
    return d


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    ap.add_argument('--backend', default='auto', choices=['auto', 'pybullet', 'fake'])
    ap.add_argument('--models', default='', help='directory with cf2x.urdf / plane.urdf (default: the reference tree)')
    args = ap.parse_args(argv)
    p, name = backend(args.backend)
    probes = probe(p, models_dir(args.models))
    doc = {'backend': name, 'pybullet_version': getattr(p, '__version__', None) or (
        p.getAPIVersion() if hasattr(p, 'getAPIVersion') else None),
        'note': 'backend "fake" = the oracle reading its own constants back: NOT a pin' if name == 'fake' else
                'measured on a real pybullet: load with bm.PhysicsParams.from_json and regenerate tests/golden',
        'probes': probes, 'PhysicsParams': physics_params(probes)}
    text = json.dumps(doc, indent=1)
    if args.out:
        open(args.out, 'w').write(text + '\n')
    else:
        print(text)
    return doc


if __name__ == '__main__':
    main()
