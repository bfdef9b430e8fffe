"""TEST INFRASTRUCTURE -- stand-in for the `pybullet` module (not installable here).

Exports the 13 backend functions the mrs-gym hot path calls (seam 2)
on top of ``oracle/bullet_model.py`` so that /root/reference/mrsgym runs VERBATIM:

  connect/setGravity/setTimeStep/setRealTimeSimulation   mrsgym/BulletSim.py:29-35
  loadURDF                                               mrsgym/EnvCreator.py:60
  resetBasePositionAndOrientation/resetBaseVelocity      mrsgym/Object.py:64-65
  getBasePositionAndOrientation/getBaseVelocity          mrsgym/Object.py:79-91
  getLinkStates                                          mrsgym/Quadcopter.py:71
  applyExternalForce/applyExternalTorque                 mrsgym/Quadcopter.py:44-45,82,93,110
  stepSimulation                                         mrsgym/BulletSim.py:47

State is float64 (PyBullet wheels are BT_USE_DOUBLE_PRECISION); getters hand back tuples
of python floats so ``torch.tensor(...)`` yields float32 exactly as with the real module.
PARITY UNPINNED for the integrator/contact part (see bullet_model.py header).
"""
from __future__ import annotations

import os

import numpy as np

from . import bullet_model as bm

DIRECT, GUI = 2, 1
LINK_FRAME, WORLD_FRAME = 1, 2
COV_ENABLE_GUI = 1

# knobs the golden generator may override before `connect`
PHYSICS = bm.PhysicsParams()
QUAD = bm.QuadParams()
RECORD = None  # optional list: per stepSimulation dict of applied wrench (debug/golden)


class _World:
    def __init__(self):
        self.gravity = 9.81
        self.dt = 1.0 / 240.0
        self.kind = []      # 'quad' | 'static'
        self.pos, self.quat, self.vel, self.angvel = [], [], [], []
        self.force, self.torque = [], []
        self.params = bm.PhysicsParams(**vars(PHYSICS))


_WORLDS = {}
_NEXT_ID = [0]


def _w(physicsClientId=0):
    return _WORLDS[physicsClientId]


def connect(mode, **kw):
    cid = _NEXT_ID[0]     # never reuse ids: a stale MRS.__del__ must not hit a newer world
    _NEXT_ID[0] += 1
    _WORLDS[cid] = _World()
    return cid


def disconnect(physicsClientId=0):
    _WORLDS.pop(physicsClientId, None)


def resetSimulation(physicsClientId=0):
    if physicsClientId in _WORLDS:
        _WORLDS[physicsClientId] = _World()


def configureDebugVisualizer(*a, **k):
    pass


def addUserDebugLine(*a, **k):
    return 0


def removeUserDebugItem(*a, **k):
    pass


def removeBody(*a, **k):
    pass


def setGravity(gravX=0, gravY=0, gravZ=-9.81, physicsClientId=0):
    _w(physicsClientId).gravity = -float(gravZ)


def setTimeStep(timeStep, physicsClientId=0):
    _w(physicsClientId).dt = float(timeStep)


def setRealTimeSimulation(enable, physicsClientId=0):
    assert not enable, "oracle backend is DIRECT only"


def loadURDF(fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), physicsClientId=0, **kw):
    w = _w(physicsClientId)
    name = os.path.basename(fileName)
    uid = len(w.kind)
    if name.startswith('plane'):
        w.kind.append('static')
        # box 30x30x1 centred on the base (plane.urdf:21-26) => top = base z + 0.5
        w.params.ground_z = float(basePosition[2]) + 0.5
    else:
        w.kind.append('quad')
    w.pos.append(np.array(basePosition, dtype=np.float64))
    w.quat.append(np.array(baseOrientation, dtype=np.float64))
    w.vel.append(np.zeros(3))
    w.angvel.append(np.zeros(3))
    w.force.append(np.zeros(3))
    w.torque.append(np.zeros(3))
    return uid


def resetBasePositionAndOrientation(bodyUniqueId, posObj, ornObj, physicsClientId=0):
    w = _w(physicsClientId)
    w.pos[bodyUniqueId] = np.array(posObj, dtype=np.float64)
    q = np.array(ornObj, dtype=np.float64)
    w.quat[bodyUniqueId] = q / np.linalg.norm(q)


def resetBaseVelocity(objectUniqueId, linearVelocity=None, angularVelocity=None, physicsClientId=0):
    w = _w(physicsClientId)
    if linearVelocity is not None:
        w.vel[objectUniqueId] = np.array(linearVelocity, dtype=np.float64)
    if angularVelocity is not None:
        w.angvel[objectUniqueId] = np.array(angularVelocity, dtype=np.float64)


def _tup(a):
    return tuple(float(x) for x in a)


def getBasePositionAndOrientation(bodyUniqueId, physicsClientId=0):
    w = _w(physicsClientId)
    return _tup(w.pos[bodyUniqueId]), _tup(w.quat[bodyUniqueId])


def getBaseVelocity(bodyUniqueId, physicsClientId=0):
    w = _w(physicsClientId)
    return _tup(w.vel[bodyUniqueId]), _tup(w.angvel[bodyUniqueId])


def _link_offset(link):
    if 0 <= link < 4:
        x, y = QUAD.prop_xy[link]
        return np.array([x, y, 0.0])
    return np.zeros(3)


def getLinkStates(bodyUniqueId, linkIndices, computeLinkVelocity=0, computeForwardKinematics=0,
                  physicsClientId=0):
    w = _w(physicsClientId)
    R = bm.quat_to_mat(w.quat[bodyUniqueId])
    out = np.empty((len(linkIndices), 8), dtype=object)   # object array: survives numpy>=1.24
    for r, li in enumerate(linkIndices):
        p = w.pos[bodyUniqueId] + R @ _link_offset(li)
        out[r, 0] = _tup(p)
        out[r, 1] = _tup(w.quat[bodyUniqueId])
        out[r, 2] = (0.0, 0.0, 0.0)
        out[r, 3] = (0.0, 0.0, 0.0, 1.0)
        out[r, 4] = _tup(p)
        out[r, 5] = _tup(w.quat[bodyUniqueId])
        out[r, 6] = _tup(w.vel[bodyUniqueId])
        out[r, 7] = _tup(w.angvel[bodyUniqueId])
    return out


def applyExternalForce(objectUniqueId, linkIndex, forceObj, posObj, flags, physicsClientId=0):
    """LINK_FRAME: force and position are in the link's inertial frame (PhysicsServer-
    CommandProcessor: forceWorld = linkBasis*f, torque = (linkBasis*pos) x forceWorld about
    the link CoM, which sits at R*offset from the base CoM through the fixed joint)."""
    w = _w(physicsClientId)
    f = np.array([float(x) for x in forceObj], dtype=np.float64)
    r = np.array([float(x) for x in posObj], dtype=np.float64)
    R = bm.quat_to_mat(w.quat[objectUniqueId])
    if flags == LINK_FRAME:
        fw = R @ f
        arm = R @ (_link_offset(linkIndex) + r)
    else:
        fw = f
        arm = r - w.pos[objectUniqueId]
    w.force[objectUniqueId] = w.force[objectUniqueId] + fw
    w.torque[objectUniqueId] = w.torque[objectUniqueId] + np.cross(arm, fw)


def applyExternalTorque(objectUniqueId, linkIndex, torqueObj, flags, physicsClientId=0):
    w = _w(physicsClientId)
    t = np.array([float(x) for x in torqueObj], dtype=np.float64)
    if flags == LINK_FRAME:
        t = bm.quat_to_mat(w.quat[objectUniqueId]) @ t
    w.torque[objectUniqueId] = w.torque[objectUniqueId] + t


def stepSimulation(physicsClientId=0):
    w = _w(physicsClientId)
    idx = [i for i, k in enumerate(w.kind) if k == 'quad']
    if not idx:
        return
    pos = np.stack([w.pos[i] for i in idx])
    quat = np.stack([w.quat[i] for i in idx])
    vel = np.stack([w.vel[i] for i in idx])
    ang = np.stack([w.angvel[i] for i in idx])
    F = np.stack([w.force[i] for i in idx])
    T = np.stack([w.torque[i] for i in idx])
    if RECORD is not None:
        RECORD.append(dict(force=F.copy(), torque=T.copy()))
    p1, q1, v1, w1 = bm.bullet_step(pos, quat, vel, ang, F, T, w.params, w.dt, w.gravity)
    for r, i in enumerate(idx):
        w.pos[i], w.quat[i], w.vel[i], w.angvel[i] = p1[r], q1[r], v1[r], w1[r]
        w.force[i] = np.zeros(3)
        w.torque[i] = np.zeros(3)


# ----------------------------------------------------------------------------- introspection (tools/pin_bullet.py)
def getDynamicsInfo(bodyUniqueId, linkIndex, physicsClientId=0):
    """(mass, lateral friction, local inertia diagonal, inertial pos, inertial orn, restitution, rolling friction,
    spinning friction, contact damping, contact stiffness, body type, collision margin) as pybullet returns it."""
    w = _w(physicsClientId)
    P = w.params
    if w.kind[bodyUniqueId] == 'static':
        return (0.0, 1.5, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), 0.0, 0.0, 0.0, -1.0, -1.0, 1, 0.001)
    return (P.mass, 0.5, tuple(P.inertia_diag()), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), 0.0, 0.0, 0.0, -1.0, -1.0, 1,
            P.col_margin)


def getPhysicsEngineParameters(physicsClientId=0):
    w = _w(physicsClientId)
    P = w.params
    return dict(fixedTimeStep=w.dt, numSubSteps=0, numSolverIterations=int(P.solver_iters), useRealTimeSimulation=0,
                gravityAccelerationX=0.0, gravityAccelerationY=0.0, gravityAccelerationZ=-w.gravity,
                erp=0.2, contactERP=P.erp2, frictionERP=0.2, contactSlop=P.slop,
                solverResidualThreshold=1e-7, numNonContactInnerIterations=1)
