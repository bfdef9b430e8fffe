"""TEST INFRASTRUCTURE -- CPU oracle, not product code.

float64 numpy restatement of what ``p.stepSimulation()`` does to the
``'simple'`` mrs-gym scene (N x cf2x.urdf multibodies + plane.urdf), i.e. the
call at /root/reference/mrsgym/BulletSim.py:46-47.

The arithmetic lives in the third-party ``pybullet`` wheel (Bullet3 C++,
``btMultiBodyDynamicsWorld``).  It is NOT vendored under /root/reference and is
un-pinned (reference setup.py:5 ``install_requires=[..., 'pybullet', ...]``; the
shipped .pyc files are cpython-37 => pybullet 3.0-3.1 era).  pybullet is not
installable here (no network) so this file restates Bullet's published
algorithm from the bullet3 sources as recalled:

  btMultiBodyDynamicsWorld::solveConstraints         gravity -> ABA -> v += dt*a
  btMultiBody::computeAccelerationsArticulatedBody…  damping 0.04, gyro term
  btMultiBody::applyDeltaVeeMultiDof                 +-100 coordinate clamp
  btMultiBodyConstraintSolver::setupMultiBodyContactConstraint   erp2 / slop rhs
  btMultiBody::stepPositionsMultiDof                 p += dt*v, exp-map quaternion
  URDF2Bullet / btCompoundShape::calculateLocalInertia   AABB box inertia

PARITY UNPINNED for this file: the reference holds no tests / golden vectors
and real PyBullet cannot be run here.  Every constant is a field of
``PhysicsParams`` so a run with a real pybullet can correct it without code
changes.  Contact is a deliberate simplification named by the north star
(ground plane + AGENT_RADIUS sphere-sphere), specified here and mirrored
exactly by the CUDA kernels.

All arrays are ``[..., 3]`` / ``[..., 4]`` (quaternion order xyzw, as PyBullet)
with arbitrary leading batch dims; agent-agent contact couples the second to
last axis (``[..., N, 3]``).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass
class QuadParams:
    """cf2x.urdf constants (/root/reference/mrsgym/models/cf2x.urdf:5,11-12,31-36,42-78)
    and the derived ones of Quadcopter.calculate_parameters
    (/root/reference/mrsgym/Quadcopter.py:153-168)."""
    mass: float = 0.027
    arm: float = 0.0397
    kf: float = 3.16e-10
    km: float = 7.94e-12
    thrust2weight: float = 2.25
    ixx: float = 1.4e-5          # file inertia: used by set_control only
    iyy: float = 1.4e-5
    izz: float = 2.17e-5
    gnd_eff_coeff: float = 11.36859
    prop_radius: float = 2.31348e-2
    drag_xy: float = 9.1785e-7
    drag_z: float = 10.311e-7
    dw1: float = 2267.18
    dw2: float = 0.16
    dw3: float = -0.11
    col_radius: float = 0.06
    col_length: float = 0.025
    # prop link CoM offsets, cf2x.urdf:42,54,66,78
    prop_xy: tuple = ((0.028, 0.028), (-0.028, 0.028), (-0.028, -0.028), (0.028, -0.028))
    # QuadControl gains (/root/reference/mrsgym/QuadControl.py:14-32)
    pos_p: float = 1.5
    pos_i: float = 0.001
    pos_d: float = 1.0
    vel_p: float = 3.0
    vel_i: float = 0.1
    vel_d: float = 1.0
    ori_p: tuple = (70000.0, 70000.0, 60000.0)
    ori_i: tuple = (0.0, 0.0, 500.0)
    ori_d: tuple = (20000.0, 20000.0, 12000.0)
    min_pwm: float = 20000.0
    max_pwm: float = 65535.0
    pwm2rpm_a: float = 0.2685
    pwm2rpm_b: float = 4070.3
    ctrl_dt: float = 0.01        # QuadControl always uses DefaultSim (QuadControl.py:10)
    ctrl_gravity: float = 9.81

    def derived(self, gravity: float = 9.81) -> dict:
        g = gravity * self.mass
        max_rpm = math.sqrt(self.thrust2weight * g / (4 * self.kf))
        max_thrust = 4.0 * self.kf * max_rpm ** 2
        hclip = 0.25 * self.prop_radius * math.sqrt(
            (15 * max_rpm ** 2 * self.kf * self.gnd_eff_coeff) / max_thrust)
        return dict(GravityForce=g, HoverRPM=math.sqrt(g / (4 * self.kf)), MaxRPM=max_rpm,
                    MaxThrust=max_thrust, GroundEffectHClip=hclip)


@dataclasses.dataclass
class PhysicsParams:
    """Bullet-side constants.  [Bullet, unverified] unless noted."""
    mass: float = 0.027
    # collision AABB half extents used for the default (no URDF_USE_INERTIA_FROM_FILE)
    # inertia: cylinder hull r=.06, half-length .0125; margins: hull recalcLocalAabb
    # (+.001) + btTransformAabb (+.001) + compound getAabb (+.001) => 3 margins.
    inertia_margins: int = 3
    col_margin: float = 0.001
    col_radius: float = 0.06
    col_halfheight: float = 0.0125
    lin_damping: float = 0.04
    ang_damping: float = 0.04
    max_coord_vel: float = 100.0
    gyro: bool = True
    ang_motion_threshold: float = 0.25 * math.pi
    # contact solver (PhysicsServerCommandProcessor defaults)
    erp2: float = 0.08
    slop: float = 1e-5
    contact_margin: float = 0.02     # contact breaking threshold: speculative contacts
    mu_ground: float = 0.75          # plane 1.5 (plane.urdf:4-6) x link default 0.5
    ground_z: float = 0.5            # plane.urdf:21-26 box 30x30x1 centred at z=0
    ground_contact: bool = True
    agent_contact: bool = True
    agent_radius: float = 0.3        # MRS.AGENT_RADIUS (/root/reference/mrsgym/MRS.py:28)

    def inertia_diag(self) -> np.ndarray:
        hx = self.col_radius + self.inertia_margins * self.col_margin
        hz = self.col_halfheight + self.inertia_margins * self.col_margin
        lx, lz = 2 * hx, 2 * hz
        return np.array([self.mass / 12.0 * (lx * lx + lz * lz),
                         self.mass / 12.0 * (lx * lx + lz * lz),
                         self.mass / 12.0 * (lx * lx + lx * lx)])


# ----------------------------------------------------------------------------- rotations
def quat_to_mat(q: np.ndarray) -> np.ndarray:
    """xyzw unit quaternion -> body->world rotation matrix [..., 3, 3]."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3), dtype=q.dtype)
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product a (x) b, xyzw."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _norm(v):
    return np.sqrt(np.sum(v * v, axis=-1, keepdims=True))


def matvec(R, v):
    return np.einsum('...ij,...j->...i', R, v)


def matTvec(R, v):
    return np.einsum('...ji,...j->...i', R, v)


# ----------------------------------------------------------------------------- step pieces
def unconstrained_velocities(quat, vel, angvel, force_w, torque_w, P: PhysicsParams,
                             dt: float, gravity: float):
    """Bullet solveConstraints front half: gravity, ABA for a free base with damping
    and gyroscopic term, v += dt*a, +-max_coord_vel clamp per coordinate."""
    I = P.inertia_diag()
    R = quat_to_mat(quat)
    F = force_w.copy()
    F[..., 2] -= P.mass * gravity
    k = P.lin_damping
    a = F / P.mass - vel * (k + k * _norm(vel))
    w_b = matTvec(R, angvel)
    tau_b = matTvec(R, torque_w)
    Iw = I * w_b
    ka = P.ang_damping
    rhs = tau_b - Iw * (ka + ka * _norm(w_b))
    if P.gyro:
        rhs = rhs - np.cross(w_b, Iw)
    wdot_w = matvec(R, rhs / I)
    v1 = np.clip(vel + dt * a, -P.max_coord_vel, P.max_coord_vel)
    w1 = np.clip(angvel + dt * wdot_w, -P.max_coord_vel, P.max_coord_vel)
    return v1, w1


def _contact_rhs(dist, vn, P: PhysicsParams, dt: float):
    """Target normal-velocity change of one contact row, after
    btMultiBodyConstraintSolver::setupMultiBodyContactConstraint (restitution 0):
    penetration = dist + slop; open gap -> may close it within the step,
    penetrating -> Baumgarte push-out erp2*depth/dt.  Lower impulse limit 0."""
    pen = dist + P.slop
    rhs = np.where(pen > 0.0, -vn - pen / dt, -vn - pen * P.erp2 / dt)
    return np.maximum(rhs, 0.0)


def agent_contact_dv(pos, v1, P: PhysicsParams, dt: float):
    """AGENT_RADIUS sphere-sphere contact (north star simplification of the quad-quad
    hull contact).  One Jacobi pass over all pairs from the unconstrained velocities:
    frictionless central impulse, equal masses => each body takes half of the row."""
    N = pos.shape[-2]
    dp = pos[..., :, None, :] - pos[..., None, :, :]            # p_i - p_j
    d = np.sqrt(np.sum(dp * dp, axis=-1))
    dist = d - 2.0 * P.agent_radius
    with np.errstate(invalid='ignore', divide='ignore'):
        n = dp / d[..., None]
    dv = v1[..., :, None, :] - v1[..., None, :, :]
    vn = np.sum(dv * n, axis=-1)
    active = (dist < P.contact_margin) & (d > 0.0) & ~np.eye(N, dtype=bool)
    rhs = np.where(active, _contact_rhs(dist, np.where(active, vn, 0.0), P, dt), 0.0)
    n = np.where(active[..., None], n, 0.0)
    return 0.5 * np.sum(rhs[..., None] * n, axis=-2)


def ground_contact(pos, quat, v, P: PhysicsParams, dt: float):
    """Ground plane z = ground_z against the quad's collision cylinder (support extent
    along world z, + one collision margin); impulse acts at the CoM (no torque);
    isotropic Coulomb friction mu_ground on the tangential velocity."""
    R22 = quat_to_mat(quat)[..., 2, 2]
    ext = (P.col_radius * np.sqrt(np.maximum(1.0 - R22 * R22, 0.0))
           + P.col_halfheight * np.abs(R22) + P.col_margin)
    dist = pos[..., 2] - ext - P.ground_z
    active = dist < P.contact_margin
    vn = v[..., 2]
    jn = np.where(active, _contact_rhs(dist, vn, P, dt), 0.0)    # per unit mass
    vt = v[..., :2]
    vt_n = np.sqrt(np.sum(vt * vt, axis=-1))
    with np.errstate(invalid='ignore', divide='ignore'):
        scale = np.where(vt_n > 0.0, np.minimum(vt_n, P.mu_ground * jn) / vt_n, 0.0)
    out = v.copy()
    out[..., 2] = vn + jn
    out[..., :2] = vt - vt * scale[..., None]
    return out


def integrate_positions(pos, quat, v, w, P: PhysicsParams, dt: float):
    """btMultiBody::stepPositionsMultiDof for the base: p += dt*v; world-frame
    exponential map q <- dq(w*dt) (x) q with |w|dt capped at pi/4 and a Taylor branch
    below 1e-3 rad/s, then normalise."""
    pos1 = pos + dt * v
    ang = _norm(w)
    ang = np.where(ang * dt > P.ang_motion_threshold, 0.5 * (0.5 * math.pi) / dt, ang)
    with np.errstate(invalid='ignore', divide='ignore'):
        big = np.sin(0.5 * ang * dt) / ang
    small = 0.5 * dt - (dt * dt * dt) * 0.020833333333 * ang * ang
    axis = w * np.where(ang < 0.001, small, big)
    dq = np.concatenate([axis, np.cos(0.5 * ang * dt)], axis=-1)
    q1 = quat_mul(dq, quat)
    q1 = q1 / _norm(q1)
    return pos1, q1


def bullet_step(pos, quat, vel, angvel, force_w, torque_w, P: PhysicsParams,
                dt: float = 0.01, gravity: float = 9.81):
    """One ``stepSimulation`` (numSubSteps=0).  Inputs float64, world frame; force/torque
    are the summed external wrench about the CoM.  Returns (pos, quat, vel, angvel)."""
    v1, w1 = unconstrained_velocities(quat, vel, angvel, force_w, torque_w, P, dt, gravity)
    if P.agent_contact and pos.shape[-2] > 1:
        v1 = v1 + agent_contact_dv(pos, v1, P, dt)
    if P.ground_contact:
        v1 = ground_contact(pos, quat, v1, P, dt)
    pos1, q1 = integrate_positions(pos, quat, v1, w1, P, dt)
    return pos1, q1, v1, w1
