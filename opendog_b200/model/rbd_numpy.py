"""Generic (any tree) rigid-body kinematics / mass matrix in numpy, MuJoCo conventions.

Used (1) by the model compiler for the compile-time constants MuJoCo's
`mj_setConst` derives at `qpos0` (`dof_invweight0`, `body_invweight0`
[3P-recalled: engine_setconst.c `set0`]), and (2) by the tests as an
implementation of M(q) and the bias force that is independent of the
structured CRBA/RNE in `oracle/odg_oracle.c`.

Conventions: free joint qpos = (x,y,z, qw,qx,qy,qz); qvel = (world linear
velocity of the body origin, angular velocity in the BODY frame); hinge angle
is measured from `ref`.
"""
from __future__ import annotations

import numpy as np

from .mjcf import MjcfModel, axis_angle_quat, quat_mul, quat_normalize, quat_to_mat


def qpos0(model: MjcfModel) -> np.ndarray:
    q = np.zeros(model.nq)
    for j in model.joints:
        if j.type == "free":
            b = model.bodies[j.body]
            q[j.qposadr:j.qposadr + 3] = b.pos
            q[j.qposadr + 3:j.qposadr + 7] = b.quat
        else:
            q[j.qposadr] = j.ref
    return q


def kinematics(model: MjcfModel, qpos: np.ndarray):
    """Body poses. Returns dict(xpos[nb,3], xmat[nb,3,3], xipos[nb,3], anchor[nj,3], axis[nj,3])."""
    nb = len(model.bodies)
    xpos = np.zeros((nb, 3)); xquat = np.zeros((nb, 4)); xquat[0, 0] = 1
    anchor = np.zeros((len(model.joints), 3)); axis = np.zeros((len(model.joints), 3))
    for i in range(1, nb):
        b = model.bodies[i]
        p = b.parent
        Rp = quat_to_mat(xquat[p])
        pos = xpos[p] + Rp @ b.pos
        quat = quat_mul(xquat[p], b.quat)
        for jid in b.joints:
            j = model.joints[jid]
            if j.type == "free":
                pos = qpos[j.qposadr:j.qposadr + 3].copy()
                quat = quat_normalize(qpos[j.qposadr + 3:j.qposadr + 7])
                anchor[jid] = pos
            else:
                R = quat_to_mat(quat)
                anchor[jid] = pos + R @ j.pos
                axis[jid] = R @ j.axis
                quat = quat_mul(quat, axis_angle_quat(j.axis, qpos[j.qposadr] - j.ref))
                pos = anchor[jid] - quat_to_mat(quat) @ j.pos
        xpos[i], xquat[i] = pos, quat_normalize(quat)
    xmat = np.stack([quat_to_mat(q) for q in xquat])
    xipos = np.stack([xpos[i] + xmat[i] @ model.bodies[i].ipos for i in range(nb)])
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, anchor=anchor, axis=axis)


def jacobian(model: MjcfModel, kin, body: int, point: np.ndarray):
    """(Jp[3,nv], Jr[3,nv]) of a world `point` rigidly attached to `body`."""
    Jp = np.zeros((3, model.nv)); Jr = np.zeros((3, model.nv))
    b = body
    while b > 0:
        for jid in model.bodies[b].joints:
            j = model.joints[jid]
            d = j.dofadr
            if j.type == "free":
                Jp[:, d:d + 3] = np.eye(3)
                R = kin["xmat"][b]
                for k in range(3):
                    Jr[:, d + 3 + k] = R[:, k]
                    Jp[:, d + 3 + k] = np.cross(R[:, k], point - kin["xpos"][b])
            else:
                Jr[:, d] = kin["axis"][jid]
                Jp[:, d] = np.cross(kin["axis"][jid], point - kin["anchor"][jid])
        b = model.bodies[b].parent
    return Jp, Jr


def mass_matrix(model: MjcfModel, qpos: np.ndarray) -> np.ndarray:
    kin = kinematics(model, qpos)
    M = np.zeros((model.nv, model.nv))
    for i, b in enumerate(model.bodies):
        if i == 0 or b.mass == 0:
            continue
        Jp, Jr = jacobian(model, kin, i, kin["xipos"][i])
        Iw = kin["xmat"][i] @ b.inertia @ kin["xmat"][i].T
        M += b.mass * Jp.T @ Jp + Jr.T @ Iw @ Jr
    for j in model.joints:
        n = 6 if j.type == "free" else 1
        M[range(j.dofadr, j.dofadr + n), range(j.dofadr, j.dofadr + n)] += j.armature
    return M


def integrate_pos(model: MjcfModel, qpos: np.ndarray, qvel: np.ndarray, h: float) -> np.ndarray:
    """mj_integratePos: q <- q (+) h*v."""
    q = qpos.copy()
    for j in model.joints:
        if j.type == "free":
            q[j.qposadr:j.qposadr + 3] += h * qvel[j.dofadr:j.dofadr + 3]
            w = qvel[j.dofadr + 3:j.dofadr + 6]
            n = np.linalg.norm(w)
            if n > 0:
                dq = axis_angle_quat(w / n, h * n)
                q[j.qposadr + 3:j.qposadr + 7] = quat_normalize(quat_mul(quat_normalize(q[j.qposadr + 3:j.qposadr + 7]), dq))
        else:
            q[j.qposadr] += h * qvel[j.dofadr]
    return q


def potential_energy(model: MjcfModel, qpos: np.ndarray) -> float:
    kin = kinematics(model, qpos)
    g = model.option["gravity"]
    return -sum(b.mass * np.dot(g, kin["xipos"][i]) for i, b in enumerate(model.bodies) if i)


def bias_force(model: MjcfModel, qpos: np.ndarray, qvel: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """c(q,v) (Coriolis + centrifugal + gravity) from Lagrange's equations by finite differences.

    For quasi-velocities v with q' = q (+) v:  c = Ṁ v − ∂T/∂q + ∂V/∂q  plus the
    non-holonomic correction of the body-frame angular velocity (ω × (I ω)-type
    terms), obtained here numerically as   c = d/dt(M v)|_{v̇=0} − ∂/∂q(½vᵀMv) + ∂V/∂q
    where ∂/∂q is taken along the dof directions with v transported consistently
    (Hamel/Boltzmann form). To stay simple and exact we use the momentum form:
        c_k = d/dt[(M v)_k] − ∂T/∂π_k + Σ_{ij} γ terms
    which for this joint set reduces to transporting v through the body-frame
    rotation. The implementation below evaluates the generalized inertial force
    Σ_b Jp_bᵀ m a_b + Jr_bᵀ (I α + ω×Iω) with accelerations from second
    differences of the kinematics along the flow q(t) of constant v.
    """
    nv = model.nv
    h = eps ** 0.5 * 1e-1

    def flow(t):
        return integrate_pos(model, qpos, qvel, t)

    k0 = kinematics(model, qpos)
    kp = kinematics(model, flow(h)); km = kinematics(model, flow(-h))
    g = model.option["gravity"]
    c = np.zeros(nv)
    for i, b in enumerate(model.bodies):
        if i == 0 or b.mass == 0:
            continue
        acc = (kp["xipos"][i] - 2 * k0["xipos"][i] + km["xipos"][i]) / h ** 2
        Jp, Jr = jacobian(model, k0, i, k0["xipos"][i])
        # angular velocity (world) at t=±h/2 from rotation increments -> alpha by differencing
        def omega(Ra, Rb):
            dR = Rb @ Ra.T
            return np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (2 * h)
        w_m = omega(km["xmat"][i], k0["xmat"][i]); w_p = omega(k0["xmat"][i], kp["xmat"][i])
        alpha = (w_p - w_m) / h
        w = Jr @ qvel
        Iw = k0["xmat"][i] @ b.inertia @ k0["xmat"][i].T
        c += Jp.T @ (b.mass * (acc - g)) + Jr.T @ (Iw @ alpha + np.cross(w, Iw @ w))
    return c


def invweight0(model: MjcfModel):
    """(dof_invweight0[nv], body_invweight0[nb,2]) as `mj_setConst` computes them at qpos0."""
    q0 = qpos0(model)
    M = mass_matrix(model, q0)
    Minv = np.linalg.inv(M)
    kin = kinematics(model, q0)
    dof_w = np.zeros(model.nv)
    for j in model.joints:
        d = j.dofadr
        if j.type == "free":
            dof_w[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            dof_w[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            dof_w[d] = Minv[d, d]
    body_w = np.zeros((len(model.bodies), 2))
    for i in range(1, len(model.bodies)):
        Jp, Jr = jacobian(model, kin, i, kin["xipos"][i])
        if not Jp.any() and not Jr.any():
            continue
        body_w[i, 0] = np.trace(Jp @ Minv @ Jp.T) / 3
        body_w[i, 1] = np.trace(Jr @ Minv @ Jr.T) / 3
    return dof_w, body_w
