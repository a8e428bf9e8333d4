"""CPU ORACLE for the dynamic (Tier-B) mode — test infrastructure, NOT product code.

PARITY UNPINNED vs PyBullet: the reference env never exercises Bullet's dynamics (gravity 0, joints teleported with
zero velocity, no motor targets: SURVEY.md section 0 facts 2-3), PyBullet is not installable here, and Bullet's
POSITION_CONTROL is a velocity-level PGS constraint, not an explicit PD law.  This file therefore defines the
semantics of ``PNR_MODE_DYNAMIC`` and checks the CUDA kernels against an INDEPENDENT float64
implementation: textbook 6x6 spatial-matrix Featherstone ABA (R. Featherstone, "Rigid Body Dynamics Algorithms",
2008, Table 7.1), itself cross-checked against the composite-rigid-body mass matrix + recursive Newton-Euler
(M(q) qdd + C(q, qd) = tau, Tables 6.2 and 5.1) in tests/test_dynamics_oracle.py.  The one check the reference
does pin (SURVEY 8(c) C6 viii): qd = 0, g = 0, tau = 0 leaves (q, qd) bit-unchanged.

Semantics of one env step (``frame_skip`` substeps of ``timestep``; reference constants bullet_env.py:38-41):
    tau_i  = clamp(kp (u_i - q_i) - kd qd_i, +-effort_i * torque_scale)      if kp != 0 or kd != 0   (PD position control,
             clamp(u_i, +-effort_i * torque_scale)                            otherwise                u = action)
    tau_i -= damping_i qd_i                                                   (URDF <dynamics damping>)
    qdd    = ABA(q, qd, tau, gravity along -z)
    qd    += qdd dt ; q += qd dt                                              (semi-implicit Euler, btMultiBody order)
    q clamped to the joint limits; a joint driven into a limit loses its velocity (inelastic stop)
Reward / done / observation are the kinematic env's (pioneer_knm_env.py:151-211) evaluated after the substeps, with
r = q, v = qd, a = the action just applied.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

f32, f64 = np.float32, np.float64
DOF = 6


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=f64)


def rot_axis(axis, q):
    k = np.asarray(axis, f64)
    c, s = math.cos(q), math.sin(q)
    return np.eye(3) * c + s * skew(k) + (1 - c) * np.outer(k, k)


def x_motion(E, p):
    """6x6 Pluecker motion transform parent -> child; E = child-from-parent rotation, p = child origin in parent."""
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    X[3:, :3] = -E @ skew(p)
    return X


def crm(v):
    X = np.zeros((6, 6))
    X[:3, :3] = skew(v[:3])
    X[3:, 3:] = skew(v[:3])
    X[3:, :3] = skew(v[3:])
    return X


def crf(v):
    return -crm(v).T


def spatial_inertia(m, c, Ic):
    I = np.zeros((6, 6))
    C = skew(c)
    I[:3, :3] = Ic + m * (C @ C.T)
    I[:3, 3:] = m * C
    I[3:, :3] = m * C.T
    I[3:, 3:] = m * np.eye(3)
    return I


@dataclass
class DynChain:
    axis: np.ndarray
    origin_xyz: np.ndarray
    origin_rot: np.ndarray
    body_mass: np.ndarray
    body_com: np.ndarray
    body_inertia: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    effort: np.ndarray
    damping: np.ndarray

    @staticmethod
    def from_model(m) -> "DynChain":
        g = lambda a: np.array(a, f64)
        return DynChain(g(m.axis), g(m.origin_xyz), g(m.origin_rot), g(m.body_mass), g(m.body_com),
                        g(m.body_inertia), g(m.lower), g(m.upper), g(m.effort), g(m.damping))

    def S(self, i):
        return np.concatenate([self.axis[i], np.zeros(3)])

    def Xup(self, i, qi):
        R = self.origin_rot[i] @ rot_axis(self.axis[i], qi)        # child axes in parent coordinates
        return x_motion(R.T, self.origin_xyz[i])

    def I(self, i):
        return spatial_inertia(self.body_mass[i], self.body_com[i], self.body_inertia[i])


def aba(ch: DynChain, q, qd, tau, gravity: float) -> np.ndarray:
    """Featherstone articulated-body algorithm, fixed base, gravity along -z of the base frame."""
    n = DOF
    Xup, v, c, IA, pA = [None] * n, [None] * n, [None] * n, [None] * n, [None] * n
    for i in range(n):
        Xup[i] = ch.Xup(i, q[i])
        vJ = ch.S(i) * qd[i]
        v[i] = vJ if i == 0 else Xup[i] @ v[i - 1] + vJ
        c[i] = crm(v[i]) @ vJ
        IA[i] = ch.I(i)
        pA[i] = crf(v[i]) @ IA[i] @ v[i]
    U, d, u = [None] * n, np.zeros(n), np.zeros(n)
    for i in range(n - 1, -1, -1):
        S = ch.S(i)
        U[i] = IA[i] @ S
        d[i] = S @ U[i]
        u[i] = tau[i] - S @ pA[i]
        if i > 0:
            Ia = IA[i] - np.outer(U[i], U[i]) / d[i]
            pa = pA[i] + Ia @ c[i] + U[i] * (u[i] / d[i])
            IA[i - 1] = IA[i - 1] + Xup[i].T @ Ia @ Xup[i]
            pA[i - 1] = pA[i - 1] + Xup[i].T @ pa
    a_prev = np.array([0, 0, 0, 0, 0, gravity], f64)                # -a_gravity: the base "accelerates upwards"
    qdd = np.zeros(n)
    for i in range(n):
        a = Xup[i] @ a_prev + c[i]
        qdd[i] = (u[i] - U[i] @ a) / d[i]
        a_prev = a + ch.S(i) * qdd[i]
    return qdd


def rnea(ch: DynChain, q, qd, qdd, gravity: float) -> np.ndarray:
    """Recursive Newton-Euler inverse dynamics (independent of aba(): used to cross-check it)."""
    n = DOF
    Xup, v, a, f = [None] * n, [None] * n, [None] * n, [None] * n
    a_base = np.array([0, 0, 0, 0, 0, gravity], f64)
    for i in range(n):
        Xup[i] = ch.Xup(i, q[i])
        S = ch.S(i)
        vJ = S * qd[i]
        vp = np.zeros(6) if i == 0 else v[i - 1]
        ap = a_base if i == 0 else a[i - 1]
        v[i] = Xup[i] @ vp + vJ
        a[i] = Xup[i] @ ap + S * qdd[i] + crm(v[i]) @ vJ
        f[i] = ch.I(i) @ a[i] + crf(v[i]) @ ch.I(i) @ v[i]
    tau = np.zeros(n)
    for i in range(n - 1, -1, -1):
        tau[i] = ch.S(i) @ f[i]
        if i > 0:
            f[i - 1] = f[i - 1] + Xup[i].T @ f[i]
    return tau


def crba(ch: DynChain, q) -> np.ndarray:
    """Joint-space inertia matrix by the composite-rigid-body algorithm."""
    n = DOF
    Xup = [ch.Xup(i, q[i]) for i in range(n)]
    IC = [ch.I(i) for i in range(n)]
    for i in range(n - 1, 0, -1):
        IC[i - 1] = IC[i - 1] + Xup[i].T @ IC[i] @ Xup[i]
    M = np.zeros((n, n))
    for i in range(n):
        fh = IC[i] @ ch.S(i)
        M[i, i] = ch.S(i) @ fh
        j = i
        while j > 0:
            fh = Xup[j].T @ fh
            j -= 1
            M[i, j] = M[j, i] = ch.S(j) @ fh
    return M


@dataclass
class DynConfig:
    timestep: float = 1 / 240
    frame_skip: int = 10
    gravity: float = 0.0
    kp: float = 0.0
    kd: float = 0.0
    torque_scale: float = 1.0


def control_torque(ch: DynChain, cfg: DynConfig, action, q, qd) -> np.ndarray:
    lim = ch.effort * cfg.torque_scale
    if cfg.kp != 0.0 or cfg.kd != 0.0:
        tau = cfg.kp * (np.asarray(action, f64) - q) - cfg.kd * qd
    else:
        tau = np.asarray(action, f64).copy()
    tau = np.minimum(np.maximum(tau, -lim), lim)
    return tau - ch.damping * qd


def dynamic_substeps(ch: DynChain, cfg: DynConfig, q, qd, action, r_lo, r_hi, n_sub: Optional[int] = None):
    """``frame_skip`` substeps in float64 from (q, qd); limits are the float32 limits the env uses."""
    q, qd = np.array(q, f64), np.array(qd, f64)
    lo, hi = np.asarray(r_lo, f64), np.asarray(r_hi, f64)
    dt = cfg.timestep
    for _ in range(cfg.frame_skip if n_sub is None else n_sub):
        tau = control_torque(ch, cfg, action, q, qd)
        qdd = aba(ch, q, qd, tau, cfg.gravity)
        qd = qd + qdd * dt
        q = q + qd * dt
        over, under = q > hi, q < lo
        qd = np.where(over & (qd > 0), 0.0, qd)
        qd = np.where(under & (qd < 0), 0.0, qd)
        q = np.minimum(np.maximum(q, lo), hi)
    return q, qd


def total_energy(ch: DynChain, q, qd, gravity: float) -> float:
    """Kinetic + potential energy (conserved by the continuous dynamics when tau = 0, no damping, off the limits)."""
    M = crba(ch, q)
    ke = 0.5 * float(np.asarray(qd) @ M @ np.asarray(qd))
    # potential: sum m g z_com in the base frame
    R, p = np.eye(3), np.zeros(3)
    pe = 0.0
    for i in range(DOF):
        p = p + R @ ch.origin_xyz[i]
        R = R @ ch.origin_rot[i] @ rot_axis(ch.axis[i], q[i])
        pe += ch.body_mass[i] * gravity * float((p + R @ ch.body_com[i])[2])
    return ke + pe
