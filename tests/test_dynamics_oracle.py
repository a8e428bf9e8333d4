"""The float64 dynamics oracle checks itself: ABA (Featherstone Table 7.1) against the independent
CRBA + RNEA route (M qdd + C = tau), physical invariants, and the one Tier-B fact the reference pins
(zero velocity, gravity and torque => state unchanged, SURVEY.md 8(c) C6 viii).  CPU only."""
import numpy as np

from oracle.dynamics_oracle import DynChain, DynConfig, aba, crba, dynamic_substeps, rnea, total_energy
from pioneer_b200.urdf import flatten_urdf

CH = DynChain.from_model(flatten_urdf())


def test_composite_bodies_from_the_urdf():
    # 11 unit-mass links with unit inertia folded into 6 moving frames (urdf:27-202): masses 2,1,1,2,1,3
    m = flatten_urdf()
    assert np.allclose(m.body_mass, [2, 1, 1, 2, 1, 3])
    # rotator1 + hinge1 share the frame origin: com 0, inertia 2*I; rotator3 + effector + pointer (offset 3.6,0,1.9)
    assert np.allclose(m.body_com[0], 0) and np.allclose(m.body_inertia[0], 2 * np.eye(3))
    assert np.allclose(m.body_com[5], np.array([3.6, 0, 1.9]) / 3)
    assert np.allclose(m.effort, 1.0) and np.allclose(m.damping, 0.0)


def test_aba_agrees_with_crba_plus_rnea():
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.uniform(CH.lower, CH.upper)
        qd = rng.normal(size=6)
        tau = rng.normal(size=6) * 5
        g = float(rng.choice([0.0, 9.81]))
        qdd = aba(CH, q, qd, tau, g)
        M = crba(CH, q)
        C = rnea(CH, q, qd, np.zeros(6), g)
        np.testing.assert_allclose(M @ qdd + C, tau, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(rnea(CH, q, qd, qdd, g), tau, rtol=1e-9, atol=1e-9)
        assert np.allclose(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)


def test_zero_input_leaves_state_unchanged_bit_for_bit():
    q0 = np.array([0.3, -0.4, 0.9, 1.1, -0.7, 2.0], np.float32).astype(np.float64)
    q, qd = dynamic_substeps(CH, DynConfig(), q0, np.zeros(6), np.zeros(6), CH.lower, CH.upper, n_sub=240)
    assert np.array_equal(q, q0) and np.array_equal(qd, np.zeros(6))


def test_energy_is_conserved_without_input():
    """Free swing under gravity, no torque, no damping, away from the limits: semi-implicit Euler keeps the
    total energy within O(dt) of its initial value."""
    cfg = DynConfig(gravity=9.81, frame_skip=1)
    q = np.array([0.2, 0.3, -0.2, 0.1, 0.2, -0.1])
    qd = np.zeros(6)
    wide = np.full(6, 1e9)
    e0 = total_energy(CH, q, qd, cfg.gravity)
    worst = 0.0
    for _ in range(240):
        q, qd = dynamic_substeps(CH, cfg, q, qd, np.zeros(6), -wide, wide)
        worst = max(worst, abs(total_energy(CH, q, qd, cfg.gravity) - e0))
    assert np.abs(qd).max() > 0.05                       # it really moved
    assert worst < 2e-2 * max(1.0, abs(e0)), (worst, e0)


def test_pd_control_settles_on_the_set_point():
    cfg = DynConfig(kp=2000.0, kd=500.0, torque_scale=1e5)
    target = np.array([0.5, -0.3, 0.4, 1.0, -0.5, 0.8])
    q, qd = np.zeros(6), np.zeros(6)
    for _ in range(600):
        q, qd = dynamic_substeps(CH, cfg, q, qd, target, CH.lower, CH.upper)
    np.testing.assert_allclose(q, target, atol=1e-3)
    assert np.abs(qd).max() < 1e-3


def test_limits_are_inelastic_stops():
    cfg = DynConfig(torque_scale=1000.0)                # pure torque control, large torque into the upper limits
    q, qd = np.zeros(6), np.zeros(6)
    lo, hi = CH.lower.astype(np.float32), CH.upper.astype(np.float32)
    for _ in range(200):
        q, qd = dynamic_substeps(CH, cfg, q, qd, np.full(6, 500.0), lo, hi)
    assert np.all(q <= hi.astype(np.float64)) and np.all(q >= lo.astype(np.float64))
    assert (q == hi.astype(np.float64)).any()
    assert np.all(qd[q == hi.astype(np.float64)] == 0)


# ---------------------------------------------------------------------------------------------------------------
# An independent formulation: Lagrange's equations from Cartesian Jacobians.  Nothing below uses spatial (6-D)
# algebra, Pluecker transforms or the recursive algorithms: only the forward kinematics of the link frames
# (child = parent * T(origin) * R(axis, q), the URDF convention) and each composite body's mass / centre of mass /
# inertia about the centre of mass.  M(q) = sum m Jv^T Jv + Jw^T (R Ic R^T) Jw;  the Coriolis / centrifugal vector
# comes from the Christoffel symbols of M (central differences);  g_a = dV/dq_a with V = sum m g z.
# ---------------------------------------------------------------------------------------------------------------
def _frames(ch, q):
    R, p = np.eye(3), np.zeros(3)
    out = []
    for i in range(6):
        p = p + R @ ch.origin_xyz[i]
        k = ch.axis[i] / np.linalg.norm(ch.axis[i])
        c, s = np.cos(q[i]), np.sin(q[i])
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = R @ ch.origin_rot[i] @ (np.eye(3) + s * K + (1 - c) * (K @ K))       # Rodrigues
        out.append((R.copy(), p.copy()))
    return out


def _mass_matrix_and_gravity(ch, q, gravity):
    fr = _frames(ch, q)
    M, g = np.zeros((6, 6)), np.zeros(6)
    for i in range(6):
        Ri, pi = fr[i]
        pc = pi + Ri @ ch.body_com[i]
        Jv, Jw = np.zeros((3, 6)), np.zeros((3, 6))
        for k in range(i + 1):
            Rk, pk = fr[k]
            z = Rk @ (ch.axis[k] / np.linalg.norm(ch.axis[k]))       # the joint axis is fixed in its own (moving) frame
            Jv[:, k] = np.cross(z, pc - pk)
            Jw[:, k] = z
        M += ch.body_mass[i] * Jv.T @ Jv + Jw.T @ (Ri @ ch.body_inertia[i] @ Ri.T) @ Jw
        g += ch.body_mass[i] * gravity * Jv[2, :]
    return M, g


def _lagrange_torque(ch, q, qd, qdd, gravity, h=1e-6):
    M, g = _mass_matrix_and_gravity(ch, q, gravity)
    dM = np.zeros((6, 6, 6))                                          # dM[a, b, c] = d M_ab / d q_c
    for c in range(6):
        e = np.zeros(6); e[c] = h
        dM[:, :, c] = (_mass_matrix_and_gravity(ch, q + e, 0.0)[0] - _mass_matrix_and_gravity(ch, q - e, 0.0)[0]) / (2 * h)
    cor = np.zeros(6)
    for a in range(6):
        for b in range(6):
            for c in range(6):
                cor[a] += 0.5 * (dM[a, b, c] + dM[a, c, b] - dM[b, c, a]) * qd[b] * qd[c]
    return M @ qdd + cor + g, M


def test_aba_satisfies_lagranges_equations():
    """tau = M(q) qdd + Gamma(q)[qd, qd] + dV/dq with qdd from the articulated-body algorithm: the recursive spatial
    algorithm and the energy-based formulation agree, on the shipped robot and on a chain with a tilted axis and a
    rotated joint origin."""
    import copy
    tilted = copy.deepcopy(CH)
    tilted.axis[2] = np.array([0.0, 0.6, 0.8])
    c, s = np.cos(0.4), np.sin(0.4)
    tilted.origin_rot[3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    tilted.body_com[1] = np.array([0.3, -0.2, 0.5])
    rng = np.random.default_rng(3)
    for ch in (CH, tilted):
        for _ in range(6):
            q = rng.uniform(ch.lower, ch.upper)
            qd = rng.normal(size=6)
            tau = rng.normal(size=6) * 50
            grav = float(rng.choice([0.0, 9.81]))
            qdd = aba(ch, q, qd, tau, grav)
            got, M = _lagrange_torque(ch, q, qd, qdd, grav)
            np.testing.assert_allclose(M, crba(ch, q), rtol=1e-10, atol=1e-9)
            scale = np.abs(tau).max() + np.abs(M @ qdd).max()
            assert np.abs(got - tau).max() <= 1e-6 * scale, (np.abs(got - tau).max(), scale)
