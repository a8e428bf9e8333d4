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
