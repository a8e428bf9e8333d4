"""The C dynamic-mode oracle (oracle/dynamics_oracle.c) against the Python float64 6x6-matrix oracle it restates
(oracle/dynamics_oracle.py, itself cross-checked against CRBA + RNEA and Lagrange's equations in test_dynamics_oracle.py),
its env layer against the kinematic oracle's, and the exact segment-box distance of oracle/contact.h against brute force.
CPU only.  PARITY UNPINNED vs PyBullet (see the oracle headers)."""
import numpy as np
import pytest

from oracle.c_dyn_oracle import CDynOracleBatch, DynEnvConfig, segment_box_distance
from oracle.dynamics_oracle import DynChain, DynConfig, aba, dynamic_substeps
from oracle.reach_oracle import OracleChain, fk_pointer, reset_draws
from pioneer_b200.urdf import DEFAULT_URDF, flatten_urdf


def tilted_model(tmp_path):
    text = open(DEFAULT_URDF).read()
    text = text.replace('<origin xyz="0 0 11"/><axis xyz="0 1 0"/>', '<origin xyz="0 0 11" rpy="0.3 -0.2 0.5"/><axis xyz="0 0.6 0.8"/>')
    text = text.replace('<origin xyz="11 0 0"/><axis xyz="0 1 0"/>', '<origin xyz="11 0 0"/><axis xyz="0 -1 0"/>')
    path = tmp_path / "tilted.urdf"
    path.write_text(text)
    return flatten_urdf(str(path))


@pytest.mark.parametrize("robot", ["shipped", "tilted"])
def test_aba_equals_the_python_oracle(robot, tmp_path):
    model = flatten_urdf() if robot == "shipped" else tilted_model(tmp_path)
    ch = DynChain.from_model(model)
    rng = np.random.default_rng(3)
    for gravity in (0.0, 9.81):
        b = CDynOracleBatch(model, 1, DynEnvConfig(gravity=gravity))
        for _ in range(25):
            q = rng.uniform(model.lower, model.upper)
            qd = rng.normal(size=6) * 2.0
            tau = rng.normal(size=6) * 50.0
            want = aba(ch, q, qd, tau, gravity)
            got = b.aba(q, qd, tau)
            assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max()), (got, want)


@pytest.mark.parametrize("gravity,kp,kd,scale", [(0.0, 0.0, 0.0, 1.0), (9.81, 0.0, 0.0, 50.0), (9.81, 2000.0, 500.0, 1e5)])
def test_substeps_equal_the_python_oracle(gravity, kp, kd, scale):
    model = flatten_urdf()
    ch = DynChain.from_model(model)
    cfg = DynConfig(gravity=gravity, kp=kp, kd=kd, torque_scale=scale)
    b = CDynOracleBatch(model, 1, DynEnvConfig(gravity=gravity, kp=kp, kd=kd, torque_scale=scale))
    rng = np.random.default_rng(4)
    for k in range(12):
        q0 = rng.uniform(b.r_lo, b.r_hi) * (1.0 if k % 3 else 0.999999)       # some starts right at the stops
        qd0 = rng.normal(size=6) * (3.0 if k % 2 else 0.3)
        act = (rng.uniform(b.r_lo, b.r_hi) if kp else rng.normal(size=6) * scale).astype(np.float32)
        want_q, want_qd = dynamic_substeps(ch, cfg, q0, qd0, act, b.r_lo, b.r_hi)
        got_q, got_qd = b.substeps(q0, qd0, act, cfg.frame_skip)
        assert np.abs(got_q - want_q).max() <= 1e-10 and np.abs(got_qd - want_qd).max() <= 1e-8


def test_zero_input_is_a_bitwise_no_op():
    """World.step() in the reference (g = 0, qd = 0, tau = 0) changes nothing."""
    model = flatten_urdf()
    b = CDynOracleBatch(model, 64, DynEnvConfig(max_episode_steps=0), seed=3)
    before = b.state()["q"].copy()
    for _ in range(5):
        out = b.step(np.zeros((64, 6), np.float32))
    st = b.state()
    assert np.array_equal(st["q"], before) and not st["qd"].any() and (st["t"] == 5).all()


def test_env_layer_follows_the_kinematic_oracle():
    """Reset draws, reward, done, TimeLimit, statistics, auto-reset and both observation modes of the dynamic oracle are
    the kinematic env's (pioneer_knm_env.py:76-105, 151-211) evaluated on (q, qd, action)."""
    model = flatten_urdf()
    kin = OracleChain.from_model(model)
    n, limit, seed = 48, 7, 21
    cfg = DynEnvConfig(gravity=9.81, kp=300.0, kd=60.0, torque_scale=1e4, max_episode_steps=limit)
    term = CDynOracleBatch(model, n, cfg, env_id_base=100, seed=seed, obs_mode="terminal", threads=3)
    auto = CDynOracleBatch(model, n, cfg, env_id_base=100, seed=seed, obs_mode="autoreset", threads=1)
    s0 = term.state()
    lo, hi = np.array(model.lower, np.float32), np.array(model.upper, np.float32)
    for i in range(n):                                              # tick 0 draws of create
        u = reset_draws(seed, 100 + i, 0)
        q = np.array([np.float32(lo[j] + np.float32((hi[j] - lo[j]) * u[j])) for j in range(6)])
        assert np.array_equal(s0["q"][i], q.astype(np.float64))
    rng = np.random.default_rng(0)
    pot = np.zeros(n)
    episodes = 0
    for t in range(1, 20):
        act = rng.uniform(lo, hi, size=(n, 6)).astype(np.float32)
        a, b = term.step(act), auto.step(act)
        assert np.array_equal(a["flags"], b["flags"]) and np.array_equal(a["reward"], b["reward"])
        assert np.array_equal(a["own_q"], b["own_q"])               # the thread count does not change results
        obs = a["obs"]
        for i in range(0, n, 5):
            ptr = fk_pointer(kin, a["own_q"][i])
            assert np.abs(obs[i, 126:129] - ptr).max() < 1e-12
            d = np.linalg.norm(obs[i, 129:132] - ptr)
            assert abs(obs[i, 135] - d) < 1e-12 and abs(obs[i, 136] - 95.0 / (d / 10.0 + 1.0)) < 1e-12
            assert abs(a["reward"][i] - (obs[i, 136] - pot[i] - 0.01)) < 1e-12
        assert np.array_equal(obs[:, 0:6], a["own_q"]) and np.array_equal(obs[:, 90:96], a["own_qd"])
        assert np.array_equal(obs[:, 108:114], act.astype(np.float64))
        assert np.allclose(obs[:, 6:12], np.cos(obs[:, 0:6])) and np.allclose(obs[:, 102:108], np.sin(obs[:, 90:96]))
        done = (a["flags"] & 1).astype(bool)
        assert done.all() == (t % limit == 0) and ((a["flags"] & 2) != 0).all() == (t % limit == 0)
        pot = np.where(done, 0.0, obs[:, 136])
        if done.all():
            episodes += n
            # autoreset mode returns the first observation of the next episode: zero rates, zero action, potential 0
            assert not b["obs"][:, 90:96].any() and not b["obs"][:, 108:114].any() and not b["obs"][:, 136].any()
            assert np.array_equal(b["obs"][:, 0:6], auto.state()["q"])
        else:
            assert np.array_equal(a["obs"], b["obs"])
    st = term.stats
    assert st[0] == episodes and st[2] == episodes * limit and st[6] == 19 * n and st[7] == 0


def test_adoption_continues_from_the_given_float32_state():
    model = flatten_urdf()
    cfg = DynEnvConfig(gravity=9.81, kp=800.0, kd=200.0, torque_scale=1e4, max_episode_steps=0)
    n = 32
    a = CDynOracleBatch(model, n, cfg, seed=1)
    b = CDynOracleBatch(model, n, cfg, seed=1)
    rng = np.random.default_rng(2)
    act = rng.uniform(a.r_lo, a.r_hi, size=(n, 6)).astype(np.float32)
    ra = a.step(act)
    q32, qd32 = ra["own_q"].astype(np.float32), ra["own_qd"].astype(np.float32)
    mask = (np.arange(n) % 2).astype(np.uint8)
    rb = b.step(act, adopt=(q32 + np.float32(1e-3), qd32, mask))
    assert np.array_equal(ra["own_q"], rb["own_q"])                 # own = before adoption
    sb = b.state()
    assert np.array_equal(sb["q"][1::2], (q32 + np.float32(1e-3)).astype(np.float64)[1::2])
    assert np.array_equal(sb["q"][0::2], ra["own_q"][0::2])
    assert np.array_equal(rb["obs"][1::2, 0:6], sb["q"][1::2])      # the env layer sees the adopted state


def brute_force_segment_box(a, b, c, e, samples=50001):
    t = np.linspace(0.0, 1.0, samples)[:, None]
    x = a + t * (b - a) - c
    q = np.abs(x) - e
    sdf = np.linalg.norm(np.maximum(q, 0.0), axis=1) + np.minimum(q.max(axis=1), 0.0)
    return float(sdf.min())


def test_segment_box_distance_known_answers():
    c, e = np.array([10.0, 5.0, 0.0]), np.array([0.5, 0.5, 5.0])       # the reference demo's box (pioneer_knm_env.py:249-255)
    # a long link that straddles the box between two of the former 8 sample points (t = 3/7 and 4/7 of an 11-unit link)
    a, b = np.array([4.5, 5.0, 2.0]), np.array([15.5, 5.0, 2.0])
    assert abs(segment_box_distance(a, b, c, e) - (-0.5)) < 1e-12       # passes through the centre line: 0.5 deep
    # parallel to a face at distance 0.25
    assert abs(segment_box_distance([4.0, 5.75, 1.0], [16.0, 5.75, 1.0], c, e) - 0.25) < 1e-12
    # closest to an edge: the segment x + y = 17 runs diagonally past the vertical edge at (10.5, 5.5), nearest point (11, 6)
    assert abs(segment_box_distance([12.5, 4.5, 1.0], [9.5, 7.5, 1.0], c, e) - np.sqrt(0.5)) < 1e-12
    # end point nearest: segment pointing away from a corner
    assert abs(segment_box_distance([11.5, 6.5, 6.0], [14, 9, 9], c, e) - np.sqrt(3.0)) < 1e-12
    # fully inside: the deepest point is where the nearest face changes from x = 9.5 to y = 4.5 (t = 0.4): -0.46
    assert abs(segment_box_distance([9.8, 5.0, -1.0], [10.2, 5.1, 1.0], c, e) - (-0.46)) < 1e-12


def test_segment_box_distance_against_brute_force():
    rng = np.random.default_rng(7)
    worst = 0.0
    for k in range(300):
        c = rng.normal(size=3) * 3
        e = rng.uniform(0.1, 4.0, size=3)
        a = c + rng.normal(size=3) * 5
        b = c + rng.normal(size=3) * 5
        if k % 5 == 0:
            b[k % 3] = a[k % 3]                                          # axis-parallel segments (zero direction components)
        want = brute_force_segment_box(a, b, c, e)
        got = segment_box_distance(a, b, c, e)
        assert got <= want + 1e-12, (k, got, want)                       # the exact minimum is never above a sampled one
        # ... and a sampled one is at most one sample spacing above it (the signed distance is 1-Lipschitz)
        assert want - got <= np.linalg.norm(b - a) / 50000, (k, got, want)
        worst = max(worst, want - got)
    assert worst > 0.0
