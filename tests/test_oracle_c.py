"""The C oracle (oracle/reach_oracle.c) against the golden vectors recorded from the unmodified reference source
and against the Python oracle.  CPU only."""
import numpy as np
import pytest

from oracle.c_oracle import COracleBatch
from oracle.reach_oracle import OracleBatch, OracleConfig
from tests._util import golden_case, load_golden, oracle_chain

G = load_golden()
CHAIN = oracle_chain()
VALUE_COLS = np.r_[0:6, 18:24, 36:42, 54:60, 72:78, 90:96, 108:114]
TRIG_COLS = np.r_[6:18, 24:36, 42:54, 60:72, 78:90, 96:108, 114:126]


def replay(name, config=None):
    case = golden_case(G, name)
    cfg = config or OracleConfig()
    env = COracleBatch(CHAIN, 1, cfg, arith="np2", auto_reset=False)
    ep = 0
    first = env.reset(q0=case["q0"][ep][None], target=case["target"][ep][None])[0]
    assert np.array_equal(first[VALUE_COLS], case["reset_obs"][ep][VALUE_COLS])
    np.testing.assert_allclose(first, case["reset_obs"][ep], rtol=0, atol=2e-7)
    obs_at = {int(t): k for k, t in enumerate(case["obs_idx"])}
    for t, action in enumerate(case["actions"]):
        obs, reward, flags = env.step(action[None])
        s = env.state()
        assert np.array_equal(s["r"][0], case["r"][t]), (name, t)          # bit-exact joint state
        assert np.array_equal(s["v"][0], case["v"][t]), (name, t)
        assert np.array_equal(s["a"][0], case["a"][t])
        assert bool(flags[0] & 1) == bool(case["done"][t]) and bool(flags[0] & 2) == bool(case["truncated"][t]), (name, t)
        np.testing.assert_allclose(obs[0, 126:137], case["tail"][t], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(reward[0], case["reward"][t], rtol=0, atol=1e-11)
        if t in obs_at:
            ref = case["obs"][obs_at[t]]
            assert np.array_equal(obs[0, VALUE_COLS], ref[VALUE_COLS]), t
            np.testing.assert_allclose(obs[0, TRIG_COLS], ref[TRIG_COLS], rtol=0, atol=2e-7)   # glibc vs numpy float32 trig
        if case["done"][t]:
            ep += 1
            env.reset(q0=case["q0"][ep][None], target=case["target"][ep][None])


@pytest.mark.parametrize("name", ["cfg1", "gentle", "bangbang", "wild", "zero", "reach", "approach"])
def test_c_oracle_matches_reference_source(name):
    replay(name)


def test_c_oracle_reward_knobs():
    replay("knobs", OracleConfig(award_potential_slope=4.0, award_done=7.5, penalty_step=0.02))


@pytest.mark.parametrize("arith", ["np2", "legacy"])
@pytest.mark.parametrize("obs_mode", ["terminal", "autoreset"])
def test_c_oracle_matches_python_oracle(arith, obs_mode):
    """Philox resets, auto-reset, statistics, both arithmetic modes: the two restatements agree."""
    n, steps = 23, 40
    cfg = OracleConfig(max_episode_steps=9)
    py = OracleBatch(CHAIN, n, cfg, arith=arith, env_id_base=77, seed=5, obs_mode=obs_mode)
    cc = COracleBatch(CHAIN, n, cfg, arith=arith, env_id_base=77, seed=5, obs_mode=obs_mode)
    rng = np.random.default_rng(3)
    for t in range(steps):
        act = (rng.uniform(-1, 1, size=(n, 6)) * cc.a_max * (30.0 if t % 7 == 0 else 1.0)).astype(np.float32)
        o1, r1, f1 = py.step(act)
        o2, r2, f2 = cc.step(act)
        assert np.array_equal(f1, f2), t
        s1, s2 = py.state(), cc.state()
        for k in ("r", "v", "a", "t", "ep_return"):
            assert np.array_equal(s1[k], s2[k]), (k, t)
        np.testing.assert_allclose(s1["target"], s2["target"], rtol=0, atol=0)
        np.testing.assert_allclose(r1, r2, rtol=0, atol=1e-11)
        np.testing.assert_allclose(o1, o2, rtol=0, atol=2e-7)
        assert np.array_equal(o1[:, VALUE_COLS], o2[:, VALUE_COLS])
    np.testing.assert_allclose(py.stats, cc.stats, rtol=1e-12)
    assert py.stats[0] == n * (steps // 9)


def test_c_oracle_matches_python_oracle_on_random_configs():
    """Property check (hypothesis): for random reward / limit / timing knobs and random (also out-of-range) actions the two
    independent restatements produce the same joint state bit for bit and the same done masks."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.floats(0.5, 4.0), st.floats(2.0, 20.0), st.floats(0.05, 8.0), st.floats(0.0, 0.1), st.integers(1, 20),
           st.integers(2, 9), st.integers(0, 2 ** 31), st.sampled_from(["np2", "legacy"]))
    def check(max_v_to_r, max_a_to_v, done_distance, penalty_step, frame_skip, limit, seed, arith):
        cfg = OracleConfig(max_v_to_r=max_v_to_r, max_a_to_v=max_a_to_v, done_distance=done_distance,
                           penalty_step=penalty_step, frame_skip=frame_skip, max_episode_steps=limit)
        n = 6
        py = OracleBatch(CHAIN, n, cfg, arith=arith, seed=seed)
        cc = COracleBatch(CHAIN, n, cfg, arith=arith, seed=seed)
        rng = np.random.default_rng(seed)
        for t in range(12):
            act = (rng.uniform(-1, 1, size=(n, 6)) * cc.a_max * rng.choice([0.01, 1.0, 25.0])).astype(np.float32)
            _, r1, f1 = py.step(act)
            _, r2, f2 = cc.step(act)
            assert np.array_equal(f1, f2)
            s1, s2 = py.state(), cc.state()
            assert np.array_equal(s1["r"], s2["r"]) and np.array_equal(s1["v"], s2["v"]) and np.array_equal(s1["t"], s2["t"])
            np.testing.assert_allclose(r1, r2, rtol=0, atol=1e-10)

    check()
