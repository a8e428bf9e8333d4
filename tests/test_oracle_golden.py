"""The CPU oracle replayed against golden vectors recorded from the unmodified reference source
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle.reach_oracle import OracleConfig, OracleEnv, fk_pointer, philox4x32_10
from tests._util import golden_case, load_golden, oracle_chain

G = load_golden()
CHAIN = oracle_chain()


def replay(case, config=None):
    env = OracleEnv(CHAIN, config, arith="np2")
    ep = 0
    env.reset_world(case["q0"][ep], case["target"][ep])
    first = env.observe()
    assert np.array_equal(first[:126], case["reset_obs"][ep][:126])
    np.testing.assert_allclose(first[126:], case["reset_obs"][ep][126:], rtol=1e-12, atol=1e-12)
    obs_at = {int(t): k for k, t in enumerate(case["obs_idx"])}
    for t, action in enumerate(case["actions"]):
        obs, reward, done, truncated = env.step(action)
        # joint state: bit-exact (pure IEEE arithmetic)
        assert np.array_equal(env.r, case["r"][t]), (t, env.r, case["r"][t])
        assert np.array_equal(env.v, case["v"][t]), (t, env.v, case["v"][t])
        assert np.array_equal(env.a, case["a"][t])
        assert done == bool(case["done"][t]), t
        assert truncated == bool(case["truncated"][t]), t
        # float64 FK by two different algorithms: agreement to rounding
        np.testing.assert_allclose(obs[126:137], case["tail"][t], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(reward, case["reward"][t], rtol=0, atol=1e-11)
        if t in obs_at:
            ref = case["obs"][obs_at[t]]
            assert np.array_equal(obs[:126], ref[:126]), t  # float32-valued entries incl. numpy's float32 sin/cos
        if done:
            ep += 1
            env.reset_world(case["q0"][ep], case["target"][ep])
            np.testing.assert_allclose(env.observe(), case["reset_obs"][ep], rtol=1e-12, atol=1e-12)
    return env


@pytest.mark.parametrize("name", ["cfg1", "gentle", "bangbang", "wild", "zero", "reach", "approach"])
def test_oracle_matches_reference_source(name):
    replay(golden_case(G, name))


def test_sampled_reset_keeps_float64_until_the_first_step_and_rounds_once():
    """The reference's own env.seed(123); env.reset() (pioneer_knm_env.py:80-94,107-109): np_random.uniform returns FLOAT64
    joint angles, `r` stays float64 until the first act(), and because a = v = 0 after a reset that act() stores
    float32(r0 + 0 + 0): the float64 start is rounded to float32 exactly once.  So rounding the sampled start to float32 at
    reset time (what the CUDA path and this oracle do) reproduces the reference trajectory BIT FOR BIT."""
    case = golden_case(G, "sampled")
    assert list(case["r_dtype_before_first_step"]) == ["float64", "float64"]
    assert not np.array_equal(case["q0_f64"][0].astype(np.float32).astype(np.float64), case["q0_f64"][0])   # really float64
    env = OracleEnv(CHAIN, arith="np2")
    rs = np.random.RandomState(int(case["seed"]))                       # gym.utils.seeding.np_random stand-in of the shim
    ep = 0
    for t, action in enumerate(case["actions"]):
        if t in case["reset_at"]:
            q0 = rs.uniform(env.r_lo, env.r_hi)                         # the reference's draw order and generator
            tgt = rs.uniform(np.array(env.config.target_lo), np.array(env.config.target_hi))
            assert np.array_equal(q0, case["q0_f64"][ep]) and np.array_equal(tgt.astype(np.float32), case["target_f64"][ep].astype(np.float32))
            env.reset_world(q0, tgt)                                    # rounds the float64 start to float32
            first = env.observe()
            # the reset observation of the reference still carries the float64 angles: float32 agreement only
            np.testing.assert_allclose(first[:126], case["reset_obs"][ep][:126], rtol=0, atol=6e-7)
            np.testing.assert_allclose(first[126:], case["reset_obs"][ep][126:], rtol=0, atol=2e-5)
            ep += 1
        obs, reward, done, truncated = env.step(action)
        assert np.array_equal(env.r, case["r"][t]) and np.array_equal(env.v, case["v"][t]), t
        # the reference keeps the sampled TARGET as float64 too (it never passes through a float32 array); the device state
        # is float32, so target / difference / distance / potential carry its rounding: |d target| <= ulp(25) / 2 = 9.5e-7
        np.testing.assert_allclose(obs[126:137], case["tail"][t], rtol=0, atol=2e-6)
        np.testing.assert_allclose(reward, case["reward"][t], rtol=0, atol=2e-5)
    assert np.array_equal(case["r"][0], case["q0_f64"][0].astype(np.float32))   # step 1 moved nothing: r1 = float32(r0)


def test_oracle_reward_knobs():
    replay(golden_case(G, "knobs"), OracleConfig(award_potential_slope=4.0, award_done=7.5, penalty_step=0.02))


def test_oracle_multi_env():
    for k in range(16):
        env = OracleEnv(CHAIN, arith="np2")
        env.reset_world(G["multi__q0"][k], G["multi__target"][k])
        for t in range(200):
            obs, reward, done, truncated = env.step(G["multi__actions"][k, t])
            assert np.array_equal(env.r, G["multi__r"][k, t]) and np.array_equal(env.v, G["multi__v"][k, t])
            np.testing.assert_allclose(reward, G["multi__reward"][k, t], atol=1e-11)
            assert done == bool(G["multi__done"][k, t])
    # env 0 of the multi-env fixture is the cfg1 trajectory
    assert np.array_equal(G["multi__r"][0], G["cfg1__r"][:200])


def test_golden_cases_exercise_the_edges():
    """The fixtures really contain the branches the parity claim is about."""
    c = golden_case(G, "cfg1")
    v_max, r_lo, r_hi = G["const_v_max"], G["const_r_lo"], G["const_r_hi"]
    assert (np.abs(c["v"]) == v_max).any(), "velocity clamp never hit"
    assert (c["r"] == r_hi).any() or (c["r"] == r_lo).any(), "position limit never hit"
    assert c["truncated"].sum() == 2 and c["done"].sum() == 2  # TimeLimit at steps 500 and 1000
    reach = golden_case(G, "reach")
    assert (reach["done"] & ~reach["truncated"]).sum() >= 5, "distance-done never exercised"
    assert (reach["tail"][:, 9] < 0.1).any() and (reach["tail"][:, 9] > 0.1).any()
    assert reach["truncated"].any()
    approach = golden_case(G, "approach")
    hits = np.flatnonzero(approach["done"] & ~approach["truncated"])
    assert len(hits) >= 3 and (approach["tail"][hits - 1, 9] > 0.1).all(), "done must fire mid-episode while moving"
    wild = golden_case(G, "wild")
    assert np.abs(wild["a"]).max() > 1e6


def test_reference_constants():
    env = OracleEnv(CHAIN)
    assert np.array_equal(env.r_lo, G["const_r_lo"]) and np.array_equal(env.r_hi, G["const_r_hi"])
    assert np.array_equal(env.v_max, G["const_v_max"]) and np.array_equal(env.a_max, G["const_a_max"])
    assert env.dt == G["const_dt_eps"][0] and env.eps == G["const_dt_eps"][1]
    assert tuple(G["const_obs_shape"]) == (137,) and str(G["const_obs_dtype"]) == "float64"
    assert int(G["const_fps"]) == 24 and int(G["const_dof"]) == 6
    np.testing.assert_allclose([env.compute_potential(d) for d in (0.0, 0.1, 5.0, 20.0)],
                               [95.0, 94.0594059406, 63.3333333333, 31.6666666667], rtol=1e-11)
    np.testing.assert_array_equal(G["const_potential_kat"], [env.compute_potential(d) for d in (0.0, 0.1, 5.0, 20.0)])


def test_fk_known_answers():
    """Hand-derived known answers, SURVEY.md section 8(c) C5."""
    kat = {
        (0, 0, 0, 0, 0, 0): (14.6, 1.0, 15.9),
        (0.5, 0, 0, 0, 0, 0): (12.333279864995, 7.877195425512, 15.9),
        (0, 0.5, 0, 0, 0, 0): (18.997294851594, 1.0, 7.321202184764),
        (0, 0, 0.5, 0, 0, 0): (13.723613926947, 1.0, 8.667794003970),
        (0, 0, 0, 0.5, 0, 0): (14.6, 0.089091476652, 15.667406867592),
        (0, 0, 0, 0, 0.5, 0): (15.070205746153, 1.0, 13.941474928617),
        (0, 0, 0, 0, 0, 0.5): (14.6, 0.089091476652, 15.667406867592),
        (0.3, -0.4, 0.9, 1.1, -0.7, 2.0): (7.904033967723, 1.072102752440, 5.624962379045),
    }
    for q, p in kat.items():
        np.testing.assert_allclose(fk_pointer(CHAIN, q), p, atol=2e-12)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_fk_matches_the_product_of_exponentials_formulation():
    """An independent route to the same forward kinematics (Lynch & Park, 'Modern Robotics', ch. 4): space-frame screw
    axes read off the ZERO configuration straight from the URDF text (the shipped asset, whose kinematic tree is the
    reference's pioneer_knm_6dof.urdf:204-275) -- joint origins summed along the chain, joint axes as written -- and T(q) = exp([S1] q1) ... exp([S6] q6) M with scipy's matrix exponential.  It
    shares no code and no recursion with oracle.fk_pointer or the flattener (pioneer_b200/urdf.py)."""
    import xml.etree.ElementTree as ET
    from scipy.linalg import expm
    from pioneer_b200.urdf import DEFAULT_URDF
    root = ET.parse(DEFAULT_URDF).getroot()
    joints = {j.find("child").get("link"): j for j in root.findall("joint")}
    chain = []                                           # walk from the pointer up to the world link
    link = "robot:pointer" if "robot:pointer" in joints else None
    if link is None:                                     # link names may carry another prefix: take the leaf
        parents = {j.find("parent").get("link") for j in root.findall("joint")}
        link = next(c for c in joints if c not in parents)
    while link in joints:
        chain.append(joints[link])
        link = joints[link].find("parent").get("link")
    chain.reverse()
    vec = lambda s: np.array([float(x) for x in s.split()])
    pos = np.zeros(3)
    screws = []
    for j in chain:                                      # the shipped URDF has no rpy on any joint origin
        origin = j.find("origin")
        assert origin is None or origin.get("rpy") in (None, "0 0 0")
        pos = pos + (vec(origin.get("xyz")) if origin is not None and origin.get("xyz") else np.zeros(3))
        if j.get("type") == "revolute":
            w = vec(j.find("axis").get("xyz"))
            screws.append(np.concatenate([w, -np.cross(w, pos)]))
        else:
            assert j.get("type") == "fixed"
    assert len(screws) == 6
    home = pos                                           # pointer position at q = 0

    def twist(S):
        w, v = S[:3], S[3:]
        m = np.zeros((4, 4))
        m[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        m[:3, 3] = v
        return m

    rng = np.random.default_rng(5)
    for _ in range(200):
        q = rng.uniform(-3.1416, 3.1416, size=6)
        T = np.eye(4)
        for S, qi in zip(screws, q):
            T = T @ expm(twist(S) * qi)
        want = (T @ np.append(home, 1.0))[:3]
        np.testing.assert_allclose(fk_pointer(CHAIN, q), want, atol=1e-10)
