"""Generate tests/golden/reach_golden.npz by running the UNMODIFIED reference env.

Runs only in the build container: needs /root/reference.  The reference modules
(pioneer.envs.pioneer.PioneerKinematicEnv, pioneer.envs.bullet.*) are imported as they are; gym and
pybullet, which cannot be installed here, are replaced by the stand-ins in tests/golden/_shim
(see its README).  So every number recorded below comes out of the reference's own act(),
observe(), reset_world(), compute_potential(), BulletEnv.step() source, executed under the NumPy
of this container (2.x => the 'np2' arithmetic mode); forward kinematics comes from the stand-in's
float64 tree FK, not from Bullet.

    python tests/golden/make_golden.py            # rewrites reach_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("PIONEER_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REFERENCE)

from gym.wrappers import TimeLimit  # noqa: E402  (stand-in)
from pioneer.envs.pioneer import PioneerKinematicEnv, PioneerKinematicConfig  # noqa: E402  (reference)

f32 = np.float32


def make_env(**cfg):
    # same construction as the reference launcher (pioneer/launch/pioneer_knm_train.py:20-27)
    return TimeLimit(PioneerKinematicEnv(pioneer_config=PioneerKinematicConfig(**cfg)), max_episode_steps=500)


def inject_reset(tl, q0, target):
    """BulletEnv.reset() (bullet_env.py:187-190) with reset_world's optional arguments filled in."""
    e = tl.env
    tl._elapsed_steps = 0
    e.reset_simulator()
    e.reset_world(joint_positions=np.asarray(q0, dtype=f32), target_position=tuple(float(x) for x in target))
    return e.observe()


def rollout(tl, starts, actions, obs_every=1):
    """Step through `actions`; at every done inject the next (q0, target) of `starts`."""
    e = tl.env
    n = len(actions)
    rec = dict(r=np.zeros((n, 6), f32), v=np.zeros((n, 6), f32), a=np.zeros((n, 6), f32),
               reward=np.zeros(n), done=np.zeros(n, bool), truncated=np.zeros(n, bool),
               tail=np.zeros((n, 11)), episode=np.zeros(n, np.int32))
    obs_idx, obs_rows = [], []
    ep = 0
    reset_obs = [inject_reset(tl, *starts[ep])]
    for t in range(n):
        obs, reward, done, info = tl.step(actions[t])
        assert obs.dtype == np.float64 and obs.shape == (137,)
        rec["r"][t], rec["v"][t], rec["a"][t] = e.r, e.v, e.a
        rec["reward"][t], rec["done"][t] = reward, done
        rec["truncated"][t] = bool(info.get("TimeLimit.truncated", False))
        rec["tail"][t] = obs[126:137]
        rec["episode"][t] = ep
        if t % obs_every == 0 or done or t < 8:
            obs_idx.append(t)
            obs_rows.append(obs)
        if done:
            ep += 1
            reset_obs.append(inject_reset(tl, *starts[ep]))
    rec["obs_idx"] = np.array(obs_idx, np.int32)
    rec["obs"] = np.array(obs_rows)
    rec["reset_obs"] = np.array(reset_obs)
    rec["q0"] = np.array([s[0] for s in starts[:ep + 1]], f32)
    rec["target"] = np.array([s[1] for s in starts[:ep + 1]], f32)
    rec["actions"] = np.asarray(actions, f32)
    return rec


def starts_from(rng, env, n):
    lo, hi = np.array(env.config.target_lo, float), np.array(env.config.target_hi, float)
    return [(rng.uniform(env.r_lo, env.r_hi).astype(f32), rng.uniform(lo, hi).astype(f32)) for _ in range(n)]


def main():
    out = {}
    tl = make_env()
    e = tl.env
    out["const_r_lo"], out["const_r_hi"] = e.r_lo, e.r_hi
    out["const_v_max"], out["const_a_max"] = e.v_max, e.a_max
    out["const_dt_eps"] = np.array([e.dt, e.eps])
    out["const_action_low"], out["const_action_high"] = e.action_space.low, e.action_space.high
    out["const_obs_shape"] = np.array(e.observation_space.shape)
    out["const_obs_dtype"] = np.array(str(e.observation_space.dtype))
    out["const_fps"] = np.array(e.metadata["video.frames_per_second"])
    out["const_joint_names"] = np.array([j.name for j in e.scene.joints])
    out["const_item_names"] = np.array([i.name for i in e.scene.items])
    out["const_dof"] = np.array(e.dof)
    out["const_potential_kat"] = np.array([e.compute_potential(d) for d in (0.0, 0.1, 5.0, 20.0)])

    def put(name, rec):
        for k, v in rec.items():
            out[f"{name}__{k}"] = v

    # cfg1: single env, 1000 fixed-seed random-action steps (BASELINE.json configs[0])
    a_max = e.a_max
    rng_s, rng_a = np.random.default_rng(0), np.random.default_rng(1)
    acts = rng_a.uniform(-a_max, a_max, size=(1000, 6)).astype(f32)
    put("cfg1", rollout(tl, starts_from(rng_s, e, 4), acts, obs_every=10))

    # gentle actions: stays mostly inside the velocity limits, different clamp mix
    rng_s, rng_a = np.random.default_rng(10), np.random.default_rng(11)
    acts = (0.05 * rng_a.uniform(-a_max, a_max, size=(300, 6))).astype(f32)
    put("gentle", rollout(tl, starts_from(rng_s, e, 2), acts, obs_every=5))

    # bang-bang at +-a_max held for long stretches: velocity saturation and both position limits
    acts = np.zeros((400, 6), f32)
    for t in range(400):
        sign = 1.0 if (t // 40) % 2 == 0 else -1.0
        acts[t] = sign * a_max * (1.0 if t % 7 else 0.3)
    put("bangbang", rollout(tl, starts_from(np.random.default_rng(20), e, 2), acts, obs_every=5))

    # actions outside the action space (stored unclipped, pioneer_knm_env.py:144) incl. huge arguments to sin/cos
    rng_a = np.random.default_rng(31)
    acts = (rng_a.uniform(-1, 1, size=(120, 6)) * np.array([3.0, 30.0, 300.0, 3e3, 3e4, 3e6]) * a_max).astype(f32)
    put("wild", rollout(tl, starts_from(np.random.default_rng(30), e, 2), acts, obs_every=1))

    # zero actions: arm never moves, reward = -penalty_step from step 2 on
    put("zero", rollout(tl, starts_from(np.random.default_rng(40), e, 2), np.zeros((60, 6), f32), obs_every=1))

    # done by distance: targets placed a hair inside/outside done_distance of the pointer; zero actions.
    # The last start lies outside and runs into the time limit.
    rng = np.random.default_rng(50)
    starts = []
    for off in [0.05, 0.0999, 0.0, 0.099999, 0.02, 0.09, 0.0995, 0.100001, 0.1001]:
        q0 = rng.uniform(e.r_lo, e.r_hi).astype(f32)
        inject_reset(tl, q0, (20, 0, 4))
        pointer = e.observe()[126:129]
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        starts.append((q0, (pointer + off * d).astype(f32)))
    starts += starts_from(rng, e, 2)
    put("reach", rollout(tl, starts, np.zeros((530, 6), f32), obs_every=1))

    # approach: the arm sweeps towards a target placed on its future path, done fires mid-episode
    rng = np.random.default_rng(55)
    starts, acts = [], []
    for k in range(4):
        q0 = (0.5 * rng.uniform(e.r_lo, e.r_hi)).astype(f32)
        seq = np.tile((rng.uniform(-0.02, 0.02, size=6) * a_max).astype(f32), (40, 1))
        inject_reset(tl, q0, (20, 0, 4))
        for t in range(15 + 5 * k):
            obs, _, _, _ = tl.step(seq[t])
        target = obs[126:129].astype(f32)
        inject_reset(tl, q0, target)          # probe: how many steps until done fires
        n_k = next(t + 1 for t in range(40) if tl.step(seq[t])[2])
        starts.append((q0, target))
        acts.append(seq[:n_k])
    starts += starts_from(rng, e, 1)
    put("approach", rollout(tl, starts, np.concatenate(acts + [np.zeros((5, 6), f32)]), obs_every=1))

    # reward-shaping knobs the launcher sets through env_config (pioneer_knm_train.py:53-57)
    tl2 = make_env(award_potential_slope=4.0, award_done=7.5, penalty_step=0.02)
    rng_s, rng_a = np.random.default_rng(60), np.random.default_rng(61)
    acts = rng_a.uniform(-a_max, a_max, size=(80, 6)).astype(f32)
    put("knobs", rollout(tl2, starts_from(rng_s, tl2.env, 2), acts, obs_every=4))

    # 16 envs x 200 steps, env k seeded [0,k]/[1,k] (SURVEY.md 8(d) D2 cfg2 seeding); env 0 == cfg1 prefix
    multi = []
    for k in range(16):
        rng_s, rng_a = (np.random.default_rng(0), np.random.default_rng(1)) if k == 0 else \
            (np.random.default_rng([0, k]), np.random.default_rng([1, k]))
        acts = rng_a.uniform(-a_max, a_max, size=(1000 if k == 0 else 200, 6)).astype(f32)[:200]
        multi.append(rollout(tl, starts_from(rng_s, e, 4), acts, obs_every=50))
    for key in ("r", "v", "a", "reward", "done", "truncated", "tail", "actions"):
        out[f"multi__{key}"] = np.stack([m[key] for m in multi])
    out["multi__q0"] = np.stack([m["q0"][0] for m in multi])
    out["multi__target"] = np.stack([m["target"][0] for m in multi])

    # sampled resets: the reference's own env.seed(s); env.reset() (pioneer_knm_env.py:80-94,107-109).  np_random.uniform
    # returns FLOAT64 joint angles and `r` stays float64 until the first act(); a = v = 0 after a reset, so that act()
    # computes r1 = float32(r0 + 0 + 0): the float64 start is rounded to float32 exactly once, with nothing added.
    tl3 = make_env()
    e3 = tl3.env
    e3.seed(123)
    rng_a = np.random.default_rng(71)
    acts = rng_a.uniform(-a_max, a_max, size=(90, 6)).astype(f32)
    rec = dict(r=np.zeros((90, 6), f32), v=np.zeros((90, 6), f32), reward=np.zeros(90), tail=np.zeros((90, 11)),
               r_dtype_before_first_step=[], q0_f64=[], target_f64=[], reset_obs=[], reset_at=[])
    for t in range(90):
        if t in (0, 40):                                       # two seeded episodes from one RandomState stream
            obs0 = tl3.reset()
            rec["reset_at"].append(t)
            rec["r_dtype_before_first_step"].append(str(np.asarray(e3.r).dtype))
            rec["q0_f64"].append(np.array(e3.r, np.float64))
            rec["target_f64"].append(np.array(obs0[129:132], np.float64))
            rec["reset_obs"].append(obs0)
        obs, reward, done, info = tl3.step(acts[t])
        rec["r"][t], rec["v"][t], rec["reward"][t], rec["tail"][t] = e3.r, e3.v, reward, obs[126:137]
        assert np.asarray(e3.r).dtype == np.float32
    rec["actions"] = acts
    rec["seed"] = np.array(123)
    for k in ("q0_f64", "target_f64", "reset_obs", "reset_at", "r_dtype_before_first_step"):
        rec[k] = np.array(rec[k])
    put("sampled", rec)

    path = os.path.join(HERE, "reach_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB, numpy {np.__version__}")


if __name__ == "__main__":
    main()
