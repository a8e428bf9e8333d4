"""CPU ORACLE — test infrastructure, NOT product code.

Line-by-line restatement of the reference reach environment for ONE env, with every NumPy dtype
promotion written out explicitly so the result does not depend on the NumPy version it runs under.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under pioneer_b200/ does.

Parity status
  * Everything that lives in the reference's own source (integrator, clamps, reward, done,
    observation layout, step/reset ordering) is PINNED: tests/golden/make_golden.py imports the
    unmodified reference modules from /root/reference (with stand-ins for gym / pybullet) and
    records their outputs; tests/test_oracle_golden.py replays them through this file.
    Those fixtures are produced under NumPy 2.x, i.e. they pin ``arith='np2'``.
  * ``arith='legacy'`` (NumPy 1.x scalar promotion, what the reference computed in its own era:
    Python 3.7, bin/docker-cli:35) is restated from NumPy's documented promotion rules and is
    UNPINNED — no NumPy 1.x is installable here.
  * Forward kinematics is Bullet's in the reference (pybullet is not installable here, SURVEY.md
    section 0.4): PARITY UNPINNED vs PyBullet.  It is restated from the URDF conventions and checked
    against the hand-derived known answers of SURVEY.md section 8(c) C5 and against an independent
    homogeneous-matrix tree FK (tests/golden/_shim/pybullet_utils/bullet_client.py).

Reference files followed (relative to /root/reference):
  pioneer/envs/pioneer/pioneer_knm_env.py:56-61    bounds
  pioneer/envs/pioneer/pioneer_knm_env.py:76-105   reset_world
  pioneer/envs/pioneer/pioneer_knm_env.py:111-182  act
  pioneer/envs/pioneer/pioneer_knm_env.py:184-211  observe
  pioneer/envs/pioneer/pioneer_knm_env.py:232-236  compute_potential
  pioneer/envs/bullet/bullet_env.py:187-197        reset / step ordering
  pioneer/envs/bullet/bullet_scene.py:273-279      frame_skip substeps, step_time
  pioneer/launch/pioneer_knm_train.py:27           TimeLimit(max_episode_steps=500)
  gym.wrappers.TimeLimit (third party, unpinned in requirements.txt:14; restated from memory)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
f64 = np.float64

DOF = 6
OBS_DIM = 21 * DOF + 11

PNR_DONE = 1
PNR_TRUNCATED = 2


# --------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) — the reset
# generator shared bit-for-bit with the CUDA path (pioneer_b200/csrc/pnr_philox.cuh).
# --------------------------------------------------------------------------------------------
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> Tuple[int, int, int, int]:
    c0, c1, c2, c3 = (int(x) & _MASK for x in ctr)
    k0, k1 = (int(x) & _MASK for x in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def _u01(x: int) -> np.float32:
    """24 high bits -> float32 in [0, 1)."""
    return f32(x >> 8) * f32(2.0 ** -24)


def reset_draws(seed: int, global_env_id: int, tick: int, n: int = 9):
    """Uniforms for one reset: 6 joints then 3 target coordinates (reference draw order, pioneer_knm_env.py:80-90); with
    ``n=14`` also the per-env box of the obstacle variant (3 half extents, 2 centre coordinates: the draw order of
    pioneer/temp/pioneer_env.py:173-174).  Counter = (env id lo, env id hi, tick, block), key = seed."""
    key = (seed & _MASK, (seed >> 32) & _MASK)
    out: List[np.float32] = []
    for block in range(4):
        out.extend(_u01(x) for x in philox4x32_10((global_env_id & _MASK, (global_env_id >> 32) & _MASK,
                                                   tick & _MASK, block), key))
    return out[:n]


# --------------------------------------------------------------------------------------------
# configuration (reference defaults)
# --------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    # PioneerKinematicConfig, pioneer_knm_env.py:19-34
    max_v_to_r: float = 2
    max_a_to_v: float = 10
    done_distance: float = 0.1
    award_max: float = 100.0
    award_done: float = 5.0
    award_potential_slope: float = 10.0
    penalty_step: float = 1 / 100
    target_lo: Tuple[float, float, float] = (15, -10, 2)
    target_hi: Tuple[float, float, float] = (25, 10, 6)
    # SimulationConfig, bullet_env.py:36-44
    timestep: float = 1 / 240
    frame_skip: int = 10
    gravity: float = 0
    # gym TimeLimit, pioneer_knm_train.py:27
    max_episode_steps: int = 500
    # obstacle variant (this repo's extension; the reference only ever instantiates a box and a
    # plane in its GUI demo, pioneer_knm_env.py:249-261): list of (kind, position, extent), kind in plane/box/sphere
    obstacles: Tuple = ()
    contact_penalty: float = 0.0
    # (pos_lo[2], pos_hi[2], size_lo[3], size_hi[3]): the first box among ``obstacles`` is redrawn at every reset, half
    # extents ~ U(size), centre = (U(pos), half height) -- the legacy randomizer, pioneer/temp/pioneer_env.py:169-192
    random_box: Optional[Tuple] = None


@dataclass
class OracleChain:
    """float64 kinematic tables (same content as pioneer_b200.urdf.ChainModel, passed in by the
    test so that this module imports nothing from the product package)."""
    axis: np.ndarray
    origin_xyz: np.ndarray
    origin_rot: np.ndarray
    tip_xyz: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    capsules: Tuple = ()          # (body, radius, p0, p1) in the moving frame of `body`

    @staticmethod
    def from_model(m) -> "OracleChain":
        caps = tuple((int(b), float(r), np.array(p0, f64), np.array(p1, f64)) for b, r, p0, p1 in getattr(m, "capsules", ()))
        return OracleChain(np.array(m.axis, f64), np.array(m.origin_xyz, f64), np.array(m.origin_rot, f64),
                           np.array(m.tip_xyz, f64), np.array(m.lower, f64), np.array(m.upper, f64), caps)


def fk_pointer(chain: OracleChain, q) -> np.ndarray:
    """World position of 'robot:pointer' in float64 (what getLinkState(...).link_world_position
    returns, bullet_scene.py:58).  URDF convention: child = parent * T(origin) * Rot(axis, q).
    Evaluated tip-to-base with Rodrigues' rotation formula."""
    p = np.array(chain.tip_xyz, f64)
    for j in range(DOF - 1, -1, -1):
        k = chain.axis[j]
        qj = float(q[j])
        c, s = math.cos(qj), math.sin(qj)
        p = p * c + np.cross(k, p) * s + k * (float(k @ p) * (1.0 - c))
        p = chain.origin_xyz[j] + chain.origin_rot[j] @ p
    return p


def fk_point(chain: OracleChain, q, body: int, point) -> np.ndarray:
    """World position of a point fixed in the moving frame of ``body`` (same recursion as fk_pointer)."""
    p = np.array(point, f64)
    for j in range(body, -1, -1):
        k = chain.axis[j]
        c, s = math.cos(float(q[j])), math.sin(float(q[j]))
        p = p * c + np.cross(k, p) * s + k * (float(k @ p) * (1.0 - c))
        p = chain.origin_xyz[j] + chain.origin_rot[j] @ p
    return p


def box_sdf(x: np.ndarray, ext: np.ndarray) -> float:
    """Signed distance of ``x`` (relative to the box centre) to the axis-aligned box with half extents ``ext``."""
    qv = np.abs(x) - ext
    return float(np.linalg.norm(np.maximum(qv, 0.0))) + min(float(qv.max()), 0.0)


def segment_box_distance(a, b, centre, ext) -> float:
    """EXACT minimum of the box signed-distance function over the segment [a, b].  The function is convex along the
    segment and piecewise: the root of a quadratic between the parameters where a coordinate crosses a face plane, linear
    inside the box between the parameters where the nearest face changes.  Every breakpoint is enumerated, every piece
    minimised in closed form (same method as oracle/contact.h; the CUDA path enumerates the same breakpoints without sorting them, pnr_segment_box_exact)."""
    a0 = np.array(a, f64) - np.array(centre, f64)
    d = np.array(b, f64) - np.array(a, f64)
    e = np.array(ext, f64)
    ts = [0.0, 1.0]
    for i in range(3):
        if d[i] != 0.0:
            ts += [(e[i] - a0[i]) / d[i], (-e[i] - a0[i]) / d[i], -a0[i] / d[i]]
    for i in range(3):
        for j in range(i + 1, 3):
            for si in (-1.0, 1.0):
                for sj in (-1.0, 1.0):
                    den = si * d[i] - sj * d[j]
                    if den != 0.0:
                        ts.append((e[i] - e[j] - si * a0[i] + sj * a0[j]) / den)
    ts = sorted(t for t in ts if 0.0 <= t <= 1.0)
    best = min(box_sdf(a0 + t * d, e) for t in ts)
    for t0, t1 in zip(ts[:-1], ts[1:]):
        if t1 <= t0:
            continue
        x = a0 + 0.5 * (t0 + t1) * d
        active = np.abs(x) - e > 0.0
        if active.any():                                    # outside: f^2 = sum over the active axes of (u + t w)^2
            sgn = np.where(x > 0.0, 1.0, -1.0)
            u, w = (sgn * a0 - e)[active], (sgn * d)[active]
            if float(w @ w) > 0.0:
                tv = min(max(-float(u @ w) / float(w @ w), t0), t1)
                best = min(best, box_sdf(a0 + tv * d, e))
    return best


def contact_depth(chain: OracleChain, q, obstacles, box=None) -> float:
    """Sum over (link capsule, obstacle) pairs of the penetration depth max(0, radius - distance(segment, obstacle)),
    distance = the minimum of the obstacle's signed-distance function over the capsule's axis segment, exact for all kinds:
    plane : linear along the segment, the nearer end point decides
    sphere: closest point of the segment to the centre
    box   : segment_box_distance
    ``box`` = (centre, half extents) replaces the first box among ``obstacles`` (the per-env random box)."""
    total = 0.0
    for body, radius, p0, p1 in chain.capsules:
        a, b = fk_point(chain, q, body, p0), fk_point(chain, q, body, p1)
        first_box = True
        for kind, pos, ext in obstacles:
            pos, ext = np.array(pos, f64), np.array(ext, f64)
            if kind == "plane":
                d = min(float((a - pos) @ ext), float((b - pos) @ ext))
            elif kind == "sphere":
                ab = b - a
                t = min(max(float((pos - a) @ ab) / max(float(ab @ ab), 1e-30), 0.0), 1.0)
                d = float(np.linalg.norm(a + t * ab - pos)) - float(ext[0])
            elif kind == "box":
                if first_box and box is not None:
                    pos, ext = np.array(box[0], f64), np.array(box[1], f64)
                first_box = False
                d = segment_box_distance(a, b, pos, ext)
            else:
                raise ValueError(kind)
            total += max(0.0, radius - d)
    return total


class OracleEnv:
    """One reference env: PioneerKinematicEnv under TimeLimit, arithmetic written out.

    ``arith='np2'``    NumPy >= 2 (NEP 50): Python floats are weak, float32 stays float32.
    ``arith='legacy'`` NumPy 1.x: np.float32 scalar (op) Python float -> float64.
    """

    def __init__(self, chain: OracleChain, config: Optional[OracleConfig] = None, arith: str = "np2",
                 global_env_id: int = 0, seed: int = 0):
        assert arith in ("np2", "legacy")
        self.chain = chain
        self.config = config or OracleConfig()
        self.arith = arith
        self.global_env_id = global_env_id
        self.rng_seed = seed
        c = self.config
        # pioneer_knm_env.py:56-58, 217-220: limits are float32; python scalar * float32 array stays float32
        self.r_lo = np.array(chain.lower, dtype=f32)
        self.r_hi = np.array(chain.upper, dtype=f32)
        self.v_max = (f32(c.max_v_to_r) * (self.r_hi - self.r_lo)).astype(f32)
        self.a_max = (f32(c.max_a_to_v) * self.v_max).astype(f32)
        self.dt = c.timestep * c.frame_skip          # bullet_scene.py:277-279 (Python float)
        self.eps = 1e-5                              # pioneer_knm_env.py:61
        self.a = np.zeros(DOF, f32)
        self.v = np.zeros(DOF, f32)
        self.r = np.zeros(DOF, f32)
        self.target = np.zeros(3, f64)
        self.potential = 0.0
        self.elapsed = 0                             # TimeLimit._elapsed_steps
        self.ep_return = f32(0)                      # float32 accumulator, as the device keeps it
        self.box = None                              # (centre, half extents) of the per-env random box

    # -- pioneer_knm_env.py:76-105 -----------------------------------------------------------
    def reset_world(self, joint_positions=None, target_position=None, tick: int = 0):
        """Injected values are rounded to float32 (the device stores float32 state); sampled values
        come from Philox keyed on (seed, global env id, tick) instead of the reference's MT19937."""
        u = reset_draws(self.rng_seed, self.global_env_id, tick, 14)
        rb = self.config.random_box
        if rb is not None:
            pos_lo, pos_hi, size_lo, size_hi = (np.array(x, f32) for x in rb)
            ext = [f32(size_lo[i] + f32((size_hi[i] - size_lo[i]) * u[9 + i])) for i in range(3)]
            pos = [f32(pos_lo[i] + f32((pos_hi[i] - pos_lo[i]) * u[12 + i])) for i in range(2)] + [ext[2]]
            self.box = (np.array(pos, f64), np.array(ext, f64))
        if joint_positions is None:
            joint_positions = [f32(self.r_lo[i] + f32((self.r_hi[i] - self.r_lo[i]) * u[i])) for i in range(DOF)]
        if target_position is None:
            lo = np.array(self.config.target_lo, f32)
            hi = np.array(self.config.target_hi, f32)
            target_position = [f32(lo[i] + f32((hi[i] - lo[i]) * u[6 + i])) for i in range(3)]
        self.a = np.zeros(DOF, f32)
        self.v = np.zeros(DOF, f32)
        self.r = np.array(joint_positions, dtype=f32)
        self.target = np.array(np.array(target_position, dtype=f32), dtype=f64)
        self.potential = 0.0                          # pioneer_knm_env.py:105
        self.elapsed = 0
        self.ep_return = f32(0)

    # -- pioneer_knm_env.py:111-146 ----------------------------------------------------------
    def _integrate(self):
        a0, v0, r0 = self.a, self.v, self.r
        v1 = np.zeros(DOF, f32)
        r1 = np.zeros(DOF, f32)
        legacy = self.arith == "legacy"
        for i in range(DOF):
            if legacy:
                dt, eps = f64(self.dt), f64(self.eps)
                v1[i] = f32(f64(v0[i]) + f64(a0[i]) * dt)                          # :121
                dt_p1, dt_p2 = dt, f64(0)
                if v1[i] > self.v_max[i]:                                             # :125
                    q = f64(f32(self.v_max[i] - v0[i])) / (f64(a0[i]) + eps)          # :126
                    dt_p1 = _clip(q, f64(0), dt)
                    dt_p2 = dt - dt_p1
                    v1[i] = self.v_max[i]
                elif v1[i] < -self.v_max[i]:                                          # :129
                    q = f64(f32(-self.v_max[i] - v0[i])) / (f64(a0[i]) + eps)         # :130
                    dt_p1 = _clip(q, f64(0), dt)
                    dt_p2 = dt - dt_p1
                    v1[i] = -self.v_max[i]
                half = f64(0.5) * f64(f32(v0[i] + v1[i]))                             # :134
                r1[i] = f32((f64(r0[i]) + half * dt_p1) + f64(v1[i]) * dt_p2)
            else:
                dt, eps = f32(self.dt), f32(self.eps)
                v1[i] = f32(v0[i] + f32(a0[i] * dt))
                dt_p1, dt_p2 = dt, f32(0)
                if v1[i] > self.v_max[i]:
                    q = f32(f32(self.v_max[i] - v0[i]) / f32(a0[i] + eps))
                    dt_p1 = _clip(q, f32(0), dt)
                    dt_p2 = f32(dt - dt_p1)
                    v1[i] = self.v_max[i]
                elif v1[i] < -self.v_max[i]:
                    q = f32(f32(-self.v_max[i] - v0[i]) / f32(a0[i] + eps))
                    dt_p1 = _clip(q, f32(0), dt)
                    dt_p2 = f32(dt - dt_p1)
                    v1[i] = -self.v_max[i]
                half = f32(f32(0.5) * f32(v0[i] + v1[i]))
                r1[i] = f32(f32(r0[i] + f32(half * dt_p1)) + f32(v1[i] * dt_p2))
            if r1[i] >= self.r_hi[i]:                                                  # :135
                r1[i] = self.r_hi[i]
                v1[i] = 0
            if r1[i] <= self.r_lo[i]:                                                  # :139
                r1[i] = self.r_lo[i]
                v1[i] = 0
        return v1, r1

    def compute_potential(self, distance: float) -> float:                            # :232-236
        c = self.config
        return (c.award_max - c.award_done) / (distance / c.award_potential_slope + 1)

    def act(self, action) -> Tuple[float, bool]:
        with np.errstate(all="ignore"):
            v1, r1 = self._integrate()
        self.a = np.array(action, dtype=f32)                                           # :144 (unclipped)
        self.v = v1
        self.r = r1
        pointer = fk_pointer(self.chain, self.r)                                       # :148-152
        diff = self.target - pointer
        distance = float(np.linalg.norm(diff))
        old_potential = self.potential
        self.potential = self.compute_potential(distance)
        done = distance < self.config.done_distance                                    # :160
        reward = (self.potential - old_potential) + (-self.config.penalty_step) \
            + (self.config.award_done if done else 0)                                  # :162-165
        if self.config.obstacles and self.config.contact_penalty:
            self.last_contact_depth = contact_depth(self.chain, self.r, self.config.obstacles, self.box)
            reward -= self.config.contact_penalty * self.last_contact_depth
        # world.step(): frame_skip x stepSimulation with g = 0, qdot = 0, tau = 0 leaves q unchanged (:181)
        return float(reward), bool(done)

    # -- pioneer_knm_env.py:184-211 ----------------------------------------------------------
    def observe(self) -> np.ndarray:
        pointer = fk_pointer(self.chain, self.r)
        diff = self.target - pointer
        distance = np.linalg.norm(diff)
        r_lo_dist = self.r - self.r_lo
        r_hi_diff = self.r_hi - self.r
        with np.errstate(all="ignore"):
            return np.concatenate([
                self.r, np.cos(self.r), np.sin(self.r),
                self.r_lo, np.cos(self.r_lo), np.sin(self.r_lo),
                self.r_hi, np.cos(self.r_hi), np.sin(self.r_hi),
                r_lo_dist, np.cos(r_lo_dist), np.sin(r_lo_dist),
                r_hi_diff, np.cos(r_hi_diff), np.sin(r_hi_diff),
                self.v, np.cos(self.v), np.sin(self.v),
                self.a, np.cos(self.a), np.sin(self.a),
                pointer, self.target, diff,
                np.array([distance]), np.array([self.potential])]).astype(f64)

    # -- bullet_env.py:192-197 through gym TimeLimit ---------------------------------------------
    def step(self, action):
        reward, done = self.act(action)
        obs = self.observe()
        self.elapsed += 1
        truncated = False
        m = self.config.max_episode_steps
        if m and self.elapsed >= m:
            truncated = not done                       # info['TimeLimit.truncated']
            done = True
        self.ep_return = f32(self.ep_return + f32(reward))
        return obs, reward, done, truncated


def _clip(x, lo, hi):
    """np.clip for scalars: min(max(x, lo), hi) with NaN propagation."""
    if x < lo:
        return lo
    if x > hi:
        return hi
    return x


class OracleBatch:
    """N independent OracleEnv with the batched call semantics of the C-ABI (include/pioneer_b200.h):
    tick-keyed resets, in-step auto-reset, terminal-or-autoreset observations, episode statistics."""

    def __init__(self, chain: OracleChain, n_envs: int, config: Optional[OracleConfig] = None,
                 arith: str = "np2", env_id_base: int = 0, seed: int = 0, auto_reset: bool = True,
                 obs_mode: str = "terminal"):
        self.envs = [OracleEnv(chain, config, arith, env_id_base + i, seed) for i in range(n_envs)]
        self.n = n_envs
        self.auto_reset = auto_reset
        self.obs_mode = obs_mode
        self.tick = 0
        self.stats = np.array([0, 0, 0, 0, -np.inf, np.inf, 0, 0], dtype=f64)
        self.reset()

    def reset(self, idx=None, q0=None, target=None) -> np.ndarray:
        idx = range(self.n) if idx is None else idx
        obs = np.zeros((len(idx), OBS_DIM), f64)
        for k, i in enumerate(idx):
            self.envs[i].reset_world(None if q0 is None else q0[k], None if target is None else target[k], self.tick)
            obs[k] = self.envs[i].observe()
        self.tick += 1
        return obs

    def step(self, actions):
        obs = np.zeros((self.n, OBS_DIM), f64)
        reward = np.zeros(self.n, f64)
        flags = np.zeros(self.n, np.uint8)
        for i, e in enumerate(self.envs):
            o, r, d, tr = e.step(actions[i])
            obs[i], reward[i] = o, r
            flags[i] = (PNR_DONE if d else 0) | (PNR_TRUNCATED if tr else 0)
            self.stats[6] += 1
            if d:
                ret, length = float(e.ep_return), float(e.elapsed)
                self.stats[0] += 1; self.stats[1] += ret; self.stats[2] += length; self.stats[3] += ret * ret
                self.stats[4] = max(self.stats[4], ret); self.stats[5] = min(self.stats[5], ret)
                self.stats[7] += 0 if tr else 1
                if self.auto_reset:
                    e.reset_world(None, None, self.tick)
                    if self.obs_mode == "autoreset":
                        obs[i] = e.observe()
        self.tick += 1
        return obs, reward, flags

    def state(self):
        g = lambda name, dt: np.array([getattr(e, name) for e in self.envs], dtype=dt)
        return dict(r=g("r", f32), v=g("v", f32), a=g("a", f32), potential=g("potential", f64),
                    target=g("target", f64), t=g("elapsed", np.int32), ep_return=g("ep_return", f32))
