"""ctypes wrapper of the C dynamic-mode oracle (oracle/dynamics_oracle.c) — test infrastructure, NOT product code.

``CDynOracleBatch`` is the dynamic-mode counterpart of ``oracle.c_oracle.COracleBatch``: a float64 env batch with the
call semantics of the C-ABI (tick-keyed Philox resets, in-step auto-reset, terminal-or-autoreset observations, episode
statistics), plus the lock-step interface described in the C file's header: ``step(actions, adopt=(q32, qd32, mask))``
returns what the oracle's own float64 substeps reached AND continues from the float32 state the device reached.
PARITY UNPINNED vs PyBullet (see oracle/dynamics_oracle.py)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import build_c

DOF, OBS_DIM = 6, 137
MAX_CAPSULES, MAX_OBSTACLES = 8, 4
_d3, _d9 = C.c_double * 3, C.c_double * 9
_KIND = {"plane": 1, "box": 2, "sphere": 3}


class orc_contact(C.Structure):
    _fields_ = [("n_capsules", C.c_int32), ("capsule_body", C.c_int32 * MAX_CAPSULES),
                ("capsule_radius", C.c_double * MAX_CAPSULES), ("capsule_p0", _d3 * MAX_CAPSULES),
                ("capsule_p1", _d3 * MAX_CAPSULES), ("n_obstacles", C.c_int32),
                ("obstacle_type", C.c_int32 * MAX_OBSTACLES), ("obstacle_p", _d3 * MAX_OBSTACLES),
                ("obstacle_e", _d3 * MAX_OBSTACLES), ("contact_penalty", C.c_double), ("random_box", C.c_int32),
                ("box_pos_lo", C.c_double * 2), ("box_pos_hi", C.c_double * 2), ("box_size_lo", _d3), ("box_size_hi", _d3)]


class dyn_params(C.Structure):
    _fields_ = [("axis", _d3 * DOF), ("origin_xyz", _d3 * DOF), ("origin_rot", _d9 * DOF), ("tip_xyz", _d3),
                ("lower", C.c_double * DOF), ("upper", C.c_double * DOF), ("effort", C.c_double * DOF),
                ("damping", C.c_double * DOF), ("body_mass", C.c_double * DOF), ("body_com", _d3 * DOF),
                ("body_inertia", _d9 * DOF),
                ("done_distance", C.c_double), ("award_max", C.c_double), ("award_done", C.c_double),
                ("award_potential_slope", C.c_double), ("penalty_step", C.c_double), ("target_lo", _d3), ("target_hi", _d3),
                ("timestep", C.c_double), ("gravity", C.c_double), ("kp", C.c_double), ("kd", C.c_double),
                ("torque_scale", C.c_double),
                ("frame_skip", C.c_int32), ("max_episode_steps", C.c_int32), ("auto_reset", C.c_int32),
                ("obs_autoreset", C.c_int32), ("stepping", C.c_int32),
                ("link_damping", C.c_double), ("max_velocity", C.c_double), ("motor_kp", C.c_double),
                ("motor_kd", C.c_double), ("motor_max_force", C.c_double), ("contact", orc_contact)]


@dataclass
class DynEnvConfig:
    """Everything pnr_config carries for the dynamic mode (defaults = the reference's, plus this repo's knobs)."""
    done_distance: float = 0.1
    award_max: float = 100.0
    award_done: float = 5.0
    award_potential_slope: float = 10.0
    penalty_step: float = 1 / 100
    target_lo: Tuple[float, float, float] = (15, -10, 2)
    target_hi: Tuple[float, float, float] = (25, 10, 6)
    timestep: float = 1 / 240
    frame_skip: int = 10
    gravity: float = 0.0
    kp: float = 0.0
    kd: float = 0.0
    torque_scale: float = 1.0
    max_episode_steps: int = 500
    obstacles: Tuple = ()                 # (kind, position, extent)
    contact_penalty: float = 0.0
    random_box: Optional[Tuple] = None    # (pos_lo[2], pos_hi[2], size_lo[3], size_hi[3]): the first box is redrawn per episode
    stepping: str = "explicit"            # 'explicit' | 'bullet'
    link_damping: float = 0.04
    max_velocity: float = 100.0
    motor_kp: float = 0.1
    motor_kd: float = 1.0
    motor_max_force: float = 0.0


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c.build())
        P = C.c_void_p
        lib.orc_dyn_create.restype = P
        lib.orc_dyn_create.argtypes = [C.POINTER(dyn_params), C.c_int64, C.c_int64, C.c_uint64]
        lib.orc_dyn_destroy.argtypes = [P]
        lib.orc_dyn_set_threads.argtypes = [P, C.c_int32]
        lib.orc_dyn_reset.argtypes = [P, P, C.c_int64, P, P, P]
        lib.orc_dyn_set_state.argtypes = [P, P, P]
        lib.orc_dyn_substeps.argtypes = [P, P, P, P, C.c_int32]
        lib.orc_dyn_substeps.restype = C.c_uint32
        lib.orc_dyn_aba.argtypes = [P, P, P, P, P]
        lib.orc_dyn_segment_box.restype = C.c_double
        lib.orc_dyn_segment_box.argtypes = [P, P, P, P]
        lib.orc_dyn_contact_depth.restype = C.c_double
        lib.orc_dyn_contact_depth.argtypes = [P, P, P, P]
        lib.orc_dyn_step.argtypes = [P] * 12
        lib.orc_dyn_get_state.argtypes = [P] * 9
        lib.orc_dyn_stats.argtypes = [P, P]
        lib.orc_dyn_sizeof_params.restype = C.c_int64
        assert lib.orc_dyn_sizeof_params() == C.sizeof(dyn_params), "dyn_params layout mismatch"
        _lib = lib
    return _lib


def _fill(dst, src):
    flat = np.ascontiguousarray(src, dtype=np.float64).ravel()
    C.memmove(dst, flat.ctypes.data, flat.nbytes)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def fill_contact(c: orc_contact, capsules: Sequence, obstacles: Sequence, contact_penalty: float, random_box) -> None:
    c.n_capsules = len(capsules)
    for i, (body, radius, p0, p1) in enumerate(capsules):
        c.capsule_body[i], c.capsule_radius[i] = int(body), float(radius)
        for k in range(3):
            c.capsule_p0[i][k], c.capsule_p1[i][k] = float(p0[k]), float(p1[k])
    c.n_obstacles = len(obstacles)
    c.random_box = -1
    for i, (kind, pos, ext) in enumerate(obstacles):
        c.obstacle_type[i] = _KIND[kind]
        for k in range(3):
            c.obstacle_p[i][k], c.obstacle_e[i][k] = float(pos[k]), float(ext[k])
        if random_box is not None and kind == "box" and c.random_box < 0:
            c.random_box = i
    c.contact_penalty = float(contact_penalty)
    if random_box is not None:
        assert c.random_box >= 0, "random_box needs a box among the obstacles"
        pos_lo, pos_hi, size_lo, size_hi = random_box
        for k in range(2):
            c.box_pos_lo[k], c.box_pos_hi[k] = float(pos_lo[k]), float(pos_hi[k])
        for k in range(3):
            c.box_size_lo[k], c.box_size_hi[k] = float(size_lo[k]), float(size_hi[k])


def segment_box_distance(a, b, centre, half) -> float:
    """min over the segment [a, b] of the signed distance to the axis-aligned box (exact, float64)."""
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (a, b, centre, half)]
    return float(load().orc_dyn_segment_box(*[x.ctypes.data for x in arrs]))


class CDynOracleBatch:
    def __init__(self, model, n_envs: int, config: Optional[DynEnvConfig] = None, env_id_base: int = 0, seed: int = 0,
                 auto_reset: bool = True, obs_mode: str = "terminal", threads: Optional[int] = None):
        """``model``: pioneer_b200.urdf.ChainModel (or anything with the same float64 tables)."""
        assert obs_mode in ("terminal", "autoreset")
        cfg = config or DynEnvConfig()
        self.lib, self.cfg = load(), cfg
        p = dyn_params()
        for name in ("axis", "origin_xyz", "origin_rot", "tip_xyz", "lower", "upper", "effort", "damping", "body_mass",
                     "body_com", "body_inertia"):
            _fill(getattr(p, name), getattr(model, name))
        p.done_distance, p.award_max, p.award_done = cfg.done_distance, cfg.award_max, cfg.award_done
        p.award_potential_slope, p.penalty_step = cfg.award_potential_slope, cfg.penalty_step
        _fill(p.target_lo, cfg.target_lo); _fill(p.target_hi, cfg.target_hi)
        p.timestep, p.gravity, p.kp, p.kd, p.torque_scale = cfg.timestep, cfg.gravity, cfg.kp, cfg.kd, cfg.torque_scale
        p.frame_skip, p.max_episode_steps = cfg.frame_skip, cfg.max_episode_steps or 0
        p.auto_reset, p.obs_autoreset = int(auto_reset), int(obs_mode == "autoreset")
        p.stepping = {"explicit": 0, "bullet": 1}[cfg.stepping]
        p.link_damping, p.max_velocity = cfg.link_damping, cfg.max_velocity
        p.motor_kp, p.motor_kd, p.motor_max_force = cfg.motor_kp, cfg.motor_kd, cfg.motor_max_force
        fill_contact(p.contact, getattr(model, "capsules", ()), cfg.obstacles, cfg.contact_penalty, cfg.random_box)
        self.n = int(n_envs)
        self.h = self.lib.orc_dyn_create(C.byref(p), self.n, int(env_id_base), int(seed) & (2 ** 64 - 1))
        assert self.h, "orc_dyn_create failed"
        if threads is None:
            try:
                threads = len(os.sched_getaffinity(0))
            except AttributeError:
                threads = os.cpu_count() or 1
        self.lib.orc_dyn_set_threads(self.h, int(threads))
        self.r_lo = np.array(model.lower, dtype=np.float32)
        self.r_hi = np.array(model.upper, dtype=np.float32)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_dyn_destroy(self.h)
            self.h = None

    def reset(self, idx=None, q0=None, target=None, observe: bool = True):
        idx_a = None if idx is None else np.ascontiguousarray(idx, dtype=np.int64)
        n = self.n if idx_a is None else len(idx_a)
        q0_a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.float32).reshape(n, DOF)
        tg_a = None if target is None else np.ascontiguousarray(target, dtype=np.float32).reshape(n, 3)
        obs = np.zeros((n, OBS_DIM), np.float64) if observe else None
        self.lib.orc_dyn_reset(self.h, _ptr(idx_a), n, _ptr(q0_a), _ptr(tg_a), _ptr(obs))
        return obs

    def set_state(self, q=None, qd=None):
        q_a = None if q is None else np.ascontiguousarray(q, dtype=np.float64).reshape(self.n, DOF)
        qd_a = None if qd is None else np.ascontiguousarray(qd, dtype=np.float64).reshape(self.n, DOF)
        self.lib.orc_dyn_set_state(self.h, _ptr(q_a), _ptr(qd_a))

    def substeps(self, q, qd, action, n_sub: int):
        """n_sub substeps of ONE env from (q, qd): the dynamics alone."""
        q_a, qd_a = np.array(q, np.float64), np.array(qd, np.float64)
        act = np.ascontiguousarray(action, dtype=np.float32)
        self.lib.orc_dyn_substeps(self.h, q_a.ctypes.data, qd_a.ctypes.data, act.ctypes.data, int(n_sub))
        return q_a, qd_a

    def aba(self, q, qd, tau):
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (q, qd, tau)]
        out = np.zeros(DOF, np.float64)
        self.lib.orc_dyn_aba(self.h, *[x.ctypes.data for x in arrs], out.ctypes.data)
        return out

    def contact_depth(self, q, box=None) -> float:
        q_a = np.ascontiguousarray(q, dtype=np.float64)
        bp = None if box is None else np.ascontiguousarray(box[:3], dtype=np.float64)
        be = None if box is None else np.ascontiguousarray(box[3:], dtype=np.float64)
        return float(self.lib.orc_dyn_contact_depth(self.h, q_a.ctypes.data, _ptr(bp), _ptr(be)))

    def step(self, actions, adopt=None, want_obs: bool = True, want_own: bool = True, want_depth: bool = False):
        """Returns dict(obs, reward, flags, own_q, own_qd, depth, touched); touched[i] bit j = joint j of env i ran into a
        stop during this step.  ``adopt`` = (q float32 [n,6], qd float32 [n,6], mask
        uint8 [n] or None): see the module docstring."""
        n = self.n
        act = np.ascontiguousarray(actions, dtype=np.float32).reshape(n, DOF)
        obs = np.zeros((n, OBS_DIM), np.float64) if want_obs else None
        reward, flags = np.zeros(n, np.float64), np.zeros(n, np.uint8)
        own_q = np.zeros((n, DOF), np.float64) if want_own else None
        own_qd = np.zeros((n, DOF), np.float64) if want_own else None
        depth = np.zeros(n, np.float64) if want_depth else None
        touched = np.zeros(n, np.uint8)
        aq = aqd = am = None
        if adopt is not None:
            aq = np.ascontiguousarray(adopt[0], dtype=np.float32).reshape(n, DOF)
            aqd = np.ascontiguousarray(adopt[1], dtype=np.float32).reshape(n, DOF)
            am = None if len(adopt) < 3 or adopt[2] is None else np.ascontiguousarray(adopt[2], dtype=np.uint8).reshape(n)
        self.lib.orc_dyn_step(self.h, act.ctypes.data, _ptr(aq), _ptr(aqd), _ptr(am), _ptr(own_q), _ptr(own_qd), _ptr(obs),
                              reward.ctypes.data, flags.ctypes.data, _ptr(depth), touched.ctypes.data)
        return dict(obs=obs, reward=reward, flags=flags, own_q=own_q, own_qd=own_qd, depth=depth, touched=touched)

    def state(self):
        n = self.n
        s = dict(q=np.zeros((n, DOF)), qd=np.zeros((n, DOF)), a=np.zeros((n, DOF), np.float32), potential=np.zeros(n),
                 target=np.zeros((n, 3)), t=np.zeros(n, np.int32), ep_return=np.zeros(n, np.float32), box=np.zeros((n, 6)))
        self.lib.orc_dyn_get_state(self.h, *[s[k].ctypes.data for k in ("q", "qd", "a", "potential", "target", "t",
                                                                        "ep_return", "box")])
        return s

    @property
    def stats(self):
        out = np.zeros(8, np.float64)
        self.lib.orc_dyn_stats(self.h, out.ctypes.data)
        return out
