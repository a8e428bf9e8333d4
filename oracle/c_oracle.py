"""ctypes wrapper of the C oracle (oracle/reach_oracle.c) — test infrastructure, NOT product code.

``COracleBatch`` has the interface of ``oracle.reach_oracle.OracleBatch`` (step / reset / state / stats) and
the same results: joint state bit for bit, float64 kinematics to rounding, float32 cos/sin to ~1 ulp (glibc
cosf/sinf vs numpy's float32 cos/sin)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import build_c
from .c_dyn_oracle import fill_contact, orc_contact
from .reach_oracle import DOF, OBS_DIM, OracleChain, OracleConfig

_d3, _d9 = C.c_double * 3, C.c_double * 9


class orc_params(C.Structure):
    _fields_ = [("axis", _d3 * DOF), ("origin_xyz", _d3 * DOF), ("origin_rot", _d9 * DOF), ("tip_xyz", _d3),
                ("lower", C.c_double * DOF), ("upper", C.c_double * DOF),
                ("max_v_to_r", C.c_double), ("max_a_to_v", C.c_double), ("done_distance", C.c_double),
                ("award_max", C.c_double), ("award_done", C.c_double), ("award_potential_slope", C.c_double),
                ("penalty_step", C.c_double), ("target_lo", _d3), ("target_hi", _d3), ("timestep", C.c_double),
                ("frame_skip", C.c_int32), ("max_episode_steps", C.c_int32), ("legacy", C.c_int32),
                ("auto_reset", C.c_int32), ("obs_autoreset", C.c_int32), ("contact", orc_contact)]


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c.build())
        P = C.c_void_p
        lib.orc_create.restype = P
        lib.orc_create.argtypes = [C.POINTER(orc_params), C.c_int64, C.c_int64, C.c_uint64]
        lib.orc_destroy.argtypes = [P]
        lib.orc_bounds.argtypes = [P, P, P, P, P]
        lib.orc_reset.argtypes = [P, P, C.c_int64, P, P, P]
        lib.orc_observe.argtypes = [P, P]
        lib.orc_step.argtypes = [P, P, P, P, P]
        lib.orc_get_state.argtypes = [P, P, P, P, P, P, P, P]
        lib.orc_stats.argtypes = [P, P]
        lib.orc_get_boxes.argtypes = [P, P]
        lib.orc_sizeof_params.restype = C.c_int64
        assert lib.orc_sizeof_params() == C.sizeof(orc_params), "orc_params layout mismatch"
        _lib = lib
    return _lib


def _fill(dst, src):
    flat = np.ascontiguousarray(src, dtype=np.float64).ravel()
    C.memmove(dst, flat.ctypes.data, flat.nbytes)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class COracleBatch:
    def __init__(self, chain: OracleChain, n_envs: int, config: Optional[OracleConfig] = None, arith: str = "np2",
                 env_id_base: int = 0, seed: int = 0, auto_reset: bool = True, obs_mode: str = "terminal"):
        assert arith in ("np2", "legacy") and obs_mode in ("terminal", "autoreset")
        cfg = config or OracleConfig()
        self.lib = load()
        p = orc_params()
        _fill(p.axis, chain.axis); _fill(p.origin_xyz, chain.origin_xyz); _fill(p.origin_rot, chain.origin_rot)
        _fill(p.tip_xyz, chain.tip_xyz); _fill(p.lower, chain.lower); _fill(p.upper, chain.upper)
        p.max_v_to_r, p.max_a_to_v, p.done_distance = cfg.max_v_to_r, cfg.max_a_to_v, cfg.done_distance
        p.award_max, p.award_done = cfg.award_max, cfg.award_done
        p.award_potential_slope, p.penalty_step = cfg.award_potential_slope, cfg.penalty_step
        _fill(p.target_lo, cfg.target_lo); _fill(p.target_hi, cfg.target_hi)
        p.timestep, p.frame_skip, p.max_episode_steps = cfg.timestep, cfg.frame_skip, cfg.max_episode_steps or 0
        p.legacy, p.auto_reset, p.obs_autoreset = int(arith == "legacy"), int(auto_reset), int(obs_mode == "autoreset")
        fill_contact(p.contact, chain.capsules, cfg.obstacles, cfg.contact_penalty, getattr(cfg, "random_box", None))
        self.n = int(n_envs)
        self.h = self.lib.orc_create(C.byref(p), self.n, int(env_id_base), int(seed) & (2 ** 64 - 1))
        assert self.h, "orc_create failed"
        b = [np.zeros(DOF, np.float32) for _ in range(4)]
        self.lib.orc_bounds(self.h, *[x.ctypes.data for x in b])
        self.r_lo, self.r_hi, self.v_max, self.a_max = b

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def reset(self, idx=None, q0=None, target=None, observe: bool = True):
        idx_a = None if idx is None else np.ascontiguousarray(idx, dtype=np.int64)
        n = self.n if idx_a is None else len(idx_a)
        q0_a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.float32).reshape(n, DOF)
        tg_a = None if target is None else np.ascontiguousarray(target, dtype=np.float32).reshape(n, 3)
        obs = np.zeros((n, OBS_DIM), np.float64) if observe else None
        self.lib.orc_reset(self.h, _ptr(idx_a), n, _ptr(q0_a), _ptr(tg_a), _ptr(obs))
        return obs

    def observe(self):
        obs = np.zeros((self.n, OBS_DIM), np.float64)
        self.lib.orc_observe(self.h, obs.ctypes.data)
        return obs

    def step(self, actions, want_obs: bool = True):
        act = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, DOF)
        obs = np.zeros((self.n, OBS_DIM), np.float64) if want_obs else None
        reward = np.zeros(self.n, np.float64)
        flags = np.zeros(self.n, np.uint8)
        self.lib.orc_step(self.h, act.ctypes.data, _ptr(obs), reward.ctypes.data, flags.ctypes.data)
        return obs, reward, flags

    def state(self):
        n = self.n
        s = dict(r=np.zeros((n, DOF), np.float32), v=np.zeros((n, DOF), np.float32), a=np.zeros((n, DOF), np.float32),
                 potential=np.zeros(n, np.float64), target=np.zeros((n, 3), np.float64), t=np.zeros(n, np.int32),
                 ep_return=np.zeros(n, np.float32))
        self.lib.orc_get_state(self.h, *[s[k].ctypes.data for k in ("r", "v", "a", "potential", "target", "t", "ep_return")])
        return s

    def boxes(self):
        out = np.zeros((self.n, 6), np.float64)
        self.lib.orc_get_boxes(self.h, out.ctypes.data)
        return out

    @property
    def stats(self):
        out = np.zeros(8, np.float64)
        self.lib.orc_stats(self.h, out.ctypes.data)
        return out
