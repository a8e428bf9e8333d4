"""N independent Pioneer reach envs on one B200, stepped by one fused CUDA kernel per call.

Host-side mirror of the reference env (pioneer/envs/pioneer/pioneer_knm_env.py,
pioneer/envs/bullet/bullet_env.py) with a leading env dimension: same method names
(reset / step / seed / reset_world / observe / compute_potential / joint_limits), same attributes
(r_lo, r_hi, v_max, a_max, dt, eps, a, v, r, potential, action_space, observation_space), tensors
instead of arrays.  PyTorch is only the owner of device memory and streams here; all env arithmetic
happens in libpioneer_b200.so through the C-ABI (include/pioneer_b200.h).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _cabi
from .config import BatchConfig, PioneerKinematicConfig, SimulationConfig
from .spaces import Box
from .urdf import DEFAULT_URDF, ChainModel, flatten_urdf

DOF = _cabi.PNR_DOF
OBS_DIM = _cabi.PNR_OBS_DIM
STATS_FIELDS = ("episodes", "sum_return", "sum_length", "sum_return_sq", "max_return", "min_return",
                "env_steps", "reached_target")

_OBSTACLE_KIND = {"plane": _cabi.PNR_OBSTACLE_PLANE, "box": _cabi.PNR_OBSTACLE_BOX,
                  "sphere": _cabi.PNR_OBSTACLE_SPHERE}


def make_config(pioneer_config: PioneerKinematicConfig, simulation_config: SimulationConfig,
                batch_config: BatchConfig) -> "_cabi.pnr_config":
    lib = _cabi.load()
    c = _cabi.pnr_config()
    lib.pnr_default_config(C.byref(c))
    pc, sc, bc = pioneer_config, simulation_config, batch_config
    c.max_v_to_r, c.max_a_to_v = float(pc.max_v_to_r), float(pc.max_a_to_v)
    c.done_distance = float(pc.done_distance)
    c.award_max, c.award_done = float(pc.award_max), float(pc.award_done)
    c.award_potential_slope, c.penalty_step = float(pc.award_potential_slope), float(pc.penalty_step)
    assert len(pc.target_lo) == 3 and len(pc.target_hi) == 3          # pioneer_knm_env.py:84-85
    for k in range(3):
        c.target_lo[k], c.target_hi[k] = float(pc.target_lo[k]), float(pc.target_hi[k])
    c.timestep, c.frame_skip, c.gravity = float(sc.timestep), int(sc.frame_skip), float(sc.gravity)
    c.max_episode_steps = int(bc.max_episode_steps or 0)
    c.arith = {"f32": _cabi.PNR_ARITH_F32, "legacy64": _cabi.PNR_ARITH_LEGACY64}[bc.arith]
    c.obs_mode = {"terminal": _cabi.PNR_OBS_TERMINAL, "autoreset": _cabi.PNR_OBS_AUTORESET}[bc.obs_mode]
    c.auto_reset = 1 if bc.auto_reset else 0
    c.mode = {"kinematic": _cabi.PNR_MODE_KINEMATIC, "dynamic": _cabi.PNR_MODE_DYNAMIC}[bc.mode]
    c.kp, c.kd, c.torque_scale = float(bc.kp), float(bc.kd), float(bc.torque_scale)
    assert len(bc.obstacles) <= _cabi.PNR_MAX_OBSTACLES, "too many obstacles"
    c.n_obstacles = len(bc.obstacles)
    for i, ob in enumerate(bc.obstacles):
        c.obstacle_type[i] = _OBSTACLE_KIND[ob.kind]
        for k in range(3):
            c.obstacle_p[i][k] = float(ob.position[k])
            c.obstacle_e[i][k] = float(ob.extent[k])
    c.contact_penalty = float(bc.contact_penalty)
    c.random_box = 1 if bc.random_box else 0
    for k in range(2):
        c.box_pos_lo[k], c.box_pos_hi[k] = float(bc.box_pos_lo[k]), float(bc.box_pos_hi[k])
    for k in range(3):
        c.box_size_lo[k], c.box_size_hi[k] = float(bc.box_size_lo[k]), float(bc.box_size_hi[k])
    c.stepping = {"explicit": _cabi.PNR_STEPPING_EXPLICIT, "bullet": _cabi.PNR_STEPPING_BULLET}[bc.stepping]
    c.link_damping, c.max_velocity = float(bc.link_damping), float(bc.max_velocity)
    c.motor_kp, c.motor_kd, c.motor_max_force = float(bc.motor_kp), float(bc.motor_kd), float(bc.motor_max_force)
    return c


class BatchedPioneerEnv:
    """``n_envs`` reach envs resident on one GPU.

    ``env_id_base`` is the global id of local env 0: reset randomness is keyed on
    (seed, global env id, call counter), so a job sharded over several GPUs draws the same episodes
    as the same job on one GPU.
    """

    def __init__(self, n_envs: int, device: Optional[torch.device] = None,
                 pioneer_config: Optional[PioneerKinematicConfig] = None,
                 simulation_config: Optional[SimulationConfig] = None,
                 batch_config: Optional[BatchConfig] = None,
                 seed: int = 0, env_id_base: int = 0, urdf_path: str = DEFAULT_URDF,
                 chain: Optional[ChainModel] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("pioneer_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = _cabi.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError(f"pioneer_b200 runs on CUDA devices only, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_envs = int(n_envs)
        self.config = pioneer_config or PioneerKinematicConfig()
        self.simulation_config = simulation_config or SimulationConfig()
        self.batch_config = batch_config or BatchConfig()
        self.chain = chain or flatten_urdf(urdf_path)
        self.env_id_base = int(env_id_base)
        self._model = _cabi.model_from_chain(self.chain)
        self._cfg = make_config(self.config, self.simulation_config, self.batch_config)
        self._h = C.c_void_p()
        self._seed = int(seed) & (2 ** 64 - 1)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_create(C.byref(self._model), C.byref(self._cfg), self.n_envs, self.env_id_base,
                                             self.device.index, self._seed, C.byref(self._h)), "pnr_create")
        bounds = [(C.c_float * DOF)() for _ in range(4)]
        _cabi.check(self._lib.pnr_get_bounds(self._h, *bounds), "pnr_get_bounds")
        self.r_lo, self.r_hi, self.v_max, self.a_max = (np.array(b, dtype=np.float32) for b in bounds)
        self.dt = self.simulation_config.timestep * self.simulation_config.frame_skip     # bullet_scene.py:277-279
        self.eps = 1e-5                                                                   # pioneer_knm_env.py:61
        self.metadata = {"render.modes": [], "video.frames_per_second": self.simulation_config.frames_per_second}
        # spaces of ONE env (pioneer_knm_env.py:72-74); the GPU emits float32 observations (the reference
        # concatenates float32 and float64 pieces into float64)
        bc = self.batch_config
        if bc.mode == "dynamic":
            # the action is a motor set point, not the kinematic env's acceleration: desired joint positions under PD /
            # POSITION_CONTROL, joint torques (clamped to +-effort * torque_scale) otherwise
            if bc.kp or bc.kd or (bc.stepping == "bullet" and bc.motor_max_force > 0):
                self.action_space = Box(self.r_lo.copy(), self.r_hi.copy(), dtype=np.float32)
            else:
                tau = (np.asarray(self.chain.effort, np.float64) * bc.torque_scale).astype(np.float32)
                self.action_space = Box(-tau, tau, dtype=np.float32)
        else:
            self.action_space = Box(-self.a_max, self.a_max, dtype=np.float32)
        self.observation_space = Box(np.full(OBS_DIM, -np.inf, np.float32), np.full(OBS_DIM, np.inf, np.float32),
                                     dtype=np.float32)
        self.reward_range = (-float("inf"), float("inf"))
        kw = dict(device=self.device)
        self._obs = torch.empty((self.n_envs, OBS_DIM), dtype=torch.float32, **kw)
        self._reward = torch.empty(self.n_envs, dtype=torch.float32, **kw)
        self._flags = torch.empty(self.n_envs, dtype=torch.uint8, **kw)
        self._host: Optional[Dict[str, torch.Tensor]] = None
        self._host_ring = None
        self._filters = []                                   # MeanStdObsFilter objects bound to this handle
        self.step_index = 0

    # ---- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            for f in getattr(self, "_filters", ()):          # a filter must not outlive the handle it points into
                f._h = None
            self._lib.pnr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # ---- helpers ----------------------------------------------------------------------------
    @property
    def dof(self) -> int:
        return DOF

    @property
    def num_envs(self) -> int:
        return self.n_envs

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, x, dtype, shape) -> torch.Tensor:
        t = torch.as_tensor(x, dtype=dtype) if not isinstance(x, torch.Tensor) else x.to(dtype)
        t = t.to(self.device, non_blocking=True).reshape(shape).contiguous()
        return t

    def joint_limits(self) -> Tuple[np.ndarray, np.ndarray]:
        return self.r_lo, self.r_hi

    def compute_potential(self, distance):
        m = self.config.award_max - self.config.award_done
        return m / (distance / self.config.award_potential_slope + 1)

    def seed(self, seed=None):
        """Re-key the reset generator (PioneerKinematicEnv.seed, pioneer_knm_env.py:107-109)."""
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
        self._seed = int(seed) & (2 ** 64 - 1)
        _cabi.check(self._lib.pnr_seed(self._h, self._seed), "pnr_seed")
        return [seed]

    # ---- reset ------------------------------------------------------------------------------
    def reset_world(self, joint_positions=None, target_position=None, indices=None, observe: bool = False):
        """reset_world(joint_positions, target_position) of the reference (pioneer_knm_env.py:76-105) for
        all envs or for ``indices``.  None = sampled uniformly on the device."""
        with torch.cuda.device(self.device):
            idx = None if indices is None else self._dev(indices, torch.int64, (-1,))
            n = self.n_envs if idx is None else idx.numel()
            q0 = None if joint_positions is None else self._dev(joint_positions, torch.float32, (n, DOF))
            tg = None if target_position is None else self._dev(target_position, torch.float32, (n, 3))
            out = None
            if observe:
                out = self._obs if idx is None else torch.empty((n, OBS_DIM), dtype=torch.float32, device=self.device)
            _cabi.check(self._lib.pnr_reset(self._h, None if idx is None else idx.data_ptr(), n,
                                            None if q0 is None else q0.data_ptr(),
                                            None if tg is None else tg.data_ptr(),
                                            None if out is None else out.data_ptr(), self._stream()), "pnr_reset")
        return out

    def reset(self, indices=None) -> torch.Tensor:
        """BulletEnv.reset (bullet_env.py:187-190): new episode(s), returns the first observation(s)."""
        if indices is None:
            self.step_index = 0
        return self.reset_world(indices=indices, observe=True)

    # ---- step -------------------------------------------------------------------------------
    def step_tensor(self, actions: torch.Tensor, out: Optional[Tuple[torch.Tensor, ...]] = None):
        """The fast path: one kernel launch, no other device work.  ``actions`` float32 [N, 6] on this
        device.  Returns (obs [N,137] f32, reward [N] f32, flags [N] u8 with PNR_DONE | PNR_TRUNCATED bits);
        unless ``out`` is given these are the env's own buffers, overwritten by the next call."""
        if actions.dtype != torch.float32 or actions.device != self.device or not actions.is_contiguous() \
                or actions.shape != (self.n_envs, DOF):
            actions = self._dev(actions, torch.float32, (self.n_envs, DOF))
        obs, reward, flags = out if out is not None else (self._obs, self._reward, self._flags)
        _cabi.check(self._lib.pnr_step(self._h, actions.data_ptr(), obs.data_ptr(), reward.data_ptr(),
                                       flags.data_ptr(), self._stream()), "pnr_step")
        self.step_index += 1
        return obs, reward, flags

    def step_many(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, flags: torch.Tensor) -> None:
        """A rollout fragment with pre-computed actions (pnr_step_many): ``actions`` [T,N,6] -> ``obs`` [T,N,137], ``reward``
        [T,N], ``flags`` [T,N]: T steps in ONE kernel launch (every CTA / warp runs the T steps on its own tiles), bit for
        bit what T ``step_tensor`` calls return."""
        T = actions.shape[0]
        assert actions.shape == (T, self.n_envs, DOF) and obs.shape == (T, self.n_envs, OBS_DIM)
        assert reward.shape == (T, self.n_envs) and flags.shape == (T, self.n_envs)
        assert actions.dtype == torch.float32 and obs.dtype == torch.float32 and flags.dtype == torch.uint8
        assert all(x[0].is_contiguous() for x in (actions, obs)) and reward.is_contiguous() and flags.is_contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_step_many(self._h, T, actions.data_ptr(), actions.stride(0), obs.data_ptr(), obs.stride(0),
                                                reward.data_ptr(), flags.data_ptr(), self._stream()), "pnr_step_many")
        self.step_index += T

    def capture_rollout(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, flags: torch.Tensor):
        """Capture T consecutive steps into ONE CUDA graph: ``actions`` [T,N,6] -> ``obs`` [T,N,137], ``reward`` [T,N],
        ``flags`` [T,N] (caller-owned device tensors, re-read / re-written on every replay).  The graph holds two nodes:
        the counter advance and one pnr_step_many launch that runs the T steps.  Returns the torch.cuda.CUDAGraph; ``graph.replay()``
        advances every env by T steps.  The host-side call counter that keys the reset generator is frozen into the
        graph, so the graph's first node advances a device-side counter by T (pnr_tick_advance): every replay draws
        fresh reset states."""
        T = actions.shape[0]
        assert actions.shape == (T, self.n_envs, DOF) and obs.shape == (T, self.n_envs, OBS_DIM)
        assert reward.shape == (T, self.n_envs) and flags.shape == (T, self.n_envs)
        for t_ in (actions, obs, reward, flags):
            assert t_.device == self.device and all(t_[k].is_contiguous() for k in range(T))
        # pnr_step stores observation tiles with 16-byte bulk copies: every obs[t] must start on a 16-byte boundary, i.e.
        # the row count of the allocation must be a multiple of 4 (allocate [T, ceil4(N), 137] and pass obs[:, :N])
        assert all(obs[k].data_ptr() % 16 == 0 for k in range(T)), "obs[t] must be 16-byte aligned (pad N to a multiple of 4)"
        assert actions.dtype == torch.float32 and obs.dtype == torch.float32 and flags.dtype == torch.uint8
        with torch.cuda.device(self.device):
            self.step_tensor(actions[0], out=(obs[0], reward[0], flags[0]))     # one-time kernel attribute setup
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                _cabi.check(self._lib.pnr_tick_advance(self._h, T, self._stream()), "pnr_tick_advance")
                self.step_many(actions, obs, reward, flags)                     # the T steps are ONE kernel node
        return graph

    def advance_reset_counter(self, n: int) -> None:
        """Advance the reset generator's counter by ``n`` on the device (pnr_tick_advance); what a captured rollout
        graph does as its first node."""
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_tick_advance(self._h, int(n), self._stream()), "pnr_tick_advance")

    def step(self, actions):
        """BulletEnv.step through gym TimeLimit (bullet_env.py:192-197, pioneer_knm_train.py:27) for every env:
        (obs, reward, done, info) with tensors; info['TimeLimit.truncated'] is a bool tensor."""
        with torch.cuda.device(self.device):
            obs, reward, flags = self.step_tensor(actions)
            done = (flags & _cabi.PNR_DONE).bool()
            info = {"TimeLimit.truncated": (flags & _cabi.PNR_TRUNCATED).bool(), "flags": flags}
        return obs, reward, done, info

    def step_host(self, actions: np.ndarray):
        """Same step with HOST arrays through pnr_step_host: H2D of the actions, the kernel, D2H of
        obs / reward / flags, all inside the library.  Returns numpy views of pinned buffers."""
        h = self._ensure_host()
        if isinstance(actions, torch.Tensor) and actions.device.type == "cpu" and actions.dtype == torch.float32 \
                and actions.is_contiguous() and actions.numel() == self.n_envs * DOF:
            src = actions                   # a host tensor (ideally pinned) is handed to the library as it is
        else:
            h["actions"].numpy()[...] = np.asarray(actions, dtype=np.float32).reshape(self.n_envs, DOF)
            src = h["actions"]
        _cabi.check(self._lib.pnr_step_host(self._h, src.data_ptr(), h["obs"].data_ptr(),
                                            h["reward"].data_ptr(), h["flags"].data_ptr()), "pnr_step_host")
        self.step_index += 1
        return h["obs"].numpy(), h["reward"].numpy(), h["flags"].numpy()

    def obs_constants(self) -> np.ndarray:
        """obs[18:54] of every row (r_lo, r_hi and their cos / sin), exactly as the kernels write them."""
        out = (C.c_float * 36)()
        _cabi.check(self._lib.pnr_get_obs_constants(self._h, out), "pnr_get_obs_constants")
        return np.array(out, dtype=np.float32)

    def expand_compact(self, compact: np.ndarray) -> np.ndarray:
        """float32 [n, 101] (PNR_HOST_COMPACT) -> float32 [n, 137], bit for bit what the full layout delivers."""
        compact = np.ascontiguousarray(compact, dtype=np.float32)
        n = compact.shape[0]
        full = np.empty((n, OBS_DIM), np.float32)
        _cabi.check(self._lib.pnr_expand_obs_host(self._h, compact.ctypes.data, full.ctypes.data, n), "pnr_expand_obs_host")
        return full

    def step_host_begin(self, actions: torch.Tensor, compact: bool = False):
        """Asynchronous, double-buffered host step (pnr_step_host_begin): ``actions`` is a pinned float32 [N,6] host tensor
        that must stay untouched until the matching ``step_host_end``.  At most two steps may be in flight; the results
        land in a ring of two pinned buffer sets.  ``compact=True`` moves 101 instead of 137 columns per row over PCIe."""
        assert actions.device.type == "cpu" and actions.dtype == torch.float32 and actions.is_contiguous() \
            and actions.numel() == self.n_envs * DOF
        if self._host_ring is None:
            pin = dict(pin_memory=True)
            self._host_ring = dict(slot=0, pending=[], sets=[
                dict(obs=torch.empty((self.n_envs, OBS_DIM), dtype=torch.float32, **pin),
                     reward=torch.empty(self.n_envs, dtype=torch.float32, **pin),
                     flags=torch.empty(self.n_envs, dtype=torch.uint8, **pin)) for _ in range(2)])
        ring = self._host_ring
        buf = ring["sets"][ring["slot"]]
        width = _cabi.PNR_OBS_COMPACT_DIM if compact else OBS_DIM
        _cabi.check(self._lib.pnr_step_host_begin(self._h, actions.data_ptr(), buf["obs"].data_ptr(), buf["reward"].data_ptr(),
                                                  buf["flags"].data_ptr(),
                                                  _cabi.PNR_HOST_COMPACT if compact else _cabi.PNR_HOST_FULL),
                    "pnr_step_host_begin")
        ring["pending"].append((ring["slot"], width, actions))
        ring["slot"] ^= 1
        self.step_index += 1

    def step_host_end(self):
        """Wait for the oldest step begun; returns numpy views (obs [N,137] or [N,101], reward, flags) of its pinned
        buffers, valid until two more steps have been begun."""
        ring = self._host_ring
        assert ring is not None and ring["pending"], "no host step in flight"
        _cabi.check(self._lib.pnr_step_host_end(self._h), "pnr_step_host_end")
        slot, width, _ = ring["pending"].pop(0)
        buf = ring["sets"][slot]
        obs = buf["obs"].numpy().reshape(-1)[:self.n_envs * width].reshape(self.n_envs, width)
        return obs, buf["reward"].numpy(), buf["flags"].numpy()

    def _ensure_host(self) -> Dict[str, torch.Tensor]:
        if self._host is None:
            pin = dict(pin_memory=True)
            self._host = dict(actions=torch.empty((self.n_envs, DOF), dtype=torch.float32, **pin),
                              obs=torch.empty((self.n_envs, OBS_DIM), dtype=torch.float32, **pin),
                              reward=torch.empty(self.n_envs, dtype=torch.float32, **pin),
                              flags=torch.empty(self.n_envs, dtype=torch.uint8, **pin))
        return self._host

    def host_action_buffer(self) -> torch.Tensor:
        """Pinned float32 [N, 6] staging tensor; step_host() skips its own copy when handed this tensor."""
        return self._ensure_host()["actions"]

    # ---- observation / state ------------------------------------------------------------------
    def observe(self, indices=None) -> torch.Tensor:
        """PioneerKinematicEnv.observe (pioneer_knm_env.py:184-211) on the current state."""
        with torch.cuda.device(self.device):
            idx = None if indices is None else self._dev(indices, torch.int64, (-1,))
            n = self.n_envs if idx is None else idx.numel()
            out = torch.empty((n, OBS_DIM), dtype=torch.float32, device=self.device)
            _cabi.check(self._lib.pnr_observe(self._h, None if idx is None else idx.data_ptr(), n, out.data_ptr(),
                                              self._stream()), "pnr_observe")
        return out

    def observe_done(self, flags: torch.Tensor, obs: torch.Tensor, terminal_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """RLlib's ``reset_at`` after a done, for all envs at once (pnr_observe_done): rows of ``obs`` whose env finished in
        the step that produced ``flags`` are replaced by the first observation of the new episode (what the policy must
        see next); their terminal rows go to ``terminal_out`` when given.  In place, fixed launch shape."""
        assert flags.dtype == torch.uint8 and flags.numel() == self.n_envs and obs.shape == (self.n_envs, OBS_DIM)
        assert obs.is_contiguous() and obs.dtype == torch.float32 and obs.device == self.device
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_observe_done(self._h, flags.data_ptr(), obs.data_ptr(),
                                                   None if terminal_out is None else terminal_out.data_ptr(),
                                                   self._stream()), "pnr_observe_done")
        return obs

    def state(self) -> Dict[str, torch.Tensor]:
        """r, v, a [N,6]; potential [N]; target [N,3]; t [N] int32 (TimeLimit._elapsed_steps); ep_return [N]."""
        kw = dict(device=self.device, dtype=torch.float32)
        s = dict(r=torch.empty((self.n_envs, DOF), **kw), v=torch.empty((self.n_envs, DOF), **kw),
                 a=torch.empty((self.n_envs, DOF), **kw), potential=torch.empty(self.n_envs, **kw),
                 target=torch.empty((self.n_envs, 3), **kw),
                 t=torch.empty(self.n_envs, device=self.device, dtype=torch.int32),
                 ep_return=torch.empty(self.n_envs, **kw))
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_get_state(self._h, s["r"].data_ptr(), s["v"].data_ptr(), s["a"].data_ptr(),
                                                s["potential"].data_ptr(), s["target"].data_ptr(), s["t"].data_ptr(),
                                                s["ep_return"].data_ptr(), self._stream()), "pnr_get_state")
        return s

    def set_state(self, r=None, v=None, a=None, potential=None, target=None, t=None, ep_return=None):
        n = self.n_envs
        with torch.cuda.device(self.device):
            keep = [None if x is None else self._dev(x, dt, shape) for x, dt, shape in (
                (r, torch.float32, (n, DOF)), (v, torch.float32, (n, DOF)), (a, torch.float32, (n, DOF)),
                (potential, torch.float32, (n,)), (target, torch.float32, (n, 3)), (t, torch.int32, (n,)),
                (ep_return, torch.float32, (n,)))]
            _cabi.check(self._lib.pnr_set_state(self._h, *[None if x is None else x.data_ptr() for x in keep],
                                                self._stream()), "pnr_set_state")

    def boxes(self) -> torch.Tensor:
        """The per-env random box of the obstacle variant (BatchConfig.random_box): float32 [N, 6] = centre, half extents."""
        out = torch.empty((self.n_envs, 6), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_get_boxes(self._h, out.data_ptr(), self._stream()), "pnr_get_boxes")
        return out

    def set_boxes(self, boxes) -> None:
        with torch.cuda.device(self.device):
            b = self._dev(boxes, torch.float32, (self.n_envs, 6))
            _cabi.check(self._lib.pnr_set_boxes(self._h, b.data_ptr(), self._stream()), "pnr_set_boxes")
            self._keep_boxes = b

    # ---- checkpoint / resume -------------------------------------------------------------------
    def state_dict(self) -> Dict[str, object]:
        """Everything needed to continue bit-identically: per-env state (host copies, incl. the random boxes), the reset
        generator's seed and call counter, the episode statistics of the current window (all eight numbers) and the running
        statistics of every observation filter bound to this env.  (The reference env has no save / restore; it is
        re-created from constructor arguments, pioneer_knm_env.py:38,51 -- RLlib checkpoints carry the filter state.)"""
        tick, steps, seed = C.c_uint32(), C.c_double(), C.c_uint64()
        _cabi.check(self._lib.pnr_get_counters(self._h, C.byref(tick), C.byref(steps), C.byref(seed)), "pnr_get_counters")
        sd = {"n_envs": self.n_envs, "env_id_base": self.env_id_base, "seed": int(seed.value), "tick": int(tick.value),
              "env_steps": float(steps.value), "step_index": self.step_index,
              "state": {k: v.cpu() for k, v in self.state().items()},
              "episode_stats": self.episode_stats(clear=False)}
        if self.batch_config.random_box and self.batch_config.contact_penalty and self.batch_config.obstacles:
            sd["boxes"] = self.boxes().cpu()
        if self._filters:
            f = self._filters[-1]
            sd["obs_filter"] = {"count": f.n, "mean": f.mean, "var": f.var}
        return sd

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        assert sd["n_envs"] == self.n_envs and sd["env_id_base"] == self.env_id_base, "checkpoint is for another shard"
        st = sd["state"]
        self.set_state(r=st["r"], v=st["v"], a=st["a"], potential=st["potential"], target=st["target"], t=st["t"],
                       ep_return=st["ep_return"])
        if "boxes" in sd:
            self.set_boxes(sd["boxes"])
        self.seed(int(sd["seed"]))
        _cabi.check(self._lib.pnr_set_counters(self._h, int(sd["tick"]), float(sd["env_steps"])), "pnr_set_counters")
        if "episode_stats" in sd:
            vals = (C.c_double * _cabi.PNR_STATS_LEN)(*[float(sd["episode_stats"][k]) for k in STATS_FIELDS])
            _cabi.check(self._lib.pnr_set_stats(self._h, vals), "pnr_set_stats")
        if "obs_filter" in sd and self._filters:
            f = sd["obs_filter"]
            self._filters[-1].set_stats(f["count"], f["mean"], f["var"])
        self.step_index = int(sd["step_index"])

    # the reference's per-env attributes, batched
    @property
    def r(self): return self.state()["r"]

    @property
    def v(self): return self.state()["v"]

    @property
    def a(self): return self.state()["a"]

    @property
    def potential(self): return self.state()["potential"]

    # ---- statistics ---------------------------------------------------------------------------
    def episode_stats(self, clear: bool = False) -> Dict[str, float]:
        """Episode statistics since the last clear (the columns the reference CLI prints, cli.py:32-38)."""
        out = (C.c_double * _cabi.PNR_STATS_LEN)()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_stats(self._h, out, 1 if clear else 0, self._stream()), "pnr_stats")
        return dict(zip(STATS_FIELDS, [float(x) for x in out]))

    def episode_stats_tensor(self, clear: bool = False) -> torch.Tensor:
        """The same 8 numbers as a float64 device tensor, not synchronised (for the NCCL all-reduce)."""
        out = torch.empty(_cabi.PNR_STATS_LEN, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.pnr_stats_device(self._h, out.data_ptr(), 1 if clear else 0, self._stream()),
                        "pnr_stats_device")
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.pnr_launch_count(self._h))
