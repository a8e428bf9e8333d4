"""Single-env facade with the reference's exact constructor and gym.Env contract
(pioneer/envs/pioneer/pioneer_knm_env.py:38-242), backed by a one-env BatchedPioneerEnv on the GPU.

Useful as a drop-in and for parity tests; throughput comes from BatchedPioneerEnv / PioneerVectorEnv.
Differences from the reference, all deliberate: observations carry float32 precision (returned as
float64 arrays of shape (137,) like the reference's), render() is not provided.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from ...batched_env import BatchedPioneerEnv
from ...config import BatchConfig, PioneerKinematicConfig, RenderConfig, SimulationConfig
from ...spaces import Box
from ...urdf import DEFAULT_URDF
from ..bullet.bullet_env import BulletEnv
from ..bullet.bullet_scene import Scene, World


def arr2str(a) -> str:
    """reference: pioneer/collections_util.py:13-14"""
    return "[" + ", ".join(f"{x:.3f}" for x in a) + "]"


class PioneerKinematicEnv(BulletEnv):
    def __init__(self, headless: bool = True,
                 pioneer_config: Optional[PioneerKinematicConfig] = None,
                 simulation_config: Optional[SimulationConfig] = None,
                 render_config: Optional[RenderConfig] = None,
                 device=None, arith: str = "f32"):
        # gym.utils.EzPickle (pioneer_knm_env.py:38,51): the env pickles as its constructor arguments
        self._ctor_args = (headless, pioneer_config, simulation_config, render_config, device, arith)
        self.np_random: Optional[np.random.RandomState] = None
        self.seed()
        self.config = pioneer_config or PioneerKinematicConfig()
        sim = simulation_config or SimulationConfig()
        # no TimeLimit and no auto-reset inside: the reference wraps the env in gym's TimeLimit itself
        self._batch = BatchedPioneerEnv(1, device=device, pioneer_config=self.config, simulation_config=sim,
                                        batch_config=BatchConfig(max_episode_steps=0, auto_reset=False, arith=arith))
        self._index = 0
        self._last_obs: Optional[np.ndarray] = None
        BulletEnv.__init__(self, DEFAULT_URDF, headless, sim, render_config)
        self.r_lo, self.r_hi = self.joint_limits()
        self.v_max = self._batch.v_max
        self.a_max = self._batch.a_max
        self.dt = self.world.step_time
        self.eps = 1e-5
        self.reset_world()
        self.action_space = Box(-self.a_max, self.a_max, dtype=np.float32)
        self.observation_space = self.observation_to_space(self.observe())
        self.reward_range = (-float("inf"), float("inf"))

    # BulletEnv.reset_simulator (bullet_env.py:90-99)
    def reset_simulator(self):
        sim = self.simulation_config
        self.world = World(sim.timestep, sim.frame_skip, sim.gravity)
        self.scene = Scene(self, self._batch.chain)
        super().reset_simulator()

    def reset_world(self, joint_positions=None, target_position: Optional[Tuple[float, float, float]] = None):
        # same draw order and host generator as the reference (pioneer_knm_env.py:80-90)
        r_lo, r_hi = self._batch.r_lo, self._batch.r_hi
        if joint_positions is None:
            joint_positions = self.np_random.uniform(r_lo, r_hi)
        if target_position is None:
            assert len(self.config.target_lo) == 3
            assert len(self.config.target_hi) == 3
            target_position = tuple(self.np_random.uniform(np.array(self.config.target_lo), np.array(self.config.target_hi)))
        assert len(list(joint_positions)) == self.dof                     # pioneer_knm_env.py:227
        self._batch.reset_world(joint_positions=np.asarray(joint_positions, np.float32)[None],
                                target_position=np.asarray(target_position, np.float32)[None])
        self._last_obs = None

    def seed(self, seed=None) -> List[int]:
        try:  # pragma: no cover - gym is not installable in the build container
            from gym.utils import seeding
            self.np_random, seed = seeding.np_random(seed)
        except Exception:  # noqa: BLE001
            if seed is None:
                seed = int(np.random.SeedSequence().generate_state(1)[0])
            self.np_random = np.random.RandomState(seed % (2 ** 32))
        return [seed]

    def act(self, action, world_index: int, step_index: int) -> Tuple[float, bool, Dict]:
        obs, reward, flags = self._batch.step_tensor(np.asarray(action, np.float32)[None])
        obs = obs[0].double().cpu().numpy()
        reward, done = float(reward[0]), bool(int(flags[0]) & 1)
        self._last_obs = obs
        pot_old_free = reward + self.config.penalty_step - (self.config.award_done if done else 0)
        info = {                                                           # pioneer_knm_env.py:167-179
            "r_pot": f"{pot_old_free:.3f}", "r_step": f"{-self.config.penalty_step:.3f}",
            "r_done": f"{(self.config.award_done if done else 0):.3f}", "rw": f"{reward:.3f}",
            "dist": f"{obs[135]:.3f}", "pot": f"{obs[136]:.3f}",
            "a": arr2str(obs[108:114]), "v": arr2str(obs[90:96]), "r": arr2str(obs[0:6]),
        }
        return reward, done, info

    def observe(self) -> np.ndarray:
        if self._last_obs is None:
            self._last_obs = self._batch.observe()[0].double().cpu().numpy()
        return self._last_obs

    @property
    def dof(self) -> int:
        return len(self.scene.joints)

    @property
    def a(self): return self._batch.state()["a"][0].cpu().numpy()

    @property
    def v(self): return self._batch.state()["v"][0].cpu().numpy()

    @property
    def r(self): return self._batch.state()["r"][0].cpu().numpy()

    @property
    def potential(self): return float(self._batch.state()["potential"][0])

    def joint_limits(self) -> Tuple[np.ndarray, np.ndarray]:
        lower = np.array([x.lower_limit for x in self.scene.joints], dtype=np.float32)
        upper = np.array([x.upper_limit for x in self.scene.joints], dtype=np.float32)
        return lower, upper

    def joint_positions(self) -> np.ndarray:
        return self.r.astype(np.float64)

    def reset_joint_positions(self, positions):
        positions_list = list(positions)
        assert len(positions_list) == self.dof
        self._batch.set_state(r=np.asarray(positions_list, np.float32)[None])
        self._last_obs = None

    def compute_potential(self, distance: float) -> float:
        return self._batch.compute_potential(distance)

    @staticmethod
    def observation_to_space(observation: np.ndarray) -> Box:
        low = np.full(observation.shape, -float("inf"), dtype=np.float32)
        high = np.full(observation.shape, float("inf"), dtype=np.float32)
        return Box(low, high, dtype=observation.dtype)

    def __reduce__(self):
        return (type(self), self._ctor_args)

    def close(self):
        self._batch.close()
