"""RLlib VectorEnv adapter over BatchedPioneerEnv.

The reference registers a plain gym.Env and lets RLlib vectorise it one env per worker
(pioneer/launch/pioneer_knm_train.py:20-29, num_envs_per_worker = 1).  This adapter gives the rollout
worker the same observations / rewards / dones for ``num_envs`` envs per call.  It subclasses
ray.rllib.env.VectorEnv when ray is importable and is a duck-typed stand-in otherwise
(vector_reset / reset_at / vector_step / get_unwrapped).

RLlib's list-of-arrays API costs a device->host copy and O(N) Python objects per step; at large N use
the tensor fast path (reset_tensor / step_tensor), which stays on the device.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .batched_env import BatchedPioneerEnv
from .config import BatchConfig, PioneerKinematicConfig, SimulationConfig

try:  # pragma: no cover - ray is not installable in the build container
    from ray.rllib.env.vector_env import VectorEnv as _RllibVectorEnv
except Exception:  # noqa: BLE001
    _RllibVectorEnv = object


class _EnvView:
    """What get_unwrapped() hands out: a light per-env view (spaces + index)."""

    def __init__(self, parent: "PioneerVectorEnv", index: int):
        self.parent, self.index = parent, index
        self.observation_space, self.action_space = parent.observation_space, parent.action_space

    def observe(self):
        return self.parent.batch.observe(indices=[self.index])[0].cpu().numpy()


class PioneerVectorEnv(_RllibVectorEnv):
    def __init__(self, num_envs: int, device=None, pioneer_config: Optional[PioneerKinematicConfig] = None,
                 simulation_config: Optional[SimulationConfig] = None, max_episode_steps: int = 500,
                 seed: int = 0, env_id_base: int = 0, batch_config: Optional[BatchConfig] = None):
        bc = batch_config or BatchConfig()
        bc.max_episode_steps, bc.auto_reset, bc.obs_mode = max_episode_steps, True, "terminal"
        self.batch = BatchedPioneerEnv(num_envs, device=device, pioneer_config=pioneer_config,
                                       simulation_config=simulation_config, batch_config=bc, seed=seed,
                                       env_id_base=env_id_base)
        self.num_envs = num_envs
        self.observation_space = self.batch.observation_space
        self.action_space = self.batch.action_space
        if _RllibVectorEnv is not object:  # pragma: no cover
            try:
                super().__init__(self.observation_space, self.action_space, num_envs)
            except TypeError:
                pass
        self._fresh: Dict[int, np.ndarray] = {}     # first observation of envs the kernel already auto-reset

    # ---- tensor fast path ---------------------------------------------------------------------
    def reset_tensor(self) -> torch.Tensor:
        self._fresh.clear()
        return self.batch.reset()

    def step_tensor(self, actions: torch.Tensor):
        """(terminal-or-current obs, reward, flags) as device tensors; finished envs are already reset,
        their first observation is available through batch.observe(indices)."""
        return self.batch.step_tensor(actions)

    # ---- RLlib VectorEnv API ------------------------------------------------------------------
    def vector_reset(self) -> List[np.ndarray]:
        obs = self.reset_tensor().cpu().numpy()
        return list(obs)

    def reset_at(self, index: int) -> np.ndarray:
        if index in self._fresh:                    # the step kernel already started the next episode
            return self._fresh.pop(index)
        return self.batch.reset(indices=[index])[0].cpu().numpy()

    def vector_step(self, actions) -> Tuple[List[np.ndarray], List[float], List[bool], List[Dict[str, Any]]]:
        act = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, self.batch.dof)
        obs, reward, flags = self.batch.step_host(act)
        done = (flags & 1) != 0
        infos: List[Dict[str, Any]] = [{} for _ in range(self.num_envs)]
        finished = np.flatnonzero(done)
        self._fresh.clear()
        if finished.size:
            first = self.batch.observe(indices=finished).cpu().numpy()
            for k, i in enumerate(finished):
                i = int(i)
                self._fresh[i] = first[k]
                if flags[i] & 2:
                    infos[i]["TimeLimit.truncated"] = True
                else:
                    infos[i]["TimeLimit.truncated"] = False
        return list(obs.copy()), reward.tolist(), done.tolist(), infos

    def get_unwrapped(self) -> List[_EnvView]:
        return [_EnvView(self, i) for i in range(self.num_envs)]

    def close(self):
        self.batch.close()
