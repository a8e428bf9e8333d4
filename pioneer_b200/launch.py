"""Env factories with the contract of the reference launcher (pioneer/launch/pioneer_knm_train.py:20-29):
``creator(env_config: dict) -> env``, registered under the same name when ray is importable.
The PPO trainer itself (Ray Tune / RLlib) is third-party and out of scope."""
from __future__ import annotations

from typing import Any, Dict

from .config import PioneerKinematicConfig
from .envs.pioneer import PioneerKinematicEnv
from .vector_env import PioneerVectorEnv
from .wrappers import TimeLimit

ENV_NAME = "Pioneer-v1"
MAX_EPISODE_STEPS = 500

DEFAULT_ENV_CONFIG = {"award_potential_slope": 10.0, "award_done": 5.0, "penalty_step": 1 / 100}


def _pioneer_config(env_config: Dict[str, Any]) -> PioneerKinematicConfig:
    return PioneerKinematicConfig(
        award_potential_slope=float(env_config["award_potential_slope"]),
        award_done=float(env_config["award_done"]),
        penalty_step=float(env_config["penalty_step"]),
    )


def prepare_env(env_config: Dict[str, Any]):
    """One env, exactly as the reference builds it: TimeLimit(PioneerKinematicEnv(cfg), 500)."""
    return TimeLimit(PioneerKinematicEnv(pioneer_config=_pioneer_config(env_config)), max_episode_steps=MAX_EPISODE_STEPS)


def prepare_vector_env(env_config: Dict[str, Any]):
    """``num_envs`` envs per rollout worker on the worker's GPU.  Extra keys: num_envs, seed,
    env_id_base, device; RLlib's worker_index / vector_index offset the global env ids."""
    num_envs = int(env_config.get("num_envs", 4096))
    worker = int(getattr(env_config, "worker_index", env_config.get("worker_index", 0)) or 0)
    base = int(env_config.get("env_id_base", worker * num_envs))
    return PioneerVectorEnv(num_envs, device=env_config.get("device"), pioneer_config=_pioneer_config(env_config),
                            max_episode_steps=MAX_EPISODE_STEPS, seed=int(env_config.get("seed", 0)), env_id_base=base)


def register(name: str = ENV_NAME, vectorized: bool = True) -> bool:
    """register_env(name, creator) when ray is importable; returns whether it was."""
    try:  # pragma: no cover - ray is not installable in the build container
        from ray.tune.registry import register_env
    except Exception:  # noqa: BLE001
        return False
    register_env(name, prepare_vector_env if vectorized else prepare_env)
    return True
