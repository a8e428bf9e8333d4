"""gym.wrappers.TimeLimit when gym is importable, otherwise an equivalent restated from its documented
behaviour (the reference launcher wraps the env with max_episode_steps=500,
pioneer/launch/pioneer_knm_train.py:27)."""
from __future__ import annotations

try:  # pragma: no cover - gym is not installable in the build container
    from gym.wrappers import TimeLimit
except Exception:  # noqa: BLE001
    class TimeLimit:  # type: ignore[no-redef]
        def __init__(self, env, max_episode_steps=None):
            self.env = env
            self._max_episode_steps = max_episode_steps
            self._elapsed_steps = None
            self.action_space = env.action_space
            self.observation_space = env.observation_space
            self.reward_range = env.reward_range
            self.metadata = env.metadata

        def step(self, action):
            assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
            observation, reward, done, info = self.env.step(action)
            self._elapsed_steps += 1
            if self._elapsed_steps >= self._max_episode_steps:
                info["TimeLimit.truncated"] = not done
                done = True
            return observation, reward, done, info

        def reset(self, **kwargs):
            self._elapsed_steps = 0
            return self.env.reset(**kwargs)

        def seed(self, seed=None):
            return self.env.seed(seed)

        def close(self):
            return self.env.close()

        @property
        def unwrapped(self):
            return getattr(self.env, "unwrapped", self.env)

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)
