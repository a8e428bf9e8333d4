"""gym.spaces.Box when gym is importable, otherwise a minimal stand-in with the same attributes
(the reference builds its spaces at pioneer/envs/pioneer/pioneer_knm_env.py:72-73, 238-242)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gym is not installable in the build container
    from gym import Env as GymEnv
    from gym.spaces import Box
except Exception:  # noqa: BLE001
    class GymEnv:  # type: ignore[no-redef]
        metadata: dict = {}
        reward_range = (-float("inf"), float("inf"))
        action_space = None
        observation_space = None

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
