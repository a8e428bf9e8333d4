"""BulletEnv: the reset()/step() template of the reference (pioneer/envs/bullet/bullet_env.py:65-209)
without a Bullet client behind it."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, Optional, Tuple

from ...config import RenderConfig, SimulationConfig
from ...spaces import GymEnv


class BulletEnv(GymEnv, ABC):
    def __init__(self, model_path: str, headless: bool = True,
                 simulation_config: Optional[SimulationConfig] = None,
                 render_config: Optional[RenderConfig] = None):
        if not headless:
            raise NotImplementedError("GUI mode is out of scope of pioneer_b200 (stepping path only)")
        self.model_path = model_path
        self.headless = headless
        self.simulation_config = simulation_config or SimulationConfig()
        self.render_config = render_config or RenderConfig()
        self.metadata = {"render.modes": ["human", "rgb_array"],
                         "video.frames_per_second": self.simulation_config.frames_per_second}
        self.world = None
        self.scene = None
        self.world_index = -1
        self.step_index = 0
        self.reset_simulator()

    def reset_simulator(self):
        """The reference rebuilds the physics client and reloads the URDF here on every episode
        (bullet_env.py:90-99); the device tables are immutable, so only the counters move."""
        self.world_index += 1
        self.step_index = 0

    def render(self, mode="human"):
        raise NotImplementedError("rendering is out of scope of pioneer_b200 (stepping path only)")

    def reset(self):
        self.reset_simulator()
        self.reset_world()
        return self.observe()

    def step(self, action) -> Tuple[object, float, bool, Dict]:
        self.step_index += 1
        reward, done, info = self.act(action, self.world_index, self.step_index)
        observation = self.observe()
        return observation, reward, done, info

    @abstractmethod
    def reset_world(self):
        pass

    @abstractmethod
    def act(self, action, world_index: int, step_index: int) -> Tuple[float, bool, Dict]:
        pass

    @abstractmethod
    def observe(self):
        pass
