"""pioneer_b200: the Pioneer 6-DoF reach environment stepped on NVIDIA B200 (sm_100a).

Importing this package does not open the CUDA library; constructing an env does, and raises if it is
missing (there is no CPU fallback)."""
from .config import BatchConfig, Obstacle, PioneerKinematicConfig, RenderConfig, SimulationConfig, demo_obstacles

__all__ = ["BatchConfig", "Obstacle", "PioneerKinematicConfig", "RenderConfig", "SimulationConfig", "demo_obstacles",
           "BatchedPioneerEnv", "PioneerKinematicEnv", "PioneerVectorEnv"]


def __getattr__(name):
    if name == "BatchedPioneerEnv":
        from .batched_env import BatchedPioneerEnv
        return BatchedPioneerEnv
    if name == "PioneerKinematicEnv":
        from .envs.pioneer import PioneerKinematicEnv
        return PioneerKinematicEnv
    if name == "PioneerVectorEnv":
        from .vector_env import PioneerVectorEnv
        return PioneerVectorEnv
    raise AttributeError(name)
