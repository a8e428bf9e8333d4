"""pioneer.envs.bullet (reference: pioneer/envs/bullet/__init__.py:1-2) -> pioneer_b200.envs.bullet"""
from pioneer_b200.envs.bullet import (BulletEnv, Item, Joint, Pose, RenderConfig, Scene, SimulationConfig, Velocity,  # noqa: F401
                                      World)
