from .bullet_scene import Pose, Velocity, Item, Joint, Scene, World
from .bullet_env import RenderConfig, SimulationConfig, BulletEnv
