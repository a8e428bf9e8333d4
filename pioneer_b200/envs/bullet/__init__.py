from .bullet_scene import Pose, Item, Joint, Scene, World
from .bullet_env import RenderConfig, SimulationConfig, BulletEnv
