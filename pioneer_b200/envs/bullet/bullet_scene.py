"""Read-only mirrors of the reference scene wrappers (pioneer/envs/bullet/bullet_scene.py) for ONE env
of a BatchedPioneerEnv.  They carry no physics client: every query is answered from the device state
(Joint.position/velocity -> pnr_get_state; Item.pose of the tracked link -> the observation's
pointer/target columns).  Only what the reach env touches is provided."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np


class Pose:
    """reference: bullet_scene.py:11-30 (position + quaternion xyzw)."""

    def __init__(self, xyz: Tuple[float, float, float], quaternion: Tuple[float, float, float, float] = (0, 0, 0, 1)):
        self.xyz = tuple(float(x) for x in xyz)
        self.quaternion = tuple(float(x) for x in quaternion)

    def __repr__(self):
        return f"Pose(xyz={self.xyz}, quaternion={self.quaternion})"


class Velocity:
    """reference: bullet_scene.py:33-36 (linear, angular); part of the import surface, unused on the stepping path."""

    def __init__(self, linear: Tuple[float, float, float] = (0, 0, 0), angular: Tuple[float, float, float] = (0, 0, 0)):
        self.linear = tuple(float(x) for x in linear)
        self.angular = tuple(float(x) for x in angular)


class Item:
    """A named body or link (reference: bullet_scene.py:39-73).  Only 'robot:pointer' and 'target'
    have a pose on this path (pioneer_knm_env.py:151-152)."""

    def __init__(self, env, name: str, obs_slice: Optional[slice]):
        self._env, self.name, self._slice = env, name, obs_slice

    def pose(self) -> Pose:
        if self._slice is None:
            raise NotImplementedError(f"pose() of {self.name!r} is not on the reach path")
        obs = self._env._batch.observe(indices=[self._env._index])[0].cpu().numpy()
        return Pose(obs[self._slice])


class Joint:
    """reference: bullet_scene.py:76-165."""

    def __init__(self, env, k: int, name: str, lower: float, upper: float, effort: float, damping: float,
                 friction: float, max_velocity: float):
        self._env, self._k, self.name = env, k, name
        self.joint_index = k
        self.lower_limit, self.upper_limit = lower, upper
        self.max_force, self.damping, self.friction, self.max_velocity = effort, damping, friction, max_velocity

    def __repr__(self):
        return f"Joint(name={self.name}, lower_limit={self.lower_limit}, upper_limit={self.upper_limit})"

    def position(self) -> float:
        return float(self._env._batch.state()["r"][self._env._index, self._k])

    def velocity(self) -> float:
        """Bullet's joint velocity: resetJointState zeroes it and nothing sets it again in kinematic
        mode; in dynamic mode it is the integrated joint rate."""
        if self._env._batch.batch_config.mode == "dynamic":
            return float(self._env._batch.state()["v"][self._env._index, self._k])
        return 0.0

    def reset_state(self, position: float, velocity: Optional[float] = None):
        s = self._env._batch.state()
        s["r"][self._env._index, self._k] = float(position)
        self._env._batch.set_state(r=s["r"])


class Scene:
    """name -> item / joint registries (reference: bullet_scene.py:168-259)."""

    def __init__(self, env, chain):
        self.joints: List[Joint] = []
        self.joints_by_name: Dict[str, Joint] = {}
        self.items_by_name: Dict[str, Item] = {}
        for k, name in enumerate(chain.joint_names):
            j = Joint(env, k, name, float(chain.lower[k]), float(chain.upper[k]), float(chain.effort[k]),
                      float(chain.damping[k]), float(chain.friction[k]), float(chain.max_velocity[k]))
            self.joints.append(j)
            self.joints_by_name[name] = j
        for link in chain.links:
            self.items_by_name[link] = Item(env, link, slice(126, 129) if link == chain.tip_name else None)
        self.items_by_name["target"] = Item(env, "target", slice(129, 132))

    @staticmethod
    def rpy2quat(rpy):
        r, p, y = (0.5 * float(v) for v in rpy)
        cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
        return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                cr * cp * cy + sr * sp * sy)


class World:
    """reference: bullet_scene.py:262-279.  step() is what the fused kernel already did."""

    def __init__(self, timestep: float, frame_skip: int, gravity: float):
        self.timestep, self.frame_skip, self.gravity_force = timestep, frame_skip, gravity

    @property
    def step_time(self) -> float:
        return self.timestep * self.frame_skip
