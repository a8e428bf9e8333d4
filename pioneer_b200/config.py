"""Configuration dataclasses with the reference's names and defaults."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np


@dataclass
class RenderConfig:
    """Kept for signature compatibility (reference: pioneer/envs/bullet/bullet_env.py:18-33).
    Rendering is out of scope of the stepping path; render() raises."""
    camera_target: Tuple[float, float, float] = (0, 0, 0)
    camera_distance: float = 100.0
    camera_yaw: float = 120.0
    camera_pitch: float = -30.0
    camera_roll: float = 0.0
    render_width: int = 1280
    render_height: int = 800
    projection_fov: float = 30
    projection_near: float = 0.1
    projection_far: float = 200.0


@dataclass
class SimulationConfig:
    """reference: pioneer/envs/bullet/bullet_env.py:36-62"""
    timestep: float = 1 / 240
    frame_skip: int = 10
    gravity: float = 0
    self_collision: bool = False
    collision_parent: bool = True

    @property
    def frames_per_second(self) -> int:
        return int(np.round(1 / (self.timestep * self.frame_skip)))


@dataclass
class PioneerKinematicConfig:
    """reference: pioneer/envs/pioneer/pioneer_knm_env.py:19-34"""
    max_v_to_r: float = 2       # seconds^-1
    max_a_to_v: float = 10      # seconds^-1

    done_distance: float = 0.1

    award_max: float = 100.0
    award_done: float = 5.0
    award_potential_slope: float = 10.0
    penalty_step: float = 1 / 100

    target_lo: Tuple[float, float, float] = (15, -10, 2)
    target_hi: Tuple[float, float, float] = (25, 10, 6)
    target_radius: float = 0.2
    target_rgba: Tuple[float, float, float, float] = (1.0, 0.0, 0.0, 0.5)


@dataclass
class Obstacle:
    """Static obstacle of the reach-with-obstacles variant.  kind: 'plane' (position = point on the
    plane, extent = unit normal), 'box' (position = centre, extent = half extents, axis aligned; the
    reference demo's box is half extents (0.5, 0.5, 5) at (10, 5, 0), pioneer_knm_env.py:249-255),
    'sphere' (extent[0] = radius)."""
    kind: str
    position: Tuple[float, float, float]
    extent: Tuple[float, float, float]


def demo_obstacles() -> List[Obstacle]:
    """The only obstacle geometry the reference instantiates (pioneer_knm_env.py:249-261) plus one sphere."""
    return [Obstacle("plane", (0, 0, 0), (0, 0, 1.0)),
            Obstacle("box", (10, 5, 0), (0.5, 0.5, 5.0)),
            Obstacle("sphere", (12.0, -6.0, 8.0), (2.0, 0, 0))]


@dataclass
class BatchConfig:
    """Knobs of the batched GPU implementation that the reference has no counterpart for."""
    max_episode_steps: int = 500          # gym TimeLimit in the reference launcher (pioneer_knm_train.py:27); 0 = off
    auto_reset: bool = True               # finished envs are reset inside the step kernel
    obs_mode: str = "terminal"            # 'terminal' (what BulletEnv.step returns) | 'autoreset'
    arith: str = "f32"                    # 'f32' | 'legacy64' (NumPy 1.x promotion of the integrator)
    mode: str = "kinematic"               # 'kinematic' (the reference env) | 'dynamic' (ABA + PD, Tier B)
    kp: float = 0.0                       # dynamic mode: tau = kp (q_des - q) - kd qd, clamped to +-effort*torque_scale
    kd: float = 0.0
    torque_scale: float = 1.0
    obstacles: List[Obstacle] = field(default_factory=list)
    contact_penalty: float = 0.0
    # per-env random box (the legacy randomizer, pioneer/temp/pioneer_env.py:169-192): the first box among `obstacles` is
    # redrawn at every reset -- half extents ~ U(box_size_lo, box_size_hi), centre = (U(box_pos_lo, box_pos_hi), half height)
    random_box: bool = False
    box_pos_lo: Tuple[float, float] = (8.0, -6.0)
    box_pos_hi: Tuple[float, float] = (14.0, 6.0)
    box_size_lo: Tuple[float, float, float] = (0.3, 0.3, 3.0)
    box_size_hi: Tuple[float, float, float] = (0.7, 0.7, 7.0)
    # dynamic mode: 'explicit'  | 'bullet' (opt-in Bullet-like substep: per-link damping, +-max_velocity
    # clamp, POSITION_CONTROL as a velocity-level motor constraint with impulse clamp motor_max_force * dt;
    # Joint.control_position, bullet_scene.py:123-142)
    stepping: str = "explicit"
    link_damping: float = 0.04
    max_velocity: float = 100.0
    motor_kp: float = 0.1                 # setJointMotorControl2 positionGain
    motor_kd: float = 1.0                 # velocityGain
    motor_max_force: float = 0.0          # `force`; 0 = motors off
