"""Small exercise of every kernel (ragged sizes, auto-reset, both observation modes, dynamic mode, obstacles,
filter) for compute-sanitizer runs:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig, demo_obstacles
from pioneer_b200.obs_filter import MeanStdObsFilter


def run(n, steps, **kw):
    sim = SimulationConfig(gravity=kw.pop("gravity", 0.0))
    env = BatchedPioneerEnv(n, seed=1, simulation_config=sim, batch_config=BatchConfig(max_episode_steps=3, **kw))
    act = (torch.rand((n, 6), device="cuda") - 0.5) * 40
    for _ in range(steps):
        obs, rew, flg = env.step_tensor(act)
    env.observe(indices=[0, n - 1])
    env.reset(indices=[n // 2])
    flt = MeanStdObsFilter(env)
    flt(obs.clone())
    flt.sync()
    torch.cuda.synchronize()
    st = env.episode_stats()
    env.close()
    return st["episodes"]


if __name__ == "__main__":
    total = 0
    for n in (1, 33, 1000, 5000):
        total += run(n, 7)
        total += run(n, 7, obs_mode="autoreset", arith="legacy64")
    total += run(777, 4, mode="dynamic", kp=100.0, kd=10.0, torque_scale=100.0, gravity=9.81)
    total += run(777, 4, obstacles=demo_obstacles(), contact_penalty=0.5)
    print("sanitize smoke ok, episodes", total)
