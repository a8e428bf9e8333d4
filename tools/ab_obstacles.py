"""Developer tool: per-step time of the obstacle variants next to the plain kernels (fragments of 8 steps, pre-aged envs,
L2 flushed between fragments -- bench.py's own timing functions).   python tools/ab_obstacles.py [n_envs ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig, demo_obstacles

dev = torch.device("cuda", 0)
flush_buf = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)


def flush():
    flush_buf.zero_()


def run(name, make, steps=160):
    e = make()
    bench.pre_age(torch, e, 500, 3)
    a, o, r, f = bench.make_buffers(torch, e, e.n_envs, dev, seed=3)
    ms, _ = bench.time_fragments(torch, e, a, o, r, f, steps, 16, flush)
    print(f"{name:44s} n={e.n_envs:8d}  {ms / steps * 1e3:8.2f} us/step  {e.n_envs * steps / ms / 1e6:9.2f} G env-steps/s", flush=True)
    e.close()


for n in [int(x) for x in sys.argv[1:]] or [16384, 65536]:
    ob = demo_obstacles()
    dyn = dict(mode="dynamic", kp=2000.0, kd=500.0, torque_scale=1e5)
    sim = SimulationConfig(gravity=9.81)
    run("kinematic", lambda: BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(max_episode_steps=500)))
    run("kinematic + obstacles", lambda: BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(
        max_episode_steps=500, obstacles=ob, contact_penalty=0.5)))
    run("kinematic + obstacles + random box", lambda: BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(
        max_episode_steps=500, obstacles=ob, contact_penalty=0.5, random_box=True)))
    run("dynamic", lambda: BatchedPioneerEnv(n, seed=0, simulation_config=sim, batch_config=BatchConfig(
        max_episode_steps=500, **dyn)))
    run("dynamic + obstacles", lambda: BatchedPioneerEnv(n, seed=0, simulation_config=sim, batch_config=BatchConfig(
        max_episode_steps=500, obstacles=ob, contact_penalty=0.5, **dyn)))
