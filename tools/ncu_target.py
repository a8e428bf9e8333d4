"""Developer tool: a short, deterministic launch sequence for ncu captures of the two step kernels.

    python tools/ncu_target.py [kinematic|dynamic|obstacles] [n_envs] [fragments]
Runs `fragments` rollout fragments of 8 steps (pnr_step_many) on pre-aged envs, an L2 flush between fragments, exactly as
bench.py times them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig

mode = sys.argv[1] if len(sys.argv) > 1 else "kinematic"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
frags = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if mode == "dynamic":
    env = BatchedPioneerEnv(n, seed=0, simulation_config=SimulationConfig(gravity=9.81),
                            batch_config=BatchConfig(mode="dynamic", kp=2000.0, kd=500.0, torque_scale=1e5, max_episode_steps=500))
elif mode == "obstacles":
    from pioneer_b200 import demo_obstacles
    env = BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(max_episode_steps=500, obstacles=demo_obstacles(),
                                                               contact_penalty=0.5))
else:
    env = BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(max_episode_steps=500))
g = torch.Generator(device="cuda").manual_seed(0)
env.set_state(t=torch.randint(0, 500, (n,), device="cuda", generator=g, dtype=torch.int32))
lo, hi = torch.as_tensor(env.action_space.low).cuda(), torch.as_tensor(env.action_space.high).cuda()
T = 8 if n <= 131072 else 2
acts = lo + torch.rand((T, n, 6), device="cuda", generator=g) * (hi - lo)
obs = torch.empty((T, n, 137), device="cuda")
rew = torch.empty((T, n), device="cuda")
flg = torch.empty((T, n), dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(frags):
    flush.zero_()
    env.step_many(acts, obs, rew, flg)
torch.cuda.synchronize()
import ctypes
env._lib.pnr_source_hash.restype = ctypes.c_char_p
print("ok", mode, n, env.episode_stats()["env_steps"], env._lib.pnr_source_hash().decode())
env.close()
