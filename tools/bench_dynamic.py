"""Developer tool: time the Tier-B dynamic-mode step kernel (ABA + PD, 10 substeps) with CUDA events.

    PYTHONPATH=. python tools/bench_dynamic.py [--envs N] [--steps K] [--obstacles]
"""
import argparse

import torch

from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig, demo_obstacles


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--obstacles", action="store_true")
    ap.add_argument("--fused-filter", action="store_true", help="normalise the observations inside the step kernel")
    a = ap.parse_args()
    bc = BatchConfig(mode="dynamic", kp=2000.0, kd=500.0, torque_scale=1e5, max_episode_steps=500,
                     obstacles=demo_obstacles() if a.obstacles else [], contact_penalty=0.5 if a.obstacles else 0.0)
    env = BatchedPioneerEnv(a.envs, seed=0, simulation_config=SimulationConfig(gravity=9.81), batch_config=bc)
    if a.fused_filter:
        from pioneer_b200.obs_filter import MeanStdObsFilter
        flt = MeanStdObsFilter(env, fused=True)                     # noqa: F841 (keeps the fused mode on)
    lo, hi = torch.as_tensor(env.r_lo).cuda(), torch.as_tensor(env.r_hi).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((8, a.envs, 6), device="cuda", generator=g) * (hi - lo) + lo
    for k in range(5):
        env.step_tensor(acts[k % 8])
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for k in range(a.steps):
        env.step_tensor(acts[k % 8])
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.steps
    print(f"dynamic mode{' + fused normaliser' if a.fused_filter else ''}: {a.envs} envs, {ms * 1e3:.1f} us/step, {a.envs / ms * 1e3:.3e} env-steps/s, "
          f"{10 * a.envs / ms * 1e3:.3e} substeps/s")
    env.close()


if __name__ == "__main__":
    main()
