"""Developer tool: A/B several builds of libpioneer_b200 on the dynamic-mode step (one subprocess per library, selected
with PIONEER_B200_LIB), timed like bench.py: fragments of 8 steps (pnr_step_many) between L2 flushes, CUDA events.

    python tools/ab_dynamic.py libA.so libB.so ... [--sizes 65536 1048576] [--kinematic]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, sys, torch
sys.path.insert(0, %(root)r)
from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig
sizes, kinematic, age = %(sizes)r, %(kinematic)r, %(age)r
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = {}
for n in sizes:
    if kinematic:
        env = BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(max_episode_steps=500))
    else:
        env = BatchedPioneerEnv(n, seed=0, simulation_config=SimulationConfig(gravity=9.81),
                                batch_config=BatchConfig(mode="dynamic", kp=2000.0, kd=500.0, torque_scale=1e5, max_episode_steps=500))
    g = torch.Generator(device="cuda").manual_seed(0)
    if age:
        env.set_state(t=torch.randint(0, 500, (n,), device="cuda", generator=g, dtype=torch.int32))
    lo, hi = torch.as_tensor(env.action_space.low).cuda(), torch.as_tensor(env.action_space.high).cuda()
    T = 8 if n <= 131072 else 2
    acts = lo + torch.rand((T, n, 6), device="cuda", generator=g) * (hi - lo)
    obs = torch.empty((T, n, 137), device="cuda")
    rew = torch.empty((T, n), device="cuda"); flg = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        env.step_many(acts, obs, rew, flg)
    evs = []
    for k in range(40 if n <= 131072 else 16):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step_many(acts, obs, rew, flg); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) / T for a, b in evs)
    out[n] = {"p50_us": ts[len(ts) // 2] * 1e3, "mean_us": sum(ts) / len(ts) * 1e3}
    env.close()
print("AB " + json.dumps(out))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--sizes", type=int, nargs="+", default=[65536, 1048576])
    ap.add_argument("--kinematic", action="store_true")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--no-age", action="store_true", help="all envs start at age 0: no auto-reset inside the timed steps")
    a = ap.parse_args()
    code = CHILD % dict(root=ROOT, sizes=a.sizes, kinematic=a.kinematic, age=not a.no_age)
    for rnd in range(a.rounds):
        for lib in a.libs:
            env = dict(os.environ, PIONEER_B200_LIB=os.path.abspath(lib))
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            line = next((ln for ln in out.stdout.splitlines() if ln.startswith("AB ")), None)
            if line is None:
                print(f"{os.path.basename(lib):32s} FAILED: {out.stderr[-600:]}")
                continue
            res = json.loads(line[3:])
            print(f"round {rnd} {os.path.basename(lib):32s} " + "  ".join(f"{n}: {v['p50_us']:8.2f} us" for n, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
