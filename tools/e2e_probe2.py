"""Developer tool: host-buffer step variants (sync / pipelined, full / compact rows, one or two copy streams)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pioneer_b200 import BatchConfig, BatchedPioneerEnv

n, steps = 65536, 200
env = BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(max_episode_steps=500))
acts = [torch.empty((n, 6), dtype=torch.float32, pin_memory=True).uniform_(-50, 50) for _ in range(4)]


def sync_full():
    for k in range(steps):
        env.step_host(acts[k % 4])


def begin_end(compact, depth):
    def run():
        if depth == 2:
            env.step_host_begin(acts[0], compact=compact)
        for k in range(steps):
            if depth == 2:
                if k + 1 < steps:
                    env.step_host_begin(acts[(k + 1) % 4], compact=compact)
            else:
                env.step_host_begin(acts[k % 4], compact=compact)
            env.step_host_end()
    return run


for name, fn in (("sync full (pnr_step_host)", sync_full), ("begin/end full depth1", begin_end(False, 1)),
                 ("begin/end full depth2", begin_end(False, 2)), ("begin/end compact depth1", begin_end(True, 1)),
                 ("begin/end compact depth2", begin_end(True, 2))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"{name:28s} {dt * 1e3:7.3f} ms/step  {n / dt:10.3e} env-steps/s", flush=True)
