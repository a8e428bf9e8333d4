"""BASELINE.json configs[4]: rollout loop (observation filter -> 137-256-256-12 policy stub -> fused env step) on the
tensor fast path, env-sharded over the ranks, one episode-stat + filter-stat all-reduce per iteration.

    python tools/bench_rollout.py [--envs-per-gpu 131072] [--fragment 8] [--iters 10]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_rollout.py
Prints one JSON line on rank 0 (device time from CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from pioneer_b200 import BatchConfig, BatchedPioneerEnv
from pioneer_b200.rollout import RolloutWorker


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--fragment", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-filter", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay each fragment from a CUDA graph")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    sys.stdout.flush()
    json_fd = os.dup(1)                 # stdout carries exactly one JSON line: NCCL prints its banner on fd 1
    os.dup2(2, 1)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.envs_per_gpu
    env = BatchedPioneerEnv(n, device=dev, seed=0, env_id_base=rank * n, batch_config=BatchConfig(max_episode_steps=500))
    torch.cuda.manual_seed(1234 + rank)
    worker = RolloutWorker(env, fragment_length=a.fragment, use_filter=not a.no_filter, seed=rank, cuda_graph=a.graph)
    for _ in range(3):
        worker.collect()
        worker.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        batch = worker.collect()
        stats = worker.sync(summary=False)         # stays on the device: the loop never waits for the GPU
    e.record()
    torch.cuda.synchronize()
    from pioneer_b200.distributed import summarize
    summary = summarize(stats)
    ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    steps = a.iters * a.fragment
    line = {"metric": "rollout env-steps/sec (filter + 137-256-256-12 policy stub + env step, device resident)",
            "value": world * n * steps / (float(ms) / 1e3), "unit": "env-steps/s", "n_gpus": world, "envs_per_gpu": n,
            "total_envs": world * n, "fragment_length": a.fragment, "iterations": a.iters,
            "ms_per_env_step_batch": float(ms) / steps, "filter": not a.no_filter, "cuda_graph": a.graph,
            "batch_shapes": {k: list(v.shape) for k, v in batch.items()}, "episode_stats": summary}
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())        # summarize() reports None (null), never NaN
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
