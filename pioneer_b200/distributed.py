"""Multi-GPU sharding of the env batch and the one collective the path has.

Envs are independent: a job of ``total_envs`` is cut into contiguous ranges of global env ids, one
per rank / GPU, with no data-path communication.  Once per training iteration the per-GPU episode
statistics (the columns the reference CLI prints, cli.py:32-38) are reduced with NCCL over
NVLink: SUM over the counters and MAX over (max_return, -min_return).  The reference moves the same
numbers from its Ray rollout workers to the trainer (pioneer/launch/pioneer_knm_train.py:49).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .batched_env import STATS_FIELDS

SUM_SLOTS = (0, 1, 2, 3, 6, 7)
MAX_SLOTS = (4, 5)


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host buffer is allocated, so
    that the host side of the H2D / D2H copies (pnr_step_host) is local memory.  One rank per GPU on a two-socket box
    otherwise lands its pinned buffers wherever the scheduler started it.  Returns the node, or None if unknown."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001 - sysfs layout / permissions differ between boxes: binding is best effort
        return None


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(env_id_base, n_local): contiguous split, the first ``total % world`` ranks get one more."""
    assert 0 <= rank < world_size and total_envs >= world_size
    q, r = divmod(total_envs, world_size)
    n_local = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, n_local


def reduce_episode_stats(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-reduce the packed float64[8] statistics vector (any backend; NCCL on the GPUs).
    Two collectives: SUM and MAX.  Returns a new tensor on the same device."""
    assert local.dtype == torch.float64 and local.numel() == len(STATS_FIELDS)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.clone()
    sums = local[list(SUM_SLOTS)].contiguous()
    maxs = torch.stack([local[4], -local[5]])
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    out = torch.empty_like(local)
    out[list(SUM_SLOTS)] = sums
    out[4], out[5] = maxs[0], -maxs[1]
    return out


def reduce_packed(stats_local: torch.Tensor, extra_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """ONE collective for everything a rollout worker exchanges per iteration: the float64[8] episode statistics (SUM /
    MAX slots) and an additive float64 vector (the observation filter's delta) are packed, all-gathered, and reduced
    locally.  Returns (stats, extra) like reduce_episode_stats + a SUM all-reduce would."""
    assert stats_local.dtype == torch.float64 and extra_local.dtype == torch.float64
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats_local.clone(), extra_local.clone()
    world = dist.get_world_size(group)
    packed = torch.cat([stats_local, extra_local])
    gathered = torch.empty((world, packed.numel()), dtype=torch.float64, device=packed.device)
    dist.all_gather(list(gathered.unbind(0)), packed, group=group)       # (works on gloo as well as on NCCL)
    k = len(STATS_FIELDS)
    stats = gathered[:, :k].sum(0)
    stats[4] = gathered[:, 4].max()
    stats[5] = gathered[:, 5].min()
    return stats, gathered[:, k:].sum(0).contiguous()


def summarize(stats: torch.Tensor) -> Dict[str, float]:
    """episode_reward_max/min/mean, episode_len_mean, episodes_total (cli.py:32-38) from the packed vector."""
    s = [float(x) for x in stats.tolist()]
    n = s[0]
    if not n:      # no episode finished in the window: None (JSON null), never NaN / inf
        return {"episodes_total": 0.0, "episode_reward_mean": None, "episode_reward_std": None, "episode_reward_max": None,
                "episode_reward_min": None, "episode_len_mean": None, "env_steps": s[6], "reached_target": s[7]}
    mean = s[1] / n
    var = max(s[3] / n - mean * mean, 0.0)
    return {"episodes_total": n, "episode_reward_mean": mean, "episode_reward_std": var ** 0.5,
            "episode_reward_max": s[4], "episode_reward_min": s[5], "episode_len_mean": s[2] / n, "env_steps": s[6],
            "reached_target": s[7]}
