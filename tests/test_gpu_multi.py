"""Two ranks on two GPUs over NCCL: env sharding by global id, the episode-statistics all-reduce and the filter
synchronisation reproduce a single-GPU run of the same job.  Skipped on boxes with one GPU (the CPU suite covers the
host logic with gloo, tests/test_distributed.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu
N_PER, STEPS, LIMIT, SEED = 1000, 6, 3, 13


def _actions(n_total):
    g = torch.Generator().manual_seed(3)
    return (torch.rand((STEPS, n_total, 6), generator=g) * 2 - 1) * 50.0


def _run(env, flt, acts, fused_sync=False):
    from pioneer_b200.distributed import reduce_episode_stats
    outs = []
    for t in range(STEPS):
        obs, rew, flg = env.step_tensor(acts[t].to(env.device).contiguous())
        flt.push(obs)
        outs.append((obs.clone(), rew.clone(), flg.clone()))
    stats = reduce_episode_stats(env.episode_stats_tensor())
    # the same reduction as ONE graph-replayed collective with static buffers (what the rollout loop and bench.py use)
    # the same reduction with static buffers (what the rollout loop and bench.py use): over NCCL, and as ONE kernel per rank
    # over peer memory (eager and replayed from a CUDA graph) -- across real GPUs and processes here
    from pioneer_b200.distributed import IterationSync
    world = dist.get_world_size() if dist.is_initialized() else 1
    for transport, graph in (("nccl", False), ("p2p", False), ("p2p", True), ("auto", True)):
        sync = IterationSync(env, None, None, clear=False, cuda_graph=graph, transport=transport)
        if world > 1 and transport != "nccl":
            assert sync.transport == "p2p", "the ranks of one node must be able to open each other's windows"
        for _ in range(3):
            again = sync()
            assert torch.equal(again, stats), (transport, graph)
        assert not sync.timed_out()
    if fused_sync:      # statistics + filter delta in the one-kernel exchange (what RolloutWorker.sync does)
        both = IterationSync(env, flt, None, clear=False, cuda_graph=True)
        assert torch.equal(both(), stats) and not both.timed_out()
    else:
        flt.sync()
    return outs, stats.cpu().numpy(), flt.n, flt.mean, flt.var


def _worker(rank, world, port, q, fused_sync):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))      # NCCL_DEBUG is left as the caller set it
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pioneer_b200 import BatchConfig, BatchedPioneerEnv
        from pioneer_b200.obs_filter import MeanStdObsFilter
        env = BatchedPioneerEnv(N_PER, device=torch.device("cuda", rank), seed=SEED, env_id_base=rank * N_PER,
                                batch_config=BatchConfig(max_episode_steps=LIMIT))
        flt = MeanStdObsFilter(env)
        acts = _actions(world * N_PER)[:, rank * N_PER:(rank + 1) * N_PER]
        outs, stats, n, mean, var = _run(env, flt, acts, fused_sync)
        q.put((rank, [o[0].cpu().numpy() for o in outs], stats, n, mean, var))
        env.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fused_sync", [False, True])
def test_two_ranks_equal_one_gpu(fused_sync):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, fused_sync)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, obs, stats, n, mean, var = q.get(timeout=300)
        got[rank] = (obs, stats, n, mean, var)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the same job on one GPU
    from pioneer_b200 import BatchConfig, BatchedPioneerEnv
    from pioneer_b200.obs_filter import MeanStdObsFilter
    env = BatchedPioneerEnv(2 * N_PER, seed=SEED, batch_config=BatchConfig(max_episode_steps=LIMIT))
    flt = MeanStdObsFilter(env)
    outs, stats, n, mean, var = _run(env, flt, _actions(2 * N_PER), fused_sync)
    for t in range(STEPS):
        whole = outs[t][0].cpu().numpy()
        assert np.array_equal(whole[:N_PER], got[0][0][t]) and np.array_equal(whole[N_PER:], got[1][0][t]), t
    for r in (0, 1):                                  # every rank holds the job-wide statistics after the all-reduce
        np.testing.assert_allclose(got[r][1], stats, rtol=1e-6)     # float32 per-tile partial sums, float64 atomics
        assert got[r][2] == n == 2 * N_PER * STEPS
        np.testing.assert_allclose(got[r][3], mean, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(got[r][4], var, rtol=1e-7, atol=1e-10)
    assert stats[0] == 2 * N_PER * (STEPS // LIMIT)
    env.close()
