"""Multi-GPU host logic on CPU: env sharding and the one collective of the path (episode-statistics all-reduce),
world size 2 over gloo.  The data path has no collective."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from pioneer_b200.distributed import reduce_episode_stats, shard_range, summarize  # noqa: E402


def test_shard_range_covers_every_env_once():
    for total, world in [(65536 * 8, 8), (1000, 3), (7, 7), (1 << 20, 8), (10, 4)]:
        seen = []
        for rank in range(world):
            base, n = shard_range(total, rank, world)
            seen.extend(range(base, base + n))
            assert n in (total // world, total // world + 1)
        assert seen == list(range(total))


def test_summarize_matches_numpy():
    rng = np.random.default_rng(0)
    returns, lengths = rng.normal(20, 5, size=1000), rng.integers(1, 500, size=1000)
    packed = torch.tensor([1000, returns.sum(), lengths.sum(), (returns ** 2).sum(), returns.max(), returns.min(),
                           lengths.sum(), 17], dtype=torch.float64)
    s = summarize(packed)
    assert s["episodes_total"] == 1000 and s["reached_target"] == 17
    np.testing.assert_allclose(s["episode_reward_mean"], returns.mean())
    np.testing.assert_allclose(s["episode_reward_std"], returns.std(), rtol=1e-9)
    np.testing.assert_allclose(s["episode_len_mean"], lengths.mean())
    assert s["episode_reward_max"] == returns.max() and s["episode_reward_min"] == returns.min()
    empty = summarize(torch.tensor([0, 0, 0, 0, -np.inf, np.inf, 0, 0], dtype=torch.float64))
    assert empty["episodes_total"] == 0 and empty["episode_reward_mean"] is None and empty["episode_reward_max"] is None
    import json
    json.loads(json.dumps(empty, allow_nan=False))                 # strict JSON


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # per-rank packed statistics as pnr_stats_device lays them out
        local = torch.tensor([[3, 30.0, 900, 350.0, 15.0, 5.0, 4096, 1],
                              [5, 80.0, 1100, 1500.0, 40.0, -2.0, 4096, 0]], dtype=torch.float64)[rank]
        total = reduce_episode_stats(local)
        out.put((rank, total.tolist()))
        # a rank that finished no episode contributes the identity (-inf / +inf) to max / min
        idle = torch.tensor([0, 0, 0, 0, -np.inf, np.inf, 10, 0], dtype=torch.float64)
        mixed = reduce_episode_stats(local if rank == 0 else idle)
        out.put((rank + 10, mixed.tolist()))
        # the rollout worker's ONE collective per iteration: statistics + an additive vector, all-gathered and reduced locally
        from pioneer_b200.distributed import reduce_packed
        extra = torch.arange(5, dtype=torch.float64) * (rank + 1)
        st, ex = reduce_packed(local, extra)
        out.put((rank + 20, st.tolist() + ex.tolist()))
    finally:
        dist.destroy_process_group()


def test_reduce_episode_stats_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in range(6))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [8, 110.0, 2000, 1850.0, 40.0, -2.0, 8192, 1]
    assert got[0] == want and got[1] == want                      # SUM over counters, MAX / MIN over returns
    want_mixed = [3, 30.0, 900, 350.0, 15.0, 5.0, 4106, 1]
    assert got[10] == want_mixed and got[11] == want_mixed
    want_packed = want + [0.0, 3.0, 6.0, 9.0, 12.0]
    assert got[20] == want_packed and got[21] == want_packed      # the packed all-gather gives the same statistics
    s = summarize(torch.tensor(got[0], dtype=torch.float64))
    assert s["episodes_total"] == 8 and s["episode_reward_mean"] == 110.0 / 8


def test_single_process_reduce_is_identity():
    x = torch.arange(8, dtype=torch.float64)
    assert torch.equal(reduce_episode_stats(x), x)
