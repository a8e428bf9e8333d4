"""pnr_iteration_sync: the once-per-iteration exchange as one kernel over peer memory.  One GPU is enough to exercise the
protocol: two handles on the same device play two ranks, their kernels run concurrently on two streams and store into each
other's windows exactly as two GPUs would over NVLink (tests/test_gpu_multi.py repeats the comparison across real GPUs
and processes).  The checker is the three-kernel path (pnr_stats_device + pnr_filter_delta_device ->
pnr_stats_merge_device -> pnr_filter_sync_device) on twin handles driven with the same actions."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
N, LIMIT = 1000, 3


def _make(seed, base):
    from pioneer_b200 import BatchConfig, BatchedPioneerEnv
    from pioneer_b200.obs_filter import MeanStdObsFilter
    env = BatchedPioneerEnv(N, seed=seed, env_id_base=base, batch_config=BatchConfig(max_episode_steps=LIMIT))
    return env, MeanStdObsFilter(env)


def _advance(env, flt, acts):
    for a in acts:
        obs, _, _ = env.step_tensor(a)
        flt.push(obs)


def _connect(envs):
    from pioneer_b200 import _cabi as c
    lib = envs[0]._lib
    ptrs = (C.c_void_p * len(envs))()
    for r, e in enumerate(envs):
        p = C.c_void_p()
        c.check(lib.pnr_sync_window_ptr(e._h, C.byref(p)))
        ptrs[r] = p.value
    for r, e in enumerate(envs):
        c.check(lib.pnr_sync_window_connect_ptrs(e._h, ptrs, len(envs), r))


def _reference(twins, with_filter, clear):
    """The multi-kernel path on the twin handles: returns merged [len]; merges the filter delta into every twin."""
    from pioneer_b200 import _cabi as c
    lib = twins[0][0]._lib
    ln = 8 + (c.PNR_FILTER_DELTA_LEN if with_filter else 0)
    gathered = torch.zeros((len(twins), ln), dtype=torch.float64, device="cuda")
    for r, (env, _) in enumerate(twins):
        c.check(lib.pnr_stats_device(env._h, gathered[r].data_ptr(), clear, None))
        if with_filter:
            c.check(lib.pnr_filter_delta_device(env._h, gathered[r].data_ptr() + 64, None))
    merged = torch.zeros(ln, dtype=torch.float64, device="cuda")
    c.check(lib.pnr_stats_merge_device(gathered.data_ptr(), len(twins), ln, merged.data_ptr(), None))
    if with_filter:
        for env, _ in twins:
            c.check(lib.pnr_filter_sync_device(env._h, merged.data_ptr() + 64, None))
    torch.cuda.synchronize()
    return merged


@pytest.mark.parametrize("with_filter", [True, False])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_iteration_sync_equals_the_gather_and_merge_path(world, with_filter):
    from pioneer_b200 import _cabi as c
    ranks = [_make(5, r * N) for r in range(world)]
    twins = [_make(5, r * N) for r in range(world)]
    lib = ranks[0][0]._lib
    if world > 1:
        _connect([e for e, _ in ranks])
    streams = [torch.cuda.Stream() for _ in range(world)]
    ln = 8 + (c.PNR_FILTER_DELTA_LEN if with_filter else 0)
    outs = [torch.zeros(ln, dtype=torch.float64, device="cuda") for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(1)
    seen = 0.0
    for it in range(12):                       # more iterations than parities: slots and sequence numbers are reused
        for r in range(world):
            acts = [(torch.rand((N, 6), device="cuda", generator=g) * 2 - 1) * 50 for _ in range(2 + (it + r) % 3)]
            _advance(*ranks[r], acts)
            _advance(*twins[r], acts)
        torch.cuda.synchronize()
        clear = it % 2
        for r in range(world):                 # launch order rotates: whoever comes first waits for the others
            q = (r + it) % world
            with torch.cuda.stream(streams[q]):
                c.check(lib.pnr_iteration_sync(ranks[q][0]._h, int(with_filter), clear, outs[q].data_ptr(), 5000,
                                               streams[q].cuda_stream))
        torch.cuda.synchronize()
        want = _reference(twins, with_filter, clear)
        for r in range(world):
            assert torch.equal(outs[r], want), (it, r)          # bit-identical on every rank
            if with_filter:
                a, b = ranks[r][1], twins[r][1]
                assert a.n == b.n and np.array_equal(a.mean, b.mean) and np.array_equal(a.var, b.var)
            flag = C.c_int(-1)
            c.check(lib.pnr_sync_status(ranks[r][0]._h, C.byref(flag)))
            assert flag.value == 0
        seen += want[0].item()
    assert seen > 0                            # episodes did finish inside the windows
    for e, _ in ranks + twins:
        e.close()


def test_iteration_sync_replays_from_graphs():
    """Captured once per rank, replayed many times: the sequence number lives in the window, not in the launch."""
    from pioneer_b200 import _cabi as c
    ranks = [_make(9, r * N) for r in range(2)]
    twins = [_make(9, r * N) for r in range(2)]
    lib = ranks[0][0]._lib
    _connect([e for e, _ in ranks])
    ln = 8 + c.PNR_FILTER_DELTA_LEN
    outs = [torch.zeros(ln, dtype=torch.float64, device="cuda") for _ in range(2)]
    graphs = []
    torch.cuda.synchronize()
    for r in range(2):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            c.check(lib.pnr_iteration_sync(ranks[r][0]._h, 1, 1, outs[r].data_ptr(), 5000,
                                           torch.cuda.current_stream().cuda_stream))
        graphs.append(gr)
    streams = [torch.cuda.Stream() for _ in range(2)]
    g = torch.Generator(device="cuda").manual_seed(2)
    for it in range(7):
        for r in range(2):
            acts = [(torch.rand((N, 6), device="cuda", generator=g) * 2 - 1) * 50 for _ in range(3)]
            _advance(*ranks[r], acts)
            _advance(*twins[r], acts)
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        want = _reference(twins, True, 1)
        assert torch.equal(outs[0], want) and torch.equal(outs[1], want), it
        assert np.array_equal(ranks[0][1].mean, twins[0][1].mean)
    for e, _ in ranks + twins:
        e.close()


def test_a_missing_peer_times_out_instead_of_hanging():
    from pioneer_b200 import _cabi as c
    ranks = [_make(3, r * N) for r in range(2)]
    lib = ranks[0][0]._lib
    _connect([e for e, _ in ranks])
    _advance(*ranks[0], [torch.zeros((N, 6), device="cuda")] * 4)
    before = ranks[0][1].n
    out = torch.zeros(8 + c.PNR_FILTER_DELTA_LEN, dtype=torch.float64, device="cuda")
    c.check(lib.pnr_iteration_sync(ranks[0][0]._h, 1, 0, out.data_ptr(), 50, None))     # rank 1 never calls
    torch.cuda.synchronize()
    assert torch.isnan(out).all()
    flag = C.c_int(0)
    c.check(lib.pnr_sync_status(ranks[0][0]._h, C.byref(flag)))
    assert flag.value == 1
    assert ranks[0][1].n == before                                  # a poisoned exchange leaves the running statistics alone
    _connect([e for e, _ in ranks])                                 # reconnecting restarts the protocol
    outs = [torch.zeros_like(out) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    for r in range(2):
        with torch.cuda.stream(streams[r]):
            c.check(lib.pnr_iteration_sync(ranks[r][0]._h, 1, 0, outs[r].data_ptr(), 5000, streams[r].cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and not torch.isnan(outs[0]).any()
    c.check(lib.pnr_sync_status(ranks[0][0]._h, C.byref(flag)))
    assert flag.value == 0
    for e, _ in ranks:
        e.close()


def test_connect_rejects_bad_arguments():
    from pioneer_b200 import _cabi as c
    env, _ = _make(1, 0)
    lib = env._lib
    ptrs = (C.c_void_p * 2)(None, None)
    assert lib.pnr_sync_window_connect_ptrs(env._h, ptrs, 2, 0) != 0           # null peer window
    assert lib.pnr_sync_window_connect_ptrs(env._h, ptrs, 17, 0) != 0
    assert lib.pnr_sync_window_connect_ptrs(env._h, ptrs, 2, 2) != 0
    ipc = (C.c_ubyte * c.PNR_SYNC_IPC_BYTES)()
    c.check(lib.pnr_sync_window_create(env._h, ipc))
    assert any(ipc)
    env.close()
