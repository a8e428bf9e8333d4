"""Developer probe: the step with and without the fused observation normaliser and its statistics (the measurement behind
the PNR_FILTER_SLOTS comment in pioneer_b200/csrc/pnr_launch.h).   python tools/fuse_probe.py"""
import sys, os; sys.path.insert(0, os.getcwd())
import torch
from pioneer_b200 import BatchedPioneerEnv, BatchConfig
from pioneer_b200.obs_filter import MeanStdObsFilter
for n in (65536, 1048576):
    env = BatchedPioneerEnv(n, seed=0, batch_config=BatchConfig(max_episode_steps=500))
    flt = MeanStdObsFilter(env)
    a_max = torch.as_tensor(env.a_max, device="cuda")
    acts = (torch.rand((8, n, 6), device="cuda") * 2 - 1) * a_max
    ring = torch.empty((max(2, min(8, (1 << 30) // (n * 137 * 4))), n, 137), device="cuda")
    rew = torch.empty(n, device="cuda"); flg = torch.empty(n, dtype=torch.uint8, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    def run(k=200):
        evs = []
        for i in range(k + 10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.step_tensor(acts[i % 8], out=(ring[i % ring.shape[0]], rew, flg)); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = sorted(x.elapsed_time(y) for x, y in evs[10:])
        return t[len(t) // 2] * 1e3
    print(n, "plain", run())
    flt.set_fused(True, update=False)
    print(n, "fused, no statistics", run())
    flt.set_fused(True, update=True)
    print(n, "fused + statistics", run())
    flt.set_fused(False)
    env.close()
