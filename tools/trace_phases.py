"""Developer tool: per-CTA phase timing of pnr_step_kernel from the -DPNR_TRACE build.

    python -m pioneer_b200.build --trace
    PYTHONPATH=. PIONEER_B200_LIB=pioneer_b200/_lib/libpioneer_b200_trace.so python tools/trace_phases.py [n_envs ...]

Slots (clock64 cycles since CTA entry; slot 0 is %globaltimer at entry): 2 loads issued, 3 before B0 (loads landed,
integrator + sincos r done), 4 after B0, 5 after B1, 6 before B2 (role work done), 7 after B2, 8 bulk store issued,
9 kernel exit (recorded once, in iteration 0's slot).
"""
import ctypes as C
import sys

import numpy as np
import torch

from pioneer_b200 import BatchedPioneerEnv, _cabi

SLOTS, CTAS, ITERS = 12, 2048, 4
NAMES = {2: "loads issued", 3: "pre-B0 (loads landed, integrate, sincos r)", 4: "post-B0", 5: "post-B1",
         6: "pre-B2 (role work done)", 7: "post-B2", 8: "store issued", 9: "exit"}


def main():
    lib = _cabi.load()
    lib.pnr_debug_trace.restype = C.c_int
    lib.pnr_debug_trace.argtypes = [C.c_void_p]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    # floor of the event-pair timing method: a tiny kernel between two events after an L2 flush
    tiny = torch.zeros(32, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    for a, b in ev:
        flush.zero_(); a.record(); tiny.add_(1.0); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    print(f"event-pair floor (1-element torch kernel after flush): median {ts[25] * 1e3:.2f} us, min {ts[0] * 1e3:.2f} us")
    # what the flush costs a 65,536-env step beyond cold DATA: (a) no flush, (b) flush, (c) flush, then one
    # 32-env launch on a scratch handle to re-warm the instruction / constant caches before the timed launch
    big, small = BatchedPioneerEnv(65536), BatchedPioneerEnv(32)
    a_big, a_small = torch.zeros((65536, 6), device="cuda"), torch.zeros((32, 6), device="cuda")
    for mode in ("no flush", "flush", "flush + code re-warm"):
        ts = []
        for k in range(60):
            if mode != "no flush":
                flush.zero_()
            if mode == "flush + code re-warm":
                small.step_tensor(a_small)
            a, b = ev[k % 50]
            a.record(); big.step_tensor(a_big); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[10:])
        print(f"65,536-env step, {mode:22s}: median {ts[len(ts) // 2] * 1e3:.2f} us, min {ts[0] * 1e3:.2f} us")
    big.close(); small.close()
    for n in [int(x) for x in sys.argv[1:]] or [4096, 65536]:
        env = BatchedPioneerEnv(n)
        act = torch.zeros((n, 6), device="cuda")
        for _ in range(5):
            flush.zero_()
            env.step_tensor(act)
        torch.cuda.synchronize()
        buf = np.zeros((CTAS, 2, ITERS, SLOTS), dtype=np.uint64)
        assert lib.pnr_debug_trace(buf.ctypes.data) == 0
        tiles = (n + 31) // 32
        b = buf.astype(np.int64)
        live = b[:, 0, 0, 1] > 0
        n_cta = int(live.sum())
        b = b[live]
        t0 = b[:, :, 0, 0].min()
        print(f"== {n} envs, {tiles} tiles, {n_cta} CTAs traced; CTA start spread (globaltimer) "
              f"{(b[:, 0, 0, 0].max() - t0)} ns; cycles since CTA entry, median / p95 over CTAs")
        for it in range(ITERS):
            has = b[:, 1, it, 7] > b[:, 1, 0, 1]
            if it and not has.any():
                break
            sel = b[has] if it else b
            print(f"  tile iteration {it}: {len(sel)} CTAs")
            for role, name in ((0, "joint warp 0"), (1, "task warp")):
                line = []
                for slot in range(2, 9):
                    if it and slot == 2:
                        continue
                    d = sel[:, role, it, slot] - sel[:, role, 0, 1]
                    line.append(f"{NAMES[slot].split(' (')[0]} {int(np.median(d))}/{int(np.percentile(d, 95))}")
                print(f"     {name:13s} " + " | ".join(line))
        d = b[:, 1, 0, 9] - b[:, 1, 0, 1]
        print(f"  exit (task warp): {int(np.median(d))}/{int(np.percentile(d, 95))} max {int(d.max())}")
        env.close()


if __name__ == "__main__":
    main()
