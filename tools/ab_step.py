"""Developer tool: A/B two builds of libpioneer_b200 on the same box (only the entry points every ABI version has).

    python tools/ab_step.py libA.so libB.so [n_envs ...]
Times pnr_step with per-step CUDA events and an L2 flush between steps, interleaving the two libraries."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pioneer_b200 import _cabi
from pioneer_b200.urdf import flatten_urdf


def open_lib(path):
    lib = C.CDLL(path)
    for name in ("pnr_default_config", "pnr_create", "pnr_destroy", "pnr_step", "pnr_last_error"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _cabi.SIGNATURES[name]
    return lib


def make(lib, n):
    model = _cabi.model_from_chain(flatten_urdf())
    cfg = _cabi.pnr_config()
    lib.pnr_default_config(C.byref(cfg))
    h = C.c_void_p()
    rc = lib.pnr_create(C.byref(model), C.byref(cfg), n, 0, 0, 0, C.byref(h))
    assert rc == 0, lib.pnr_last_error()
    return h


def main():
    paths = sys.argv[1:3]
    sizes = [int(x) for x in sys.argv[3:]] or [65536]
    libs = [open_lib(p) for p in paths]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for n in sizes:
        hs = [make(lib, n) for lib in libs]
        act = (torch.rand((8, n, 6), device="cuda") * 2 - 1) * 50
        obs = torch.empty((4, n, 137), device="cuda")
        rew = torch.empty(n, device="cuda")
        flg = torch.empty(n, dtype=torch.uint8, device="cuda")
        times = [[], []]
        for k in range(640):
            i = k & 1
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = libs[i].pnr_step(hs[i], act[k % 8].data_ptr(), obs[k % 4].data_ptr(), rew.data_ptr(), flg.data_ptr(), stream)
            b.record()
            assert rc == 0
            times[i].append((a, b))
        torch.cuda.synchronize()
        for i in range(2):
            ts = sorted(a.elapsed_time(b) for a, b in times[i][20:])
            print(f"{n:9d} envs  {os.path.basename(paths[i]):28s} mean {sum(ts) / len(ts) * 1e3:7.2f} us  p50 {ts[len(ts) // 2] * 1e3:7.2f} us")
        for lib, h in zip(libs, hs):
            lib.pnr_destroy(h)


if __name__ == "__main__":
    main()
