"""Developer probe: is a pitched (2D) device->host copy of only the CHANGING observation columns faster than the
contiguous copy of all 137?  Columns 18:54 of every row are per-handle constants (pioneer_knm_env.py:190-198: r_lo,
cos r_lo, sin r_lo, r_hi, cos r_hi, sin r_hi), so a host buffer that already holds them needs only columns 0:18
(72 B per row) and 54:137 (332 B per row).  Prints GB/s of useful bytes and ms per 65,536-row batch for each variant."""
import ctypes
import sys
import time

import torch

N, W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 137
rt = ctypes.CDLL("libcudart.so.12")
D2H = 2
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

dev = torch.randn((N, W), device="cuda")
host = torch.empty((N, W), dtype=torch.float32, pin_memory=True)
compact_dev = torch.randn((N, 101), device="cuda")
compact_host = torch.empty((N, 101), dtype=torch.float32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def chk(rc):
    assert rc == 0, rc


def full():
    chk(rt.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), N * W * 4, D2H, s1.cuda_stream))


def full_two_engines():
    half = (N // 2) * W * 4
    chk(rt.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), half, D2H, s1.cuda_stream))
    chk(rt.cudaMemcpyAsync(host.data_ptr() + half, dev.data_ptr() + half, N * W * 4 - half, D2H, s2.cuda_stream))


def pitched(stream_a, stream_b):
    chk(rt.cudaMemcpy2DAsync(host.data_ptr(), W * 4, dev.data_ptr(), W * 4, 18 * 4, N, D2H, stream_a.cuda_stream))
    chk(rt.cudaMemcpy2DAsync(host.data_ptr() + 54 * 4, W * 4, dev.data_ptr() + 54 * 4, W * 4, 83 * 4, N, D2H,
                             stream_b.cuda_stream))


def pitched_one():
    pitched(s1, s1)


def pitched_two():
    pitched(s1, s2)


def compact_contiguous():
    # lower bound for any scheme that moves 101 columns: the same bytes as one contiguous copy
    chk(rt.cudaMemcpyAsync(compact_host.data_ptr(), compact_dev.data_ptr(), N * 101 * 4, D2H, s1.cuda_stream))


def pitched_wide_only():
    chk(rt.cudaMemcpy2DAsync(host.data_ptr() + 54 * 4, W * 4, dev.data_ptr() + 54 * 4, W * 4, 83 * 4, N, D2H,
                             s1.cuda_stream))


for name, fn, useful in (("contiguous 137 cols, 1 engine", full, N * W * 4),
                         ("contiguous 137 cols, 2 engines", full_two_engines, N * W * 4),
                         ("pitched 18 + 83 cols, 1 stream", pitched_one, N * 101 * 4),
                         ("pitched 18 + 83 cols, 2 streams", pitched_two, N * 101 * 4),
                         ("pitched 83 cols only", pitched_wide_only, N * 83 * 4),
                         ("contiguous 101 cols (bound)", compact_contiguous, N * 101 * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        fn()
        s1.synchronize()
        s2.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:34s} {dt * 1e3:8.3f} ms  {useful / dt / 1e9:7.1f} GB/s useful", flush=True)
# correctness of the pitched variant
host.zero_()
pitched_two()
torch.cuda.synchronize()
ref = dev.cpu()
assert torch.equal(host[:, :18], ref[:, :18]) and torch.equal(host[:, 54:], ref[:, 54:]) and float(host[:, 18:54].abs().sum()) == 0
print("pitched copy content ok")
