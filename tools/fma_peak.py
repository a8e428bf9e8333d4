"""FP32 FMA peak of this GPU (TFLOP/s, 2 flops per FMA): python tools/fma_peak.py"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libfma_peak.so")


def build():
    src = os.path.join(HERE, "csrc", "fma_peak.cu")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-Xcompiler", "-fPIC",
                        "-shared", src, "-o", SO], check=True)
    return SO


def measure(iters=20000):
    lib = ctypes.CDLL(build())
    lib.fma_peak_tflops.restype = ctypes.c_double
    lib.fma_peak_tflops.argtypes = [ctypes.c_int]
    return lib.fma_peak_tflops(iters)


if __name__ == "__main__":
    print(f"FP32 FMA peak: {measure():.1f} TFLOP/s")
