"""FP32 throughput by instruction form (uniform-operand FFMA, three-register FFMA, packed FFMA2, FMUL, FADD):
    python tools/fp32_forms.py"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libfp32_forms.so")


def build():
    src = os.path.join(HERE, "csrc", "fp32_forms.cu")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-Xcompiler", "-fPIC",
                        "-shared", src, "-o", SO], check=True)
    return SO


if __name__ == "__main__":
    lib = ctypes.CDLL(build())
    lib.fp32_forms.restype = ctypes.c_double
    out = (ctypes.c_double * 5)()
    ops = lib.fp32_forms(ctypes.c_int(20000), out)
    names = ["FFMA uniform operands", "FFMA three registers", "FFMA2 packed (register pairs)", "FMUL two registers", "FADD two registers"]
    flop = [2, 2, 2, 1, 1]
    instr_per_op = [1, 1, 0.5, 1, 1]
    for n, s, f, ipo in zip(names, out, flop, instr_per_op):
        warp_instr = ops * ipo / 32
        print(f"{n:32s} {s * 1e3:8.3f} ms  {ops * f / s / 1e12:7.2f} TFLOP/s  {warp_instr / s / 1e9:8.1f} G warp-instr/s "
              f"= {warp_instr / s / (148 * 4) / 1.965e9:5.3f} per clock per SMSP at 1965 MHz")
