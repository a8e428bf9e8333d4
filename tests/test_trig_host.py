"""Accuracy of the straight-line sine/cosine of the step kernel (pioneer_b200/csrc/pnr_trig.cuh), checked on the
CPU: the header is host+device, tests/csrc/trig_check.cu compiles it with nvcc for the host and sweeps each
argument range against float64 libm.  Bound: 1.5e-7 absolute (observed 9.3e-8), which is what lets the GPU parity
tests hold cos/sin columns to 5e-7 against numpy's float32 cos/sin."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_fast_sincos_accuracy(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "trig_check"
    subprocess.run([nvcc, "-Wno-deprecated-gpu-targets", "-O2", "-o", str(exe),
                    os.path.join(ROOT, "tests", "csrc", "trig_check.cu")], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    rows = {ln.split()[0]: ln.split()[1:] for ln in out}
    for name in ("bounded_pi", "bounded_2pi", "bounded_4pi", "bounded_64", "fast_126", "fast_1e4", "fast_limit"):
        n, es, ec = rows[name]
        assert int(n) >= 4000000
        assert float(es) < 1.5e-7 and float(ec) < 1.5e-7, (name, es, ec)
    # the reset state (v = a = 0) must give exactly sin 0 = 0, cos 0 = 1
    assert rows["zero"][1:] == ["0", "1"] and rows["zero_fast"][1:] == ["0", "1"]
