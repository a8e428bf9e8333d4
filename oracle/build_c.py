"""Build the C restatement of the oracle (test infrastructure): gcc -> oracle/_build/libreach_oracle.so.

    python -m oracle.build_c [--force]

-ffp-contract=off: the integrator's float32 operations must round one at a time, like NumPy's (no FMA)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "reach_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(OUT_DIR, "libreach_oracle.so")
FLAGS = ["-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Werror"]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC):
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [os.environ.get("CC", "gcc")] + FLAGS + [SRC, "-o", LIB_PATH, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
