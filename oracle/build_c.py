"""Build the C restatements of the oracle (test infrastructure): gcc reach_oracle.c dynamics_oracle.c -> oracle/_build/libreach_oracle.so.

    python -m oracle.build_c [--force]

-ffp-contract=off: the integrator's float32 operations must round one at a time, like NumPy's (no FMA)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(HERE, f) for f in ("reach_oracle.c", "dynamics_oracle.c")]
HEADERS = [os.path.join(HERE, "contact.h")]
OUT_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(OUT_DIR, "libreach_oracle.so")
# -pthread: the dynamic-mode oracle steps its envs on plain pthreads (65,536 envs x 10 float64 ABA substeps per step)
FLAGS = ["-O3", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-Wall", "-Werror"]


def build(force: bool = False) -> str:
    """Built into a temporary file and renamed into place, so concurrent callers (pool workers of the CPU baseline)
    never open a half-written library."""
    newest = max(os.path.getmtime(f) for f in SOURCES + HEADERS + [os.path.abspath(__file__)])
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
    cmd = [os.environ.get("CC", "gcc")] + FLAGS + SOURCES + ["-o", tmp, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
