"""CPU-only checks of the drop-in boundary: the library loads, exports every entry point that
include/pioneer_b200.h declares, the ctypes structs agree with the C structs, and without a CUDA
device the product path fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from pioneer_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pioneer_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnr_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_name_the_same_entry_points():
    assert declared_functions() == sorted(_cabi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/pioneer_b200.h but not exported"
    assert lib.pnr_abi_version() == _cabi.PNR_ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    src = tmp_path / "layout.c"
    fields_m = ["dof", "axis", "tip_xyz", "lower", "effort", "body_mass", "body_inertia", "capsule_body", "capsule_p1"]
    fields_c = ["max_v_to_r", "target_lo", "timestep", "frame_skip", "gravity", "max_episode_steps", "arith",
                "obs_mode", "auto_reset", "mode", "kp", "n_obstacles", "obstacle_type", "obstacle_p", "contact_penalty"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
             'printf("%zu %zu\\n", sizeof(pnr_model), sizeof(pnr_config));']
    lines += [f'printf("%zu\\n", offsetof(pnr_model, {f}));' for f in fields_m]
    lines += [f'printf("%zu\\n", offsetof(pnr_config, {f}));' for f in fields_c]
    lines += ["return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    nums = [int(x) for x in out]
    assert nums[0] == C.sizeof(_cabi.pnr_model) and nums[1] == C.sizeof(_cabi.pnr_config)
    want = [getattr(_cabi.pnr_model, f).offset for f in fields_m] + [getattr(_cabi.pnr_config, f).offset for f in fields_c]
    assert nums[2:] == want


def test_default_config_is_the_reference_defaults():
    lib = _cabi.load()
    c = _cabi.pnr_config()
    lib.pnr_default_config(C.byref(c))
    assert (c.max_v_to_r, c.max_a_to_v, c.done_distance) == (2.0, 10.0, 0.1)          # pioneer_knm_env.py:21-24
    assert (c.award_max, c.award_done, c.award_potential_slope, c.penalty_step) == (100.0, 5.0, 10.0, 0.01)
    assert list(c.target_lo) == [15, -10, 2] and list(c.target_hi) == [25, 10, 6]
    assert c.timestep == 1 / 240 and c.frame_skip == 10 and c.gravity == 0           # bullet_env.py:38-41
    assert c.max_episode_steps == 500                                                 # pioneer_knm_train.py:27


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pioneer_b200 import BatchedPioneerEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedPioneerEnv(4)
    # and the C-ABI itself refuses, too
    from pioneer_b200.urdf import flatten_urdf
    lib = _cabi.load()
    model = _cabi.model_from_chain(flatten_urdf())
    cfg = _cabi.pnr_config()
    lib.pnr_default_config(C.byref(cfg))
    h = C.c_void_p()
    rc = lib.pnr_create(C.byref(model), C.byref(cfg), 4, 0, 0, 0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.pnr_last_error()


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "pioneer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
    code = "import sys, pioneer_b200, pioneer_b200.batched_env, pioneer_b200.vector_env, pioneer_b200.launch; " \
           "assert not [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_build_freshness_is_decided_by_content_not_by_file_times(tmp_path, monkeypatch):
    """A snapshot copy to another box does not keep file times: the library is stale iff the hash of its sources
    changed (pioneer_b200/build.py), so N ranks importing the package on a fresh box do not all start nvcc."""
    from pioneer_b200 import build
    build.build()
    assert not build._stale() and build._built_hash(build.LIB_PATH) == build._source_hash()    # the stamp is inside the .so
    src = os.path.join(build.CSRC, "pnr_launch.h")
    os.utime(src, None)                                  # newer file time, same content
    assert not build._stale()
    h = build._source_hash()
    monkeypatch.setattr(build, "NVCC_FLAGS", build.NVCC_FLAGS + ["-DX"])
    assert build._source_hash() != h and build._stale()
