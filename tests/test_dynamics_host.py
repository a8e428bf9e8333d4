"""The float32 articulated-body algorithm of the dynamic-mode kernel (pioneer_b200/csrc/pnr_dynamics.cuh), compiled for
the HOST with nvcc (tests/csrc/aba_check.cu) and compared with the float64 oracle (oracle/dynamics_oracle.py) on the
CPU.  Four variants must agree: the run-time generic chain, the first specialisation for the shipped robot, the
sparsity-aware specialisation (pnr_aba_pioneer<false>) and the one the kernel runs for the shipped URDF's stub
inertials (pnr_aba_pioneer<true>).  The parameter block comes from the product
library's own host-side builder (pnr_debug_build_params), so the constants under test are the ones the GPU gets.
No CUDA device is needed and no kernel runs here; the GPU tests (tests/test_gpu_dynamic.py) check the kernel itself.
PARITY UNPINNED vs PyBullet (see the oracle's header)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle.dynamics_oracle import DynChain, DynConfig, aba, dynamic_substeps
from pioneer_b200 import _cabi
from pioneer_b200.urdf import flatten_urdf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="needs nvcc")

F = C.POINTER(C.c_float)


def fp(a):
    return a.ctypes.data_as(F)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("aba") / "aba_check.so"
    subprocess.run([NVCC, "-Wno-deprecated-gpu-targets", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                    "-o", str(out), os.path.join(ROOT, "tests", "csrc", "aba_check.cu")], check=True)
    lib = C.CDLL(str(out))
    lib.aba_params_size.restype = C.c_long
    lib.aba_check.argtypes = [C.c_void_p, C.c_int, C.c_long, F, F, F, F]
    lib.aba_substeps.argtypes = [C.c_void_p, C.c_int, C.c_long, F, F, F]
    return lib


def params_blob(harness, gravity=9.81, kp=0.0, kd=0.0, torque_scale=1.0, chain=None, bullet=None):
    lib = _cabi.load()
    lib.pnr_debug_build_params.restype = C.c_int64
    lib.pnr_debug_build_params.argtypes = [C.POINTER(_cabi.pnr_model), C.POINTER(_cabi.pnr_config), C.c_int64,
                                           C.c_void_p, C.c_int64]
    chain = chain or flatten_urdf()
    model = _cabi.model_from_chain(chain)
    cfg = _cabi.pnr_config()
    lib.pnr_default_config(C.byref(cfg))
    cfg.mode, cfg.gravity, cfg.kp, cfg.kd, cfg.torque_scale = _cabi.PNR_MODE_DYNAMIC, gravity, kp, kd, torque_scale
    if bullet is not None:
        cfg.stepping = _cabi.PNR_STEPPING_BULLET
        cfg.link_damping, cfg.max_velocity = bullet["link_damping"], bullet["max_velocity"]
        cfg.motor_kp, cfg.motor_kd, cfg.motor_max_force = bullet["motor_kp"], bullet["motor_kd"], bullet["motor_max_force"]
    size = lib.pnr_debug_build_params(C.byref(model), C.byref(cfg), 1, None, 0)
    assert size == harness.aba_params_size(), "harness and library disagree on sizeof(PnrParams)"
    blob = C.create_string_buffer(size)
    assert lib.pnr_debug_build_params(C.byref(model), C.byref(cfg), 1, blob, size) == size
    return chain, blob


def states(chain, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(chain.lower), np.asarray(chain.upper)
    q = rng.uniform(lo, hi, size=(n, 6)).astype(np.float32)
    qd = (rng.normal(size=(n, 6)) * 1.5).astype(np.float32)
    return rng, q, qd


@pytest.mark.parametrize("gravity", [0.0, 9.81])
def test_three_float32_variants_against_the_float64_oracle(harness, gravity):
    chain, blob = params_blob(harness, gravity=gravity)
    dyn = DynChain.from_model(chain)
    n = 400
    rng, q, qd = states(chain, n, seed=2)
    tau = (rng.normal(size=(n, 6)) * 300.0).astype(np.float32)
    ref = np.stack([aba(dyn, q[e].astype(np.float64), qd[e].astype(np.float64), tau[e].astype(np.float64), gravity)
                    for e in range(n)])
    scale = np.abs(ref).max(axis=0) + 1e-9
    outs = []
    for variant in (0, 1, 2, 3):
        out = np.empty((n, 6), np.float32)
        assert harness.aba_check(blob, variant, n, fp(q), fp(qd), fp(tau), fp(out)) == 0
        err = np.abs(out - ref) / scale
        assert err.max() < 1e-5, (variant, err.max())     # float32 against float64, relative to each joint's range (observed 9e-7)
        outs.append(out)
    # the sparsity-aware specialisation is at least as close to the oracle as the version it replaces
    e1 = (np.abs(outs[1] - ref) / scale).mean()
    for k in (2, 3):
        e2 = (np.abs(outs[k] - ref) / scale).mean()
        assert e2 <= 1.5 * e1 + 1e-7, (k, e1, e2)


def test_zero_input_gives_exact_zero_acceleration(harness):
    """zero velocity, gravity and torque -> every term is an exact zero."""
    chain, blob = params_blob(harness, gravity=0.0)
    n = 64
    _, q, _ = states(chain, n, seed=4)
    zero = np.zeros((n, 6), np.float32)
    for variant in (0, 1, 2, 3):
        out = np.full((n, 6), np.nan, np.float32)
        assert harness.aba_check(blob, variant, n, fp(q), fp(zero), fp(zero), fp(out)) == 0
        assert (out == 0).all(), variant


@pytest.mark.parametrize("kp,kd,scale", [(0.0, 0.0, 50.0), (800.0, 200.0, 1e4)])
def test_one_env_step_of_substeps_against_the_oracle(harness, kp, kd, scale):
    """frame_skip (10) substeps of control + ABA + semi-implicit Euler + limit stops; bars as on the GPU:
    |dq| <= 2e-5 rad, |dqd| <= 2e-4 rad/s."""
    gravity = 9.81
    chain, blob = params_blob(harness, gravity=gravity, kp=kp, kd=kd, torque_scale=scale)
    dyn = DynChain.from_model(chain)
    cfg = DynConfig(gravity=gravity, kp=kp, kd=kd, torque_scale=scale)
    n = 200
    rng, q, qd = states(chain, n, seed=7)
    q = (q * 0.8).astype(np.float32)
    qd = (qd * 0.3).astype(np.float32)
    lo32, hi32 = np.asarray(chain.lower, np.float32), np.asarray(chain.upper, np.float32)
    if kp or kd:
        act = rng.uniform(lo32, hi32, size=(n, 6)).astype(np.float32)
    else:
        act = (rng.normal(size=(n, 6)) * scale).astype(np.float32)
    ref_q, ref_qd = np.empty((n, 6)), np.empty((n, 6))
    for e in range(n):
        ref_q[e], ref_qd[e] = dynamic_substeps(dyn, cfg, q[e].astype(np.float64), qd[e].astype(np.float64),
                                               act[e].astype(np.float64), lo32.astype(np.float64), hi32.astype(np.float64))
    for pioneer_chain in (0, 1, 2):
        q1, qd1 = q.copy(), qd.copy()
        assert harness.aba_substeps(blob, pioneer_chain, n, fp(q1), fp(qd1), fp(act)) == 0
        assert np.abs(q1 - ref_q).max() <= 2e-5, (pioneer_chain, np.abs(q1 - ref_q).max())
        assert np.abs(qd1 - ref_qd).max() <= 2e-4, (pioneer_chain, np.abs(qd1 - ref_qd).max())


@pytest.mark.parametrize("case", ["motors", "weak motors", "no motors", "velocity clamp"])
def test_bullet_like_substeps_against_the_compiled_oracle(harness, case):
    """PNR_STEPPING_BULLET (opt-in): per-link damping, POSITION_CONTROL as a velocity-level motor constraint with impulse
    clamp (projected Gauss-Seidel on M^-1), +-max_velocity clamp -- the float32 header against the float64 restatement in
    oracle/dynamics_oracle.c (stepping = 1; both written from memory of btMultiBody, UNPINNED vs PyBullet).
    One env step = 10 substeps; bars: |dq| <= 5e-5, |dqd| <= 2e-3 (the motor solve divides by dt: qd carries 240 x the
    float32 rounding of q)."""
    from oracle.c_dyn_oracle import CDynOracleBatch, DynEnvConfig
    bullet = dict(link_damping=0.04, max_velocity=100.0, motor_kp=0.1, motor_kd=1.0,
                  motor_max_force={"motors": 5e4, "weak motors": 300.0, "no motors": 0.0, "velocity clamp": 5e4}[case])
    if case == "velocity clamp":
        bullet.update(max_velocity=2.0, motor_kp=0.5)
    gravity = 9.81
    chain, blob = params_blob(harness, gravity=gravity, bullet=bullet)
    orc = CDynOracleBatch(chain, 1, DynEnvConfig(gravity=gravity, stepping="bullet", **bullet))
    n = 120
    rng, q, qd = states(chain, n, seed=17)
    q, qd = (q * 0.8).astype(np.float32), (qd * 0.3).astype(np.float32)
    lo32, hi32 = np.asarray(chain.lower, np.float32), np.asarray(chain.upper, np.float32)
    act = rng.uniform(lo32, hi32, size=(n, 6)).astype(np.float32)
    ref_q, ref_qd = np.empty((n, 6)), np.empty((n, 6))
    for e in range(n):
        ref_q[e], ref_qd[e] = orc.substeps(q[e], qd[e], act[e], 10)
    moved = np.abs(ref_q - q).max()
    assert moved > 0.05                                              # the step really does something
    if case == "velocity clamp":
        assert np.abs(ref_qd).max() <= 2.0 + 1e-12 and (np.abs(ref_qd) > 1.99).any()
    for pioneer_chain in (0, 1):
        q1, qd1 = q.copy(), qd.copy()
        assert harness.aba_substeps(blob, pioneer_chain, n, fp(q1), fp(qd1), fp(act)) == 0
        dq, dqd = np.abs(q1 - ref_q).max(), np.abs(qd1 - ref_qd).max()
        assert dq <= 5e-5 and dqd <= 2e-3, (case, pioneer_chain, dq, dqd)


def test_general_inertials_take_the_non_isotropic_specialisation(harness):
    """The shipped URDF's links have stub inertials (centre of mass on the frame origin, isotropic), which hides every
    term that multiplies a centre-of-mass offset or an off-diagonal inertia.  Same kinematic structure with random
    full inertias and centre-of-mass offsets on ALL six composite bodies: the generic chain, the first specialisation and
    pnr_aba_pioneer<false> must still agree with the float64 oracle; the isotropic variant must be refused."""
    import copy
    rng = np.random.default_rng(11)
    chain = copy.deepcopy(flatten_urdf())
    com, inertia = [], []
    for j in range(6):
        a = rng.normal(size=(3, 3))
        inertia.append((a @ a.T + 2.0 * np.eye(3)) * float(chain.body_mass[j]))       # symmetric positive definite
        com.append(rng.uniform(-1.5, 1.5, size=3))
    chain.body_com = np.array(com)
    chain.body_inertia = np.array(inertia)
    gravity = 9.81
    chain_out, blob = params_blob(harness, gravity=gravity, chain=chain)
    dyn = DynChain.from_model(chain_out)
    n = 300
    rng2, q, qd = states(chain_out, n, seed=12)
    tau = (rng2.normal(size=(n, 6)) * 300.0).astype(np.float32)
    ref = np.stack([aba(dyn, q[e].astype(np.float64), qd[e].astype(np.float64), tau[e].astype(np.float64), gravity)
                    for e in range(n)])
    scale = np.abs(ref).max(axis=0) + 1e-9
    for variant in (0, 1, 2):
        out = np.empty((n, 6), np.float32)
        assert harness.aba_check(blob, variant, n, fp(q), fp(qd), fp(tau), fp(out)) == 0
        err = np.abs(out - ref) / scale
        assert err.max() < 2e-5, (variant, err.max())
    out = np.empty((n, 6), np.float32)
    assert harness.aba_check(blob, 3, n, fp(q), fp(qd), fp(tau), fp(out)) == -1       # dyn_iso_links is off for this model
    # one env step of substeps through the kernel's own driver, chain kinds 0 and 1
    cfg = DynConfig(gravity=gravity, kp=0.0, kd=0.0, torque_scale=50.0)
    _, blob2 = params_blob(harness, gravity=gravity, torque_scale=50.0, chain=chain)
    lo32, hi32 = np.asarray(chain_out.lower, np.float32), np.asarray(chain_out.upper, np.float32)
    q0, qd0 = (q[:100] * 0.8).astype(np.float32), (qd[:100] * 0.3).astype(np.float32)
    act = (rng2.normal(size=(100, 6)) * 50.0).astype(np.float32)
    want = [dynamic_substeps(dyn, cfg, q0[e].astype(np.float64), qd0[e].astype(np.float64), act[e].astype(np.float64),
                             lo32.astype(np.float64), hi32.astype(np.float64)) for e in range(100)]
    for kind in (0, 1):
        q1, qd1 = q0.copy(), qd0.copy()
        assert harness.aba_substeps(blob2, kind, 100, fp(q1), fp(qd1), fp(act)) == 0
        assert max(np.abs(q1[e] - want[e][0]).max() for e in range(100)) <= 2e-5
        assert max(np.abs(qd1[e] - want[e][1]).max() for e in range(100)) <= 2e-4
