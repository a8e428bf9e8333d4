"""The base-to-tip frame walk the CUDA contact code uses for the capsule end points (pnr_frame_advance in
pioneer_b200/csrc/pnr_kernels.cuh: o += R origin_j; R <- R R_origin_j Rot(axis_j, q_j), column mixing for coordinate axes,
Rodrigues' matrix otherwise), restated in numpy and checked against the oracle's tip-to-base point FK
(oracle/reach_oracle.py::fk_point).  CPU only: pins the recursion; tests/test_gpu_obstacles.py pins the kernels."""
import numpy as np

from oracle.reach_oracle import OracleChain, fk_point
from pioneer_b200.urdf import flatten_urdf


def _advance(chain, j, q, R, o):
    o = o + R @ chain.origin_xyz[j]
    R = R @ chain.origin_rot[j]
    k = chain.axis[j]
    c, s = np.cos(q), np.sin(q)
    code = [tuple(k) == t for t in ((1, 0, 0), (0, 1, 0), (0, 0, 1))]
    if any(code):                                   # mix two columns: u' = c u + s w, w' = -s u + c w
        u, w = ((1, 2), (2, 0), (0, 1))[code.index(True)]
        cu, cw = R[:, u].copy(), R[:, w].copy()
        R = R.copy()
        R[:, u] = c * cu + s * cw
        R[:, w] = c * cw - s * cu
    else:                                           # Rot = c I + s [k]x + (1 - c) k k^T
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = R @ (c * np.eye(3) + s * K + (1 - c) * np.outer(k, k))
    return R, o


def _walk(chain, q):
    """End points of every capsule, visiting the capsules in body order while the chain is walked once."""
    R, o, joint, out = np.eye(3), np.zeros(3), 0, {}
    order = sorted(range(len(chain.capsules)), key=lambda i: chain.capsules[i][0])
    for i in order:
        body, _, p0, p1 = chain.capsules[i]
        while joint <= body:
            R, o = _advance(chain, joint, q[joint], R, o)
            joint += 1
        out[i] = (o + R @ p0, o + R @ p1)
    return out


def test_frame_walk_matches_point_fk_on_the_shipped_robot():
    chain = OracleChain.from_model(flatten_urdf())
    assert len(chain.capsules) == 5
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.uniform(chain.lower, chain.upper)
        got = _walk(chain, q)
        for i, (body, _, p0, p1) in enumerate(chain.capsules):
            np.testing.assert_allclose(got[i][0], fk_point(chain, q, body, p0), atol=1e-11)
            np.testing.assert_allclose(got[i][1], fk_point(chain, q, body, p1), atol=1e-11)


def test_frame_walk_with_general_axes_origin_rotations_and_unordered_capsules():
    rng = np.random.default_rng(1)
    base = OracleChain.from_model(flatten_urdf())
    axis = rng.normal(size=(6, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    axis[2] = (0, 1, 0)                             # one coordinate axis among general ones
    rots = []
    for _ in range(6):                              # random proper rotations for the joint origins
        m, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        rots.append(m * np.sign(np.linalg.det(m)))
    caps = tuple((int(b), 0.5, rng.normal(size=3), rng.normal(size=3)) for b in (4, 0, 5, 2, 2, 1))
    chain = OracleChain(axis, rng.normal(size=(6, 3)), np.array(rots), base.tip_xyz, base.lower, base.upper, caps)
    for _ in range(50):
        q = rng.uniform(-3, 3, 6)
        got = _walk(chain, q)
        for i, (body, _, p0, p1) in enumerate(caps):
            np.testing.assert_allclose(got[i][0], fk_point(chain, q, body, p0), atol=1e-10)
            np.testing.assert_allclose(got[i][1], fk_point(chain, q, body, p1), atol=1e-10)
