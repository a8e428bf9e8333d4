"""The closed-form segment / box distance of the CUDA path (pnr_segment_box_exact + pnr_segment_box_inside in
pioneer_b200/csrc/pnr_kernels.cuh), restated operation by operation in float32 numpy and checked against the float64
oracle (oracle/reach_oracle.py::segment_box_distance, which sorts the breakpoints and minimises every piece).  CPU only:
this pins the ALGORITHM; tests/test_gpu_obstacles.py pins the kernels."""
import numpy as np

from oracle.reach_oracle import segment_box_distance

F = np.float32


def _gap_slope(t, a, d, e):
    """g(t) = F'(t) / 2 with F = sum_i max(|x_i| - e_i, 0)^2 (pnr_box_gap_slope)."""
    x = (a + t * d).astype(F)
    q = np.maximum(np.abs(x) - e, F(0)).astype(F)
    return F(np.sum((np.copysign(q, x) * d).astype(F), dtype=F))


def _face_term(t, a, d, e):
    t = F(0) if t != t else F(min(max(t, F(0)), F(1)))           # fmaxf drops a NaN candidate
    x = (a + t * d).astype(F)
    return F(np.max(np.abs(x) - e))


def _inside(a, d, e):
    best = min(_face_term(F(0), a, d, e), _face_term(F(1), a, d, e))
    with np.errstate(all="ignore"):
        for i in range(3):
            best = min(best, _face_term(F(-a[i] / d[i]), a, d, e))
        for i, j in ((0, 1), (0, 2), (1, 2)):
            for si in (F(1), F(-1)):
                for sj in (F(1), F(-1)):
                    num = F((sj * a[j] - e[j]) - (si * a[i] - e[i]))
                    den = F(si * d[i] - sj * d[j])
                    best = min(best, _face_term(F(num / den), a, d, e))
    return best


def device_distance(a, b, e):
    a, e = a.astype(F), e.astype(F)
    d = (b.astype(F) - a).astype(F)
    lo, hi = F(0), F(1)
    glo, ghi = _gap_slope(lo, a, d, e), _gap_slope(hi, a, d, e)
    if glo >= 0:
        t = F(0)
    elif ghi <= 0:
        t = F(1)
    else:
        with np.errstate(all="ignore"):
            for i in range(3):
                inv = F(1) / d[i]
                for s in (e[i], -e[i]):
                    tb = F((s - a[i]) * inv)
                    if lo < tb < hi:                              # inf / NaN candidates fail both tests
                        gb = _gap_slope(tb, a, d, e)
                        if gb < 0:
                            lo, glo = tb, gb
                        else:
                            hi, ghi = tb, gb
        w = F(ghi - glo)
        t = F(lo + (hi - lo) * (-glo / w)) if w > 0 else lo
        t = min(max(t, lo), hi)
    x = (a + t * d).astype(F)
    q = np.maximum(np.abs(x) - e, F(0))
    f2 = F(np.sum(q * q, dtype=F))
    if f2 > F(1e-8):
        return float(np.sqrt(f2)), "outside"
    inner = _inside(a, d, e)
    return (float(inner), "inside") if inner < 0 else (float(np.sqrt(f2)), "grazing")


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    for k in range(n):
        e = rng.uniform(0.2, 5.0, 3)
        a, b = rng.uniform(-12, 12, 3), rng.uniform(-12, 12, 3)
        if k % 7 == 0:
            b[rng.integers(3)] = a[rng.integers(3)]
        if k % 11 == 0:                                           # parallel to a coordinate axis
            b = a.copy()
            b[rng.integers(3)] += rng.uniform(-5, 5)
        if k % 13 == 0:                                           # short segments near the box: grazing and inside cases
            a = rng.uniform(-1.2, 1.2, 3) * e
            b = a + rng.uniform(-1, 1, 3)
        if k % 97 == 0:
            b = a.copy()                                          # a point
        yield a.astype(F).astype(float), b.astype(F).astype(float), e.astype(F).astype(float)


def test_closed_form_agrees_with_the_sorted_breakpoint_oracle():
    worst, kinds = 0.0, {"outside": 0, "inside": 0, "grazing": 0}
    for a, b, e in _cases(12000, 5):
        got, kind = device_distance(a, b, e)
        want = segment_box_distance(a, b, np.zeros(3), e)
        kinds[kind] += 1
        worst = max(worst, abs(got - want))
        assert abs(got - want) < 1e-5, (a, b, e, got, want, kind)
    assert kinds["outside"] > 1000 and kinds["inside"] > 1000, kinds
    assert worst < 1e-5


def test_known_answers():
    e = np.array([1.0, 2.0, 3.0])
    # through the centre along x: deepest point is the centre, depth = the smallest half extent
    assert abs(device_distance(np.array([-5.0, 0, 0]), np.array([5.0, 0, 0]), e)[0] + 1.0) < 1e-6
    # parallel to a face at distance 0.5
    assert abs(device_distance(np.array([1.5, -9.0, 0]), np.array([1.5, 9.0, 0]), e)[0] - 0.5) < 1e-6
    # nearest to an edge: from (2, 3, z) the closest box point is (1, 2, z)
    assert abs(device_distance(np.array([2.0, 3.0, -1.0]), np.array([2.0, 3.0, 1.0]), e)[0] - 2 ** 0.5) < 1e-6
    # a point at a corner's diagonal
    assert abs(device_distance(np.array([2.0, 3.0, 4.0]), np.array([2.0, 3.0, 4.0]), e)[0] - 3 ** 0.5) < 1e-6
    # end point is the minimiser
    assert abs(device_distance(np.array([4.0, 0, 0]), np.array([9.0, 1.0, 0]), e)[0] - 3.0) < 1e-6
