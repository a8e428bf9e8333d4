/*
 * CPU ORACLE (C) — test infrastructure, NOT product code.  Shared by reach_oracle.c and dynamics_oracle.c.
 *
 * Obstacle variant (BASELINE.json configs[3]).  PARITY UNPINNED: the reference robot has no collision geometry and its
 * reward has no contact term; a box (half extents (0.5, 0.5, 5) at (10, 5, 0)) and a ground plane appear only in its GUI
 * demo (pioneer/envs/pioneer/pioneer_knm_env.py:249-261), a per-episode random box only in the legacy MuJoCo path
 * (pioneer/temp/pioneer_env.py:169-192: size ~ U(size space), centre = (pos ~ U(pos space), size_z)).  This file DEFINES
 *     depth(capsule, obstacle) = max(0, radius - min over the capsule's axis segment of sdf_obstacle(x))
 * and evaluates it EXACTLY in float64:
 *     plane   sdf is linear along the segment: the nearer end point decides
 *     sphere  closest point of the segment to the centre
 *     box     sdf_box(x) = |max(q, 0)| + min(max(q_x, q_y, q_z), 0), q = |x - c| - e, is convex, so along the segment it
 *             is a convex piecewise function: quadratic (under the root) between the parameters where a coordinate
 *             crosses a face plane, linear inside the box between the parameters where the deepest face changes.
 *             All breakpoints are enumerated and sorted, every piece is minimised in closed form.
 * The CUDA path finds the same minimum from the same breakpoints, unsorted: a bracket on the piecewise-linear derivative
 * outside the box, a candidate list inside (pnr_segment_box_exact / pnr_segment_box_inside in pnr_kernels.cuh).
 */
#ifndef ORC_CONTACT_H_
#define ORC_CONTACT_H_
#include <math.h>
#include <stdint.h>

#define ORC_MAX_CAPSULES 8
#define ORC_MAX_OBSTACLES 4
#define ORC_OBST_PLANE 1
#define ORC_OBST_BOX 2
#define ORC_OBST_SPHERE 3

typedef struct {
    int32_t n_capsules;
    int32_t capsule_body[ORC_MAX_CAPSULES];
    double capsule_radius[ORC_MAX_CAPSULES];
    double capsule_p0[ORC_MAX_CAPSULES][3], capsule_p1[ORC_MAX_CAPSULES][3];
    int32_t n_obstacles;
    int32_t obstacle_type[ORC_MAX_OBSTACLES];
    double obstacle_p[ORC_MAX_OBSTACLES][3], obstacle_e[ORC_MAX_OBSTACLES][3];
    double contact_penalty;
    /* per-env random box (pioneer/temp/pioneer_env.py:169-192): obstacle `random_box` (-1 = none) is redrawn at every
     * reset: half extents ~ U(size_lo, size_hi), centre = (U(pos_lo, pos_hi), half height) -- the box stands on z = 0 */
    int32_t random_box;
    double box_pos_lo[2], box_pos_hi[2], box_size_lo[3], box_size_hi[3];
} orc_contact;

static double orc_box_sdf(const double x[3], const double e[3]) {
    double out2 = 0.0, inside = -INFINITY;
    for (int i = 0; i < 3; ++i) {
        const double q = fabs(x[i]) - e[i];
        if (q > 0.0) out2 += q * q;
        if (q > inside) inside = q;
    }
    return sqrt(out2) + (inside < 0.0 ? inside : 0.0);
}

static double orc_sdf_at(const double a[3], const double d[3], const double e[3], double t) {
    const double x[3] = {a[0] + t * d[0], a[1] + t * d[1], a[2] + t * d[2]};
    return orc_box_sdf(x, e);
}

/* min over t in [0, 1] of sdf_box(a + t d); a is relative to the box centre */
static double orc_segment_box(const double a[3], const double d[3], const double e[3]) {
    double ts[32];
    int n = 0;
    ts[n++] = 0.0; ts[n++] = 1.0;
    for (int i = 0; i < 3; ++i) {
        if (d[i] == 0.0) continue;
        const double cand[3] = {(e[i] - a[i]) / d[i], (-e[i] - a[i]) / d[i], -a[i] / d[i]};
        for (int k = 0; k < 3; ++k) if (cand[k] > 0.0 && cand[k] < 1.0) ts[n++] = cand[k];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            for (int si = -1; si <= 1; si += 2)
                for (int sj = -1; sj <= 1; sj += 2) {          /* si x_i - e_i == sj x_j - e_j */
                    const double den = si * d[i] - sj * d[j];
                    if (den == 0.0) continue;
                    const double t = (e[i] - e[j] - si * a[i] + sj * a[j]) / den;
                    if (t > 0.0 && t < 1.0) ts[n++] = t;
                }
    for (int i = 1; i < n; ++i) {                               /* insertion sort, n <= 23 */
        const double v = ts[i];
        int k = i - 1;
        while (k >= 0 && ts[k] > v) { ts[k + 1] = ts[k]; --k; }
        ts[k + 1] = v;
    }
    double best = INFINITY;
    for (int k = 0; k < n; ++k) {
        const double s = orc_sdf_at(a, d, e, ts[k]);
        if (s < best) best = s;
    }
    for (int k = 0; k + 1 < n; ++k) {
        const double t0 = ts[k], t1 = ts[k + 1];
        if (!(t1 > t0)) continue;
        const double tm = 0.5 * (t0 + t1);
        double uw = 0.0, ww = 0.0;                              /* outside: f^2 = sum_active (u_i + t w_i)^2 */
        for (int i = 0; i < 3; ++i) {
            const double x = a[i] + tm * d[i];
            if (fabs(x) - e[i] > 0.0) {
                const double s = x > 0.0 ? 1.0 : -1.0, u = s * a[i] - e[i], w = s * d[i];
                uw += u * w; ww += w * w;
            }
        }
        if (ww > 0.0) {
            double tv = -uw / ww;
            tv = tv < t0 ? t0 : (tv > t1 ? t1 : tv);
            const double s = orc_sdf_at(a, d, e, tv);
            if (s < best) best = s;
        }                                                       /* inside the box the piece is linear: its ends decide */
    }
    return best;
}

/* distance of the segment [a, b] (world) to one obstacle; kind-specific, exact */
static double orc_segment_obstacle(int kind, const double p[3], const double e[3], const double a[3], const double b[3]) {
    if (kind == ORC_OBST_PLANE) {
        double da = 0.0, db = 0.0;
        for (int i = 0; i < 3; ++i) { da += (a[i] - p[i]) * e[i]; db += (b[i] - p[i]) * e[i]; }
        return da < db ? da : db;
    }
    if (kind == ORC_OBST_SPHERE) {
        double ab2 = 0.0, pa = 0.0;
        for (int i = 0; i < 3; ++i) { ab2 += (b[i] - a[i]) * (b[i] - a[i]); pa += (p[i] - a[i]) * (b[i] - a[i]); }
        double t = pa / (ab2 > 1e-30 ? ab2 : 1e-30);
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        double c2 = 0.0;
        for (int i = 0; i < 3; ++i) { const double c = a[i] + t * (b[i] - a[i]) - p[i]; c2 += c * c; }
        return sqrt(c2) - e[0];
    }
    const double ra[3] = {a[0] - p[0], a[1] - p[1], a[2] - p[2]};
    const double d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    return orc_segment_box(ra, d, e);
}

/* world position of a point fixed in the moving frame of `body` (URDF: child = parent * T(origin) * Rot(axis, q)) */
static void orc_fk_point(const double axis[][3], const double origin_xyz[][3], const double origin_rot[][9],
                         const double* q, int body, const double point[3], double out[3]) {
    double x = point[0], y = point[1], z = point[2];
    for (int j = body; j >= 0; --j) {
        const double c = cos(q[j]), s = sin(q[j]);
        const double kx = axis[j][0], ky = axis[j][1], kz = axis[j][2];
        const double kp = (kx * x + ky * y + kz * z) * (1.0 - c);
        const double nx = x * c + (ky * z - kz * y) * s + kx * kp;
        const double ny = y * c + (kz * x - kx * z) * s + ky * kp;
        const double nz = z * c + (kx * y - ky * x) * s + kz * kp;
        const double* R = origin_rot[j];
        x = origin_xyz[j][0] + R[0] * nx + R[1] * ny + R[2] * nz;
        y = origin_xyz[j][1] + R[3] * nx + R[4] * ny + R[5] * nz;
        z = origin_xyz[j][2] + R[6] * nx + R[7] * ny + R[8] * nz;
    }
    out[0] = x; out[1] = y; out[2] = z;
}

/* sum over (capsule, obstacle) pairs of the penetration depth; `box_p` / `box_e` replace obstacle c->random_box */
static double orc_contact_depth(const orc_contact* c, const double axis[][3], const double origin_xyz[][3],
                                const double origin_rot[][9], const double* q, const double* box_p, const double* box_e) {
    double total = 0.0;
    for (int k = 0; k < c->n_capsules; ++k) {
        double a[3], b[3];
        orc_fk_point(axis, origin_xyz, origin_rot, q, c->capsule_body[k], c->capsule_p0[k], a);
        orc_fk_point(axis, origin_xyz, origin_rot, q, c->capsule_body[k], c->capsule_p1[k], b);
        for (int o = 0; o < c->n_obstacles; ++o) {
            const double* p = (o == c->random_box && box_p) ? box_p : c->obstacle_p[o];
            const double* e = (o == c->random_box && box_e) ? box_e : c->obstacle_e[o];
            const double dist = orc_segment_obstacle(c->obstacle_type[o], p, e, a, b);
            if (c->capsule_radius[k] - dist > 0.0) total += c->capsule_radius[k] - dist;
        }
    }
    return total;
}
#endif
