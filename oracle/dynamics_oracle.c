/*
 * CPU ORACLE (C) for the dynamic (Tier-B) mode — test infrastructure, NOT product code.
 *
 * The compiled twin of oracle/dynamics_oracle.py (float64 Featherstone articulated-body algorithm, PD / torque control,
 * semi-implicit Euler substeps, inelastic joint stops) PLUS the env around it, which the Python file does not have:
 * reward, done / TimeLimit flags, the 137-column observation, episode statistics and the Philox auto-reset, with the
 * batched call semantics of include/pioneer_b200.h.  Fast enough (pthreads over envs) to check the CUDA path at
 * BASELINE.json's full sizes: 65,536 envs x 200 steps x 10 substeps in seconds on the GPU box's host cores.
 *
 * PARITY UNPINNED vs PyBullet: the reference env never runs Bullet's dynamics with non-zero inputs (gravity 0, joints
 * teleported with zero velocity, no motor targets) and PyBullet is not installable here.  The
 * dynamics semantics are DEFINED by oracle/dynamics_oracle.py; tests/test_oracle_dyn_c.py checks
 * this file against it (and through it against CRBA + RNEA and Lagrange's equations).  The env layer around the dynamics
 * follows the reference source exactly as oracle/reach_oracle.c does:
 *   pioneer/envs/pioneer/pioneer_knm_env.py:76-105   reset_world (a = v = 0, potential = 0)
 *   pioneer/envs/pioneer/pioneer_knm_env.py:151-165  distance, potential, done, reward
 *   pioneer/envs/pioneer/pioneer_knm_env.py:184-211  observe
 *   pioneer/envs/bullet/bullet_env.py:187-197        step ordering (act, observe, counters)
 *   pioneer/envs/bullet/bullet_scene.py:123-155      Joint.control_position / control_velocity (the motor interface)
 *   pioneer/envs/bullet/bullet_scene.py:273-275      World.step: frame_skip x stepSimulation
 *   pioneer/launch/pioneer_knm_train.py:27           TimeLimit(max_episode_steps=500)
 *
 * `stepping = 1` selects the opt-in BULLET-LIKE substep [UPSTREAM-MEMORY: btMultiBody / btMultiBodyJointMotor as of
 * Bullet 2.8x; nothing in /root/reference pins it]: per-link linear / angular velocity damping, POSITION_CONTROL as a
 * velocity-level motor constraint solved per joint with the impulse clamp force * dt, joint velocities clamped to
 * +-max_coordinate_velocity.  See dyn_substep_bullet below.
 *
 * Lock-step checking: the CUDA path computes in float32, this file in float64, and the dynamics are not contracting in
 * general, so a free-running comparison drifts.  orc_dyn_step therefore (a) reports the state it reached on its own
 * (own_q / own_qd: compared with the device state inside the stated bars) and (b) can then ADOPT the device's float32
 * state for the env layer and the next step, so that reward / flags / observation / reset are compared on identical
 * inputs, exactly like the kinematic mode.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "contact.h"

#define DOF 6
#define OBS_DIM 137
#define ORC_DONE 1
#define ORC_TRUNCATED 2

typedef struct {
    /* chain, same content as pnr_model (float64) */
    double axis[DOF][3], origin_xyz[DOF][3], origin_rot[DOF][9], tip_xyz[3];
    double lower[DOF], upper[DOF], effort[DOF], damping[DOF];
    double body_mass[DOF], body_com[DOF][3], body_inertia[DOF][9];
    /* env configuration */
    double done_distance, award_max, award_done, award_potential_slope, penalty_step;
    double target_lo[3], target_hi[3];
    double timestep, gravity, kp, kd, torque_scale;
    int32_t frame_skip, max_episode_steps, auto_reset, obs_autoreset;
    /* stepping: 0 = explicit PD + semi-implicit Euler, 1 = Bullet-like (see header) */
    int32_t stepping;
    double link_damping;               /* Bullet: linear and angular damping of every link, default 0.04           */
    double max_velocity;               /* Bullet: m_maxCoordinateVelocity, default 100                              */
    double motor_kp, motor_kd;         /* setJointMotorControl2 positionGain / velocityGain (bullet_scene.py:123-142) */
    double motor_max_force;            /* `force` of the motor; impulse clamp = force * dt                          */
    orc_contact contact;
} dyn_params;

typedef struct {
    double q[DOF], qd[DOF];
    float a[DOF];
    double target[3];
    double potential;
    int32_t elapsed;
    float ep_return;
    double box_p[3], box_e[3];
} dyn_env;

typedef struct {
    dyn_params p;
    int64_t n, env_id_base;
    uint64_t seed;
    uint32_t tick;
    float r_lo[DOF], r_hi[DOF];
    dyn_env* envs;
    double stats[8];
    /* scratch of one step (per env): decided in the parallel part, consumed by the sequential part */
    uint8_t* s_done;
    int32_t n_threads;
} dyn_batch;

/* ---- Philox4x32-10, bit-identical to reach_oracle.c / reach_oracle.py / the CUDA path ---- */
static void philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
static float uniform32(float lo, float hi, float u) { const float span = hi - lo; const float m = span * u; return lo + m; }

/* 14 uniforms per reset: 6 joints, 3 target coordinates (reference draw order, pioneer_knm_env.py:80-90), then the
 * per-env box of the obstacle variant: 3 half extents, 2 centre coordinates (pioneer/temp/pioneer_env.py:173-174) */
static void reset_draws(const dyn_batch* b, int64_t global_env, uint32_t tick, float u[16]) {
    uint32_t out[16];
    for (uint32_t blk = 0; blk < 4; ++blk) {
        const uint32_t ctr[4] = {(uint32_t)global_env, (uint32_t)((uint64_t)global_env >> 32), tick, blk};
        philox4x32_10(ctr, (uint32_t)b->seed, (uint32_t)(b->seed >> 32), out + 4 * blk);
    }
    for (int i = 0; i < 16; ++i) u[i] = u01(out[i]);
}

/* ---- 3-vector / 3x3 helpers (row-major) ---- */
static void m3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void m3_mul_tn(const double* A, const double* B, double* C) {      /* A^T B */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
static void m3_mul_nt(const double* A, const double* B, double* C) {      /* A B^T */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
static void m3_vec(const double* A, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static void m3_tvec(const double* A, const double* v, double* o) {        /* A^T v */
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
static void v3_cross(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void skew(const double* v, double* S) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
/* Rot(axis, q) = 1 c + s [k]x + (1 - c) k k^T   (dynamics_oracle.py::rot_axis) */
static void rot_axis(const double* k, double q, double* R) {
    const double c = cos(q), s = sin(q);
    double K[9];
    skew(k, K);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = (i == j ? c : 0.0) + s * K[3 * i + j] + (1.0 - c) * k[i] * k[j];
}

/* A spatial (articulated) inertia [[I, H], [H^T, M]]: n = I w + H v, f = H^T w + M v; spatial vectors are [angular; linear]
 * in the body's own frame.  The Pluecker transform parent -> child is X = [[E, 0], [-E [p]x, E]] with E = R^T
 * (dynamics_oracle.py::x_motion); everything below is that 6x6 algebra written in 3x3 blocks. */
typedef struct { double I[9], H[9], M[9]; } sp_inertia;

typedef struct {
    double E[DOF][9];                 /* parent -> child rotation */
    double w[DOF][3], v[DOF][3];      /* body velocity            */
    double c_ang[DOF][3], c_lin[DOF][3];
    double U_ang[DOF][3], U_lin[DOF][3], d[DOF], u[DOF];
} aba_work;

/* body velocities from (q, qd): pass 1 of the ABA, also used by the Bullet-like link damping */
static void body_velocities(const dyn_params* p, const double* q, const double* qd, aba_work* W) {
    for (int i = 0; i < DOF; ++i) {
        double Rq[9], R[9];
        rot_axis(p->axis[i], q[i], Rq);
        m3_mul(p->origin_rot[i], Rq, R);                       /* child axes in parent coordinates */
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) W->E[i][3 * a + c] = R[3 * c + a];
        if (i == 0) {
            for (int k = 0; k < 3; ++k) { W->w[0][k] = p->axis[0][k] * qd[0]; W->v[0][k] = 0.0; }
        } else {
            double pxw[3], t[3];
            v3_cross(p->origin_xyz[i], W->w[i - 1], pxw);
            for (int k = 0; k < 3; ++k) t[k] = W->v[i - 1][k] - pxw[k];
            m3_vec(W->E[i], W->w[i - 1], W->w[i]);
            m3_vec(W->E[i], t, W->v[i]);
            for (int k = 0; k < 3; ++k) W->w[i][k] += p->axis[i][k] * qd[i];
        }
    }
}

static void body_inertia(const dyn_params* p, int i, sp_inertia* S) {      /* dynamics_oracle.py::spatial_inertia */
    const double m = p->body_mass[i];
    const double* c = p->body_com[i];
    double C[9], CCt[9];
    skew(c, C);
    m3_mul_nt(C, C, CCt);
    for (int k = 0; k < 9; ++k) {
        S->I[k] = p->body_inertia[i][k] + m * CCt[k];
        S->H[k] = m * C[k];
        S->M[k] = (k % 4 == 0) ? m : 0.0;
    }
}

/* qdd = ABA(q, qd, tau), fixed base, gravity along -z of the base frame, plus an optional external spatial force on each
 * body (f_ext_ang / f_ext_lin in body coordinates, NULL = none: the Bullet-like link damping).  dynamics_oracle.py::aba */
static void aba(const dyn_params* p, const double* q, const double* qd, const double* tau,
                const double (*f_ext_ang)[3], const double (*f_ext_lin)[3], double* qdd, aba_work* W) {
    sp_inertia IA[DOF];
    double pA_ang[DOF][3], pA_lin[DOF][3];
    body_velocities(p, q, qd, W);
    for (int i = 0; i < DOF; ++i) {
        double s[3], n[3], f[3], t1[3], t2[3];
        for (int k = 0; k < 3; ++k) s[k] = p->axis[i][k] * qd[i];
        v3_cross(W->w[i], s, W->c_ang[i]);                     /* crm(v) vJ */
        v3_cross(W->v[i], s, W->c_lin[i]);
        body_inertia(p, i, &IA[i]);
        m3_vec(IA[i].I, W->w[i], n); m3_vec(IA[i].H, W->v[i], t1);
        for (int k = 0; k < 3; ++k) n[k] += t1[k];
        m3_tvec(IA[i].H, W->w[i], f); m3_vec(IA[i].M, W->v[i], t1);
        for (int k = 0; k < 3; ++k) f[k] += t1[k];
        v3_cross(W->w[i], n, t1); v3_cross(W->v[i], f, t2);    /* crf(v) (I v) = [w x n + v x f; w x f] */
        for (int k = 0; k < 3; ++k) pA_ang[i][k] = t1[k] + t2[k];
        v3_cross(W->w[i], f, pA_lin[i]);
        if (f_ext_ang) for (int k = 0; k < 3; ++k) { pA_ang[i][k] -= f_ext_ang[i][k]; pA_lin[i][k] -= f_ext_lin[i][k]; }
    }
    for (int i = DOF - 1; i >= 0; --i) {
        const double* S = p->axis[i];
        m3_vec(IA[i].I, S, W->U_ang[i]);                       /* U = IA [S; 0] */
        m3_tvec(IA[i].H, S, W->U_lin[i]);
        W->d[i] = v3_dot(S, W->U_ang[i]);
        W->u[i] = tau[i] - v3_dot(S, pA_ang[i]);
        if (i == 0) break;
        sp_inertia Ia = IA[i];
        const double dinv = 1.0 / W->d[i];
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) {                      /* Ia = IA - U U^T / d */
                Ia.I[3 * a + c] -= W->U_ang[i][a] * W->U_ang[i][c] * dinv;
                Ia.H[3 * a + c] -= W->U_ang[i][a] * W->U_lin[i][c] * dinv;
                Ia.M[3 * a + c] -= W->U_lin[i][a] * W->U_lin[i][c] * dinv;
            }
        double pa_ang[3], pa_lin[3], t1[3], t2[3];
        const double ud = W->u[i] * dinv;
        m3_vec(Ia.I, W->c_ang[i], t1); m3_vec(Ia.H, W->c_lin[i], t2);
        for (int k = 0; k < 3; ++k) pa_ang[k] = pA_ang[i][k] + t1[k] + t2[k] + W->U_ang[i][k] * ud;
        m3_tvec(Ia.H, W->c_ang[i], t1); m3_vec(Ia.M, W->c_lin[i], t2);
        for (int k = 0; k < 3; ++k) pa_lin[k] = pA_lin[i][k] + t1[k] + t2[k] + W->U_lin[i][k] * ud;
        /* IA[i-1] += X^T Ia X: rotate the blocks to parent axes (B' = E^T B E), then move the reference point by p:
         *   M'' = M', H'' = H' + P M', I'' = I' - H' P + P H'^T - P M' P      with P = [p]x */
        const double* E = W->E[i];
        double T[9], Ir[9], Hr[9], Mr[9], P[9], PM[9], HP[9], PHt[9], PMP[9];
        m3_mul_tn(E, Ia.I, T); m3_mul(T, E, Ir);
        m3_mul_tn(E, Ia.H, T); m3_mul(T, E, Hr);
        m3_mul_tn(E, Ia.M, T); m3_mul(T, E, Mr);
        skew(p->origin_xyz[i], P);
        m3_mul(P, Mr, PM); m3_mul(Hr, P, HP); m3_mul_nt(P, Hr, PHt); m3_mul(PM, P, PMP);
        for (int k = 0; k < 9; ++k) {
            IA[i - 1].I[k] += Ir[k] - HP[k] + PHt[k] - PMP[k];
            IA[i - 1].H[k] += Hr[k] + PM[k];
            IA[i - 1].M[k] += Mr[k];
        }
        /* pA[i-1] += X^T pa: f_p = E^T f, n_p = E^T n + p x f_p */
        double fp[3], np_[3], pxf[3];
        m3_tvec(E, pa_lin, fp); m3_tvec(E, pa_ang, np_);
        v3_cross(p->origin_xyz[i], fp, pxf);
        for (int k = 0; k < 3; ++k) { pA_ang[i - 1][k] += np_[k] + pxf[k]; pA_lin[i - 1][k] += fp[k]; }
    }
    double a_ang[3] = {0, 0, 0}, a_lin[3] = {0, 0, p->gravity};  /* the base "accelerates upwards" by g */
    for (int i = 0; i < DOF; ++i) {
        double pxa[3], t[3], na[3], nl[3];
        v3_cross(p->origin_xyz[i], a_ang, pxa);
        for (int k = 0; k < 3; ++k) t[k] = a_lin[k] - pxa[k];
        m3_vec(W->E[i], a_ang, na); m3_vec(W->E[i], t, nl);
        for (int k = 0; k < 3; ++k) { na[k] += W->c_ang[i][k]; nl[k] += W->c_lin[i][k]; }
        qdd[i] = (W->u[i] - v3_dot(W->U_ang[i], na) - v3_dot(W->U_lin[i], nl)) / W->d[i];
        for (int k = 0; k < 3; ++k) { a_ang[k] = na[k] + p->axis[i][k] * qdd[i]; a_lin[k] = nl[k]; }
    }
}

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* inelastic joint stops at the float32 limits the env uses (dynamics_oracle.py::dynamic_substeps) */
static unsigned joint_stops(const dyn_batch* b, double* q, double* qd) {
    unsigned touched = 0;                                      /* bit i: joint i ran into a stop in this substep */
    for (int i = 0; i < DOF; ++i) {
        const double lo = (double)b->r_lo[i], hi = (double)b->r_hi[i];
        if (q[i] > hi || q[i] < lo) touched |= 1u << i;
        if (q[i] > hi && qd[i] > 0.0) qd[i] = 0.0;
        if (q[i] < lo && qd[i] < 0.0) qd[i] = 0.0;
        q[i] = clampd(q[i], lo, hi);
    }
    return touched;
}

/* stepping 0: tau = clamp(kp (u - q) - kd qd | u, +-effort * torque_scale) - damping qd; qd += qdd dt; q += qd dt */
static unsigned dyn_substep_explicit(const dyn_batch* b, double* q, double* qd, const float* action) {
    const dyn_params* p = &b->p;
    const int pd = (p->kp != 0.0 || p->kd != 0.0);
    double tau[DOF], qdd[DOF];
    aba_work W;
    for (int i = 0; i < DOF; ++i) {
        const double lim = p->effort[i] * p->torque_scale;
        double t = pd ? p->kp * ((double)action[i] - q[i]) - p->kd * qd[i] : (double)action[i];
        t = clampd(t, -lim, lim);
        tau[i] = t - p->damping[i] * qd[i];
    }
    aba(p, q, qd, tau, NULL, NULL, qdd, &W);
    for (int i = 0; i < DOF; ++i) { qd[i] += qdd[i] * p->timestep; q[i] += qd[i] * p->timestep; }
    return joint_stops(b, q, qd);
}

/* stepping 1: BULLET-LIKE substep [UPSTREAM-MEMORY of btMultiBody / btMultiBodyJointMotor / btMultiBodyConstraintSolver as
 * of Bullet 2.8x; nothing in /root/reference pins it -- see the header]
 *   1. external forces: gravity, URDF joint damping (-damping qd) and per-link damping: every link feels, at its centre of
 *      mass, the force -m v_com (k + k |v_com|) and the torque -(I_com w) (k + k |w|) with k = link_damping (Bullet's
 *      m_linearDamping = m_angularDamping = 0.04, used for both the linear and the quadratic coefficient)
 *   2. unconstrained velocity: v* = qd + dt * ABA(q, qd, tau_ext, f_ext)
 *   3. POSITION_CONTROL motor on every joint (setJointMotorControl2 through Joint.control_position,
 *      bullet_scene.py:123-142): a velocity-level constraint with the target
 *          rhs_j = positionGain * (u_j - q_j) / dt + (1 - velocityGain) * v*_j + velocityGain * 0
 *      and the impulse limit force * dt, solved by BULLET_ITERATIONS sweeps of projected Gauss-Seidel over the joints with
 *      accumulated-impulse clamping; A = M(q)^-1 is the response of the joint velocities to unit joint impulses (one ABA
 *      call per joint with qd = 0, g = 0, tau = e_j).
 *   4. qd clamped to +-max_velocity (m_maxCoordinateVelocity = 100), q += qd dt, inelastic stops at the joint limits.
 */
#define BULLET_ITERATIONS 10
static unsigned dyn_substep_bullet(const dyn_batch* b, double* q, double* qd, const float* action) {
    const dyn_params* p = &b->p;
    const double dt = p->timestep, kdamp = p->link_damping;
    double tau[DOF], qdd[DOF], f_ang[DOF][3], f_lin[DOF][3];
    aba_work W;
    body_velocities(p, q, qd, &W);
    for (int i = 0; i < DOF; ++i) {
        double wxc[3], Iw[3], vc[3], f[3], cxf[3];
        v3_cross(W.w[i], p->body_com[i], wxc);                 /* velocity of the centre of mass: v + w x c */
        for (int k = 0; k < 3; ++k) vc[k] = W.v[i][k] + wxc[k];
        const double lin = kdamp + kdamp * sqrt(v3_dot(vc, vc)), ang = kdamp + kdamp * sqrt(v3_dot(W.w[i], W.w[i]));
        for (int k = 0; k < 3; ++k) f[k] = -lin * p->body_mass[i] * vc[k];
        m3_vec(p->body_inertia[i], W.w[i], Iw);
        v3_cross(p->body_com[i], f, cxf);                      /* the force acts at the centre of mass */
        for (int k = 0; k < 3; ++k) { f_ang[i][k] = -ang * Iw[k] + cxf[k]; f_lin[i][k] = f[k]; }
        tau[i] = -p->damping[i] * qd[i];
    }
    aba(p, q, qd, tau, (const double (*)[3])f_ang, (const double (*)[3])f_lin, qdd, &W);
    double v[DOF];
    for (int i = 0; i < DOF; ++i) v[i] = qd[i] + qdd[i] * dt;
    if (p->motor_max_force > 0.0) {
        dyn_params pz = *p;
        pz.gravity = 0.0;
        double A[DOF][DOF], zero[DOF] = {0, 0, 0, 0, 0, 0}, lam[DOF] = {0, 0, 0, 0, 0, 0}, rhs[DOF];
        for (int j = 0; j < DOF; ++j) {
            double e[DOF] = {0, 0, 0, 0, 0, 0}, col[DOF];
            e[j] = 1.0;
            aba(&pz, q, zero, e, NULL, NULL, col, &W);
            for (int i = 0; i < DOF; ++i) A[i][j] = col[i];
            rhs[j] = p->motor_kp * ((double)action[j] - q[j]) / dt + (1.0 - p->motor_kd) * v[j];
        }
        const double max_imp = p->motor_max_force * dt;
        for (int it = 0; it < BULLET_ITERATIONS; ++it)
            for (int j = 0; j < DOF; ++j) {
                double dl = (rhs[j] - v[j]) / A[j][j];
                const double nl = clampd(lam[j] + dl, -max_imp, max_imp);
                dl = nl - lam[j];
                lam[j] = nl;
                for (int i = 0; i < DOF; ++i) v[i] += A[i][j] * dl;
            }
    }
    for (int i = 0; i < DOF; ++i) {
        qd[i] = clampd(v[i], -p->max_velocity, p->max_velocity);
        q[i] += qd[i] * dt;
    }
    return joint_stops(b, q, qd);
}

static void fk_pointer(const dyn_params* p, const double* q, double out[3]) {
    orc_fk_point(p->axis, p->origin_xyz, p->origin_rot, q, DOF - 1, p->tip_xyz, out);
}

static void reset_env(dyn_batch* b, int64_t i, const float* q0, const float* target, uint32_t tick) {
    dyn_env* e = &b->envs[i];
    const dyn_params* p = &b->p;
    float u[16];
    reset_draws(b, b->env_id_base + i, tick, u);
    for (int j = 0; j < DOF; ++j) {
        e->q[j] = (double)(q0 ? q0[j] : uniform32(b->r_lo[j], b->r_hi[j], u[j]));
        e->qd[j] = 0.0; e->a[j] = 0.f;
    }
    for (int k = 0; k < 3; ++k)
        e->target[k] = (double)(target ? target[k] : uniform32((float)p->target_lo[k], (float)p->target_hi[k], u[6 + k]));
    if (p->contact.random_box >= 0) {
        const orc_contact* c = &p->contact;
        for (int k = 0; k < 3; ++k) e->box_e[k] = (double)uniform32((float)c->box_size_lo[k], (float)c->box_size_hi[k], u[9 + k]);
        for (int k = 0; k < 2; ++k) e->box_p[k] = (double)uniform32((float)c->box_pos_lo[k], (float)c->box_pos_hi[k], u[12 + k]);
        e->box_p[2] = e->box_e[2];
    }
    e->potential = 0.0; e->elapsed = 0; e->ep_return = 0.f;
}

/* pioneer_knm_env.py:184-211 on the float64 state (r = q, v = qd, a = the action applied in this step) */
static void observe(const dyn_batch* b, const dyn_env* e, double* o) {
    double ptr[3];
    fk_pointer(&b->p, e->q, ptr);
    for (int i = 0; i < DOF; ++i) {
        const double r = e->q[i], lo = (double)b->r_lo[i], hi = (double)b->r_hi[i], dlo = r - lo, dhi = hi - r;
        const double v = e->qd[i], a = (double)e->a[i];
        o[0 + i] = r;       o[6 + i] = cos(r);       o[12 + i] = sin(r);
        o[18 + i] = lo;     o[24 + i] = cos(lo);     o[30 + i] = sin(lo);
        o[36 + i] = hi;     o[42 + i] = cos(hi);     o[48 + i] = sin(hi);
        o[54 + i] = dlo;    o[60 + i] = cos(dlo);    o[66 + i] = sin(dlo);
        o[72 + i] = dhi;    o[78 + i] = cos(dhi);    o[84 + i] = sin(dhi);
        o[90 + i] = v;      o[96 + i] = cos(v);      o[102 + i] = sin(v);
        o[108 + i] = a;     o[114 + i] = cos(a);     o[120 + i] = sin(a);
    }
    double d2 = 0;
    for (int k = 0; k < 3; ++k) {
        const double diff = e->target[k] - ptr[k];
        o[126 + k] = ptr[k]; o[129 + k] = e->target[k]; o[132 + k] = diff;
        d2 += diff * diff;
    }
    o[135] = sqrt(d2);
    o[136] = e->potential;
}

/* ---- exported API ---- */
dyn_batch* orc_dyn_create(const dyn_params* p, int64_t n, int64_t env_id_base, uint64_t seed) {
    dyn_batch* b = (dyn_batch*)calloc(1, sizeof(dyn_batch));
    if (!b) return NULL;
    b->p = *p; b->n = n; b->env_id_base = env_id_base; b->seed = seed; b->tick = 0;
    for (int j = 0; j < DOF; ++j) { b->r_lo[j] = (float)p->lower[j]; b->r_hi[j] = (float)p->upper[j]; }
    b->envs = (dyn_env*)calloc((size_t)n, sizeof(dyn_env));
    b->s_done = (uint8_t*)calloc((size_t)n, 1);
    if (!b->envs || !b->s_done) { free(b->envs); free(b->s_done); free(b); return NULL; }
    b->stats[4] = -INFINITY; b->stats[5] = INFINITY;
    b->n_threads = 1;
    for (int64_t i = 0; i < n; ++i) reset_env(b, i, NULL, NULL, b->tick);
    b->tick += 1;
    return b;
}

void orc_dyn_set_threads(dyn_batch* b, int32_t n) { b->n_threads = n; }

void orc_dyn_destroy(dyn_batch* b) { if (b) { free(b->envs); free(b->s_done); free(b); } }

void orc_dyn_reset(dyn_batch* b, const int64_t* idx, int64_t n, const float* q0, const float* target, double* obs_out) {
    for (int64_t k = 0; k < n; ++k) {
        const int64_t i = idx ? idx[k] : k;
        reset_env(b, i, q0 ? q0 + k * DOF : NULL, target ? target + k * 3 : NULL, b->tick);
        if (obs_out) observe(b, &b->envs[i], obs_out + k * OBS_DIM);
    }
    b->tick += 1;
}

/* q, qd float64 [n,6] (any NULL): overwrite the joint state (pnr_set_state) */
void orc_dyn_set_state(dyn_batch* b, const double* q, const double* qd) {
    for (int64_t i = 0; i < b->n; ++i)
        for (int j = 0; j < DOF; ++j) {
            if (q) b->envs[i].q[j] = q[i * DOF + j];
            if (qd) b->envs[i].qd[j] = qd[i * DOF + j];
        }
}

/* ONE env's substeps without the env layer (cross-check against dynamics_oracle.py): n_sub substeps from (q, qd).
 * Returns the joints that ran into a stop in any substep (bit i = joint i): an inelastic stop is a discontinuity, so a
 * float32 trajectory that touches it a substep earlier or later legitimately differs by up to qd * dt there. */
uint32_t orc_dyn_substeps(dyn_batch* b, double* q, double* qd, const float* action, int32_t n_sub) {
    unsigned touched = 0;
    for (int s = 0; s < n_sub; ++s) {
        if (b->p.stepping == 1) touched |= dyn_substep_bullet(b, q, qd, action);
        else touched |= dyn_substep_explicit(b, q, qd, action);
    }
    return touched;
}

/* qdd = ABA(q, qd, tau) for one configuration (cross-check against dynamics_oracle.py::aba) */
void orc_dyn_aba(const dyn_batch* b, const double* q, const double* qd, const double* tau, double* qdd) {
    aba_work W;
    aba(&b->p, q, qd, tau, NULL, NULL, qdd, &W);
}

/* min over the segment of the box signed-distance function (contact.h), for the known-answer tests */
double orc_dyn_segment_box(const double* a, const double* bpt, const double* centre, const double* half) {
    return orc_segment_obstacle(ORC_OBST_BOX, centre, half, a, bpt);
}

double orc_dyn_contact_depth(const dyn_batch* b, const double* q, const double* box_p, const double* box_e) {
    return orc_contact_depth(&b->p.contact, b->p.axis, b->p.origin_xyz, b->p.origin_rot, q, box_p, box_e);
}

/* One env step of every env.
 *   actions    float32 [n,6]: PD set points (kp / kd / motor) or joint torques
 *   adopt_q / adopt_qd float32 [n,6] + adopt_mask uint8 [n] (all NULL = free running): after the substeps, rows with a
 *              non-zero mask continue from the given float32 state (what the device reached) instead of the oracle's own
 *   own_q / own_qd float64 [n,6] or NULL: what the oracle's own substeps reached (before adoption)
 *   obs float64 [n,137] or NULL; reward float64 [n]; flags uint8 [n]; depth float64 [n] or NULL (contact depth);
 *   touched uint8 [n] or NULL: bit i = joint i ran into a stop during this step's substeps */
typedef struct {
    dyn_batch* b;
    const float *actions, *adopt_q, *adopt_qd;
    const uint8_t* adopt_mask;
    double *own_q, *own_qd, *obs, *reward, *depth_out;
    uint8_t *flags, *touched;
    int64_t begin, end;
} step_job;

static void* step_range(void* arg) {
    const step_job* J = (const step_job*)arg;
    dyn_batch* b = J->b;
    const dyn_params* p = &b->p;
    const int use_contact = p->contact.n_obstacles > 0 && p->contact.contact_penalty != 0.0;
    for (int64_t i = J->begin; i < J->end; ++i) {
        dyn_env* e = &b->envs[i];
        for (int j = 0; j < DOF; ++j) e->a[j] = J->actions[i * DOF + j];
        const uint32_t touched = orc_dyn_substeps(b, e->q, e->qd, e->a, p->frame_skip);
        if (J->touched) J->touched[i] = (uint8_t)touched;
        if (J->own_q) for (int j = 0; j < DOF; ++j) J->own_q[i * DOF + j] = e->q[j];
        if (J->own_qd) for (int j = 0; j < DOF; ++j) J->own_qd[i * DOF + j] = e->qd[j];
        if (J->adopt_q && (!J->adopt_mask || J->adopt_mask[i]))
            for (int j = 0; j < DOF; ++j) { e->q[j] = (double)J->adopt_q[i * DOF + j]; e->qd[j] = (double)J->adopt_qd[i * DOF + j]; }
        double ptr[3], d2 = 0;
        fk_pointer(p, e->q, ptr);
        for (int k = 0; k < 3; ++k) { const double diff = e->target[k] - ptr[k]; d2 += diff * diff; }
        const double distance = sqrt(d2);
        const double old_potential = e->potential;
        e->potential = (p->award_max - p->award_done) / (distance / p->award_potential_slope + 1);
        int done = distance < p->done_distance;
        double rew = (e->potential - old_potential) + (-p->penalty_step) + (done ? p->award_done : 0.0);
        if (use_contact) {
            const int rb = p->contact.random_box >= 0;
            const double depth = orc_contact_depth(&p->contact, p->axis, p->origin_xyz, p->origin_rot, e->q,
                                                   rb ? e->box_p : NULL, rb ? e->box_e : NULL);
            rew -= p->contact.contact_penalty * depth;
            if (J->depth_out) J->depth_out[i] = depth;
        } else if (J->depth_out) J->depth_out[i] = 0.0;
        if (J->obs) observe(b, e, J->obs + i * OBS_DIM);
        e->elapsed += 1;
        int truncated = 0;
        if (p->max_episode_steps > 0 && e->elapsed >= p->max_episode_steps) { truncated = !done; done = 1; }
        e->ep_return = e->ep_return + (float)rew;
        J->reward[i] = rew;
        J->flags[i] = (uint8_t)((done ? ORC_DONE : 0) | (truncated ? ORC_TRUNCATED : 0));
        b->s_done[i] = (uint8_t)done;
    }
    return NULL;
}

#define ORC_MAX_THREADS 64
void orc_dyn_step(dyn_batch* b, const float* actions, const float* adopt_q, const float* adopt_qd, const uint8_t* adopt_mask,
                  double* own_q, double* own_qd, double* obs, double* reward, uint8_t* flags, double* depth_out,
                  uint8_t* touched) {
    const dyn_params* p = &b->p;
    /* envs are independent: contiguous ranges on plain pthreads (no OpenMP runtime needed on the box) */
    int n_thr = b->n_threads < 1 ? 1 : (b->n_threads > ORC_MAX_THREADS ? ORC_MAX_THREADS : b->n_threads);
    if ((int64_t)n_thr * 64 > b->n) n_thr = (int)((b->n + 63) / 64);
    step_job jobs[ORC_MAX_THREADS];
    pthread_t tid[ORC_MAX_THREADS];
    for (int t = 0; t < n_thr; ++t) {
        step_job j = {b, actions, adopt_q, adopt_qd, adopt_mask, own_q, own_qd, obs, reward, depth_out, flags, touched,
                      b->n * t / n_thr, b->n * (t + 1) / n_thr};
        jobs[t] = j;
    }
    int started = 0;
    for (int t = 1; t < n_thr; ++t, ++started)
        if (pthread_create(&tid[t], NULL, step_range, &jobs[t]) != 0) break;
    step_range(&jobs[0]);
    for (int t = 1; t <= started; ++t) pthread_join(tid[t], NULL);
    for (int t = started + 1; t < n_thr; ++t) step_range(&jobs[t]);     /* thread creation failed: run the rest here */
    /* statistics and auto-reset in env order (deterministic sums) */
    for (int64_t i = 0; i < b->n; ++i) {
        b->stats[6] += 1;
        if (!b->s_done[i]) continue;
        dyn_env* e = &b->envs[i];
        const double ret = (double)e->ep_return, len = (double)e->elapsed;
        b->stats[0] += 1; b->stats[1] += ret; b->stats[2] += len; b->stats[3] += ret * ret;
        if (ret > b->stats[4]) b->stats[4] = ret;
        if (ret < b->stats[5]) b->stats[5] = ret;
        b->stats[7] += (flags[i] & ORC_TRUNCATED) ? 0 : 1;
        if (p->auto_reset) {
            reset_env(b, i, NULL, NULL, b->tick);
            if (obs && p->obs_autoreset) observe(b, e, obs + i * OBS_DIM);
        }
    }
    b->tick += 1;
}

/* q, qd float64 [n,6]; a float32 [n,6]; potential float64 [n]; target float64 [n,3]; t int32 [n]; ep_return float32 [n];
 * box float64 [n,6] (centre, half extents of the per-env box); any NULL */
void orc_dyn_get_state(const dyn_batch* b, double* q, double* qd, float* a, double* potential, double* target, int32_t* t,
                       float* ep_return, double* box) {
    for (int64_t i = 0; i < b->n; ++i) {
        const dyn_env* e = &b->envs[i];
        for (int j = 0; j < DOF; ++j) {
            if (q) q[i * DOF + j] = e->q[j];
            if (qd) qd[i * DOF + j] = e->qd[j];
            if (a) a[i * DOF + j] = e->a[j];
        }
        if (potential) potential[i] = e->potential;
        if (target) for (int k = 0; k < 3; ++k) target[i * 3 + k] = e->target[k];
        if (t) t[i] = e->elapsed;
        if (ep_return) ep_return[i] = e->ep_return;
        if (box) for (int k = 0; k < 3; ++k) { box[i * 6 + k] = e->box_p[k]; box[i * 6 + 3 + k] = e->box_e[k]; }
    }
}

void orc_dyn_stats(const dyn_batch* b, double* out8) { memcpy(out8, b->stats, sizeof b->stats); }
int64_t orc_dyn_sizeof_params(void) { return (int64_t)sizeof(dyn_params); }
