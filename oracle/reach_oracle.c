/*
 * CPU ORACLE (C) — test infrastructure, NOT product code.
 *
 * Plain-C restatement of the reference reach environment, the compiled twin of oracle/reach_oracle.py (same
 * arithmetic, same Philox reset stream, same batched call semantics as include/pioneer_b200.h), fast enough to
 * check the CUDA path at BASELINE.json's full sizes (4,096 envs x 1,000 steps in about a second) and to serve as
 * the "best case" CPU baseline of bench.py.  Only tests/, __graft_entry__ and bench.py's CPU-baseline legs may
 * load it.  Compile with -ffp-contract=off (no FMA contraction): the integrator must round like NumPy does.
 *
 * Parity status: identical to oracle/reach_oracle.py — pinned against the golden vectors recorded from the
 * unmodified reference source (tests/test_oracle_c.py replays them through this file); forward kinematics is
 * Bullet's in the reference and is UNPINNED vs PyBullet (see the Python oracle's header).
 *
 * Reference files followed (relative to /root/reference):
 *   pioneer/envs/pioneer/pioneer_knm_env.py:56-61    bounds
 *   pioneer/envs/pioneer/pioneer_knm_env.py:76-105   reset_world
 *   pioneer/envs/pioneer/pioneer_knm_env.py:111-182  act
 *   pioneer/envs/pioneer/pioneer_knm_env.py:184-211  observe
 *   pioneer/envs/pioneer/pioneer_knm_env.py:232-236  compute_potential
 *   pioneer/envs/bullet/bullet_env.py:187-197        reset / step ordering
 *   pioneer/launch/pioneer_knm_train.py:27           TimeLimit(max_episode_steps=500)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "contact.h"

#define DOF 6
#define OBS_DIM 137
#define ORC_DONE 1
#define ORC_TRUNCATED 2

typedef struct {
    /* kinematic chain (float64), same content as pnr_model */
    double axis[DOF][3], origin_xyz[DOF][3], origin_rot[DOF][9], tip_xyz[3];
    double lower[DOF], upper[DOF];
    /* OracleConfig */
    double max_v_to_r, max_a_to_v, done_distance, award_max, award_done, award_potential_slope, penalty_step;
    double target_lo[3], target_hi[3];
    double timestep;
    int32_t frame_skip;
    int32_t max_episode_steps;
    int32_t legacy;      /* 0 = 'np2' (float32 arithmetic), 1 = 'legacy' (NumPy 1.x promotion) */
    int32_t auto_reset;
    int32_t obs_autoreset; /* 0 = terminal observation, 1 = first observation of the next episode */
    orc_contact contact;   /* obstacle variant (contact.h): reward -= contact_penalty * sum of capsule penetration depths */
} orc_params;

typedef struct {
    float r[DOF], v[DOF], a[DOF];
    double target[3];
    double potential;
    int32_t elapsed;
    float ep_return;
    double box_p[3], box_e[3];  /* the per-env random box of the obstacle variant (centre, half extents) */
} orc_env;

typedef struct {
    orc_params p;
    int64_t n, env_id_base;
    uint64_t seed;
    uint32_t tick;
    float r_lo[DOF], r_hi[DOF], v_max[DOF], a_max[DOF];
    double dt, eps;
    orc_env* envs;
    double stats[8]; /* episodes, sum_return, sum_length, sum_return^2, max, min, env_steps, reached */
} orc_batch;

/* ---- Philox4x32-10 (Salmon et al., SC'11), bit-identical to the Python oracle and the CUDA path ---- */
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

/* 6 joints, 3 target coordinates (reference draw order, pioneer_knm_env.py:80-90), then the per-env box of the obstacle
 * variant: 3 half extents, 2 centre coordinates (pioneer/temp/pioneer_env.py:173-174) */
static void reset_draws(const orc_batch* b, int64_t global_env, uint32_t tick, float u[16]) {
    uint32_t out[16];
    for (uint32_t blk = 0; blk < 4; ++blk) {
        const uint32_t ctr[4] = {(uint32_t)global_env, (uint32_t)((uint64_t)global_env >> 32), tick, blk};
        philox4x32_10(ctr, (uint32_t)b->seed, (uint32_t)(b->seed >> 32), out + 4 * blk);
    }
    for (int i = 0; i < 16; ++i) u[i] = u01(out[i]);
}

/* ---- forward kinematics of 'robot:pointer', float64, tip-to-base with Rodrigues' formula ---- */
static void fk_pointer(const orc_params* p, const float* q, double out[3]) {
    double x = p->tip_xyz[0], y = p->tip_xyz[1], z = p->tip_xyz[2];
    for (int j = DOF - 1; j >= 0; --j) {
        const double c = cos((double)q[j]), s = sin((double)q[j]);
        const double kx = p->axis[j][0], ky = p->axis[j][1], kz = p->axis[j][2];
        const double kp = (kx * x + ky * y + kz * z) * (1.0 - c);
        const double nx = x * c + (ky * z - kz * y) * s + kx * kp;
        const double ny = y * c + (kz * x - kx * z) * s + ky * kp;
        const double nz = z * c + (kx * y - ky * x) * s + kz * kp;
        const double* R = p->origin_rot[j];
        x = p->origin_xyz[j][0] + R[0] * nx + R[1] * ny + R[2] * nz;
        y = p->origin_xyz[j][1] + R[3] * nx + R[4] * ny + R[5] * nz;
        z = p->origin_xyz[j][2] + R[6] * nx + R[7] * ny + R[8] * nz;
    }
    out[0] = x; out[1] = y; out[2] = z;
}

static double potential_of(const orc_params* p, double d) {
    return (p->award_max - p->award_done) / (d / p->award_potential_slope + 1);
}

static float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* pioneer_knm_env.py:111-146 for one joint; volatile stores force the float32 rounding points */
static void integrate_joint(const orc_batch* b, int i, float a0, float v0, float r0, float* v1o, float* r1o) {
    const float vmax = b->v_max[i];
    float v1, r1;
    if (!b->p.legacy) {
        const float dt = (float)b->dt, eps = (float)b->eps;
        const float adt = a0 * dt;
        v1 = v0 + adt;
        float dt_p1 = dt, dt_p2 = 0.f;
        if (v1 > vmax || v1 < -vmax) {
            const float vsat = v1 > vmax ? vmax : -vmax;
            const float num = vsat - v0, den = a0 + eps;
            const float q = num / den;
            dt_p1 = clipf(q, 0.f, dt);
            dt_p2 = dt - dt_p1;
            v1 = vsat;
        }
        const float sum = v0 + v1;
        const float half = 0.5f * sum;
        const float t1 = half * dt_p1, t2 = v1 * dt_p2;
        const float r01 = r0 + t1;
        r1 = r01 + t2;
    } else {
        const double dt = b->dt, eps = b->eps;
        v1 = (float)((double)v0 + (double)a0 * dt);
        double dt_p1 = dt, dt_p2 = 0.0;
        if (v1 > vmax || v1 < -vmax) {
            const float vsat = v1 > vmax ? vmax : -vmax;
            const float num = vsat - v0;
            const double q = (double)num / ((double)a0 + eps);
            dt_p1 = clipd(q, 0.0, dt);
            dt_p2 = dt - dt_p1;
            v1 = vsat;
        }
        const float sum = v0 + v1;
        const double half = 0.5 * (double)sum;
        r1 = (float)(((double)r0 + half * dt_p1) + (double)v1 * dt_p2);
    }
    if (r1 >= b->r_hi[i]) { r1 = b->r_hi[i]; v1 = 0.f; }
    if (r1 <= b->r_lo[i]) { r1 = b->r_lo[i]; v1 = 0.f; }
    *v1o = v1; *r1o = r1;
}

static void reset_env(orc_batch* b, int64_t i, const float* q0, const float* target, uint32_t tick) {
    orc_env* e = &b->envs[i];
    float u[16];
    reset_draws(b, b->env_id_base + i, tick, u);
    for (int j = 0; j < DOF; ++j) {
        if (q0) e->r[j] = q0[j];
        else { const float span = b->r_hi[j] - b->r_lo[j]; const float m = span * u[j]; e->r[j] = b->r_lo[j] + m; }
        e->v[j] = 0.f; e->a[j] = 0.f;
    }
    for (int k = 0; k < 3; ++k) {
        if (target) e->target[k] = (double)target[k];
        else {
            const float lo = (float)b->p.target_lo[k], hi = (float)b->p.target_hi[k];
            const float span = hi - lo; const float m = span * u[6 + k];
            e->target[k] = (double)(float)(lo + m);
        }
    }
    if (b->p.contact.random_box >= 0) {
        const orc_contact* c = &b->p.contact;
        for (int k = 0; k < 3; ++k) {
            const float lo = (float)c->box_size_lo[k], span = (float)c->box_size_hi[k] - lo; const float m = span * u[9 + k];
            e->box_e[k] = (double)(float)(lo + m);
        }
        for (int k = 0; k < 2; ++k) {
            const float lo = (float)c->box_pos_lo[k], span = (float)c->box_pos_hi[k] - lo; const float m = span * u[12 + k];
            e->box_p[k] = (double)(float)(lo + m);
        }
        e->box_p[2] = e->box_e[2];                              /* the box stands on z = 0 */
    }
    e->potential = 0.0; e->elapsed = 0; e->ep_return = 0.f;
}

/* pioneer_knm_env.py:184-211; cos/sin of float32 values are float32 evaluations (np.cos on a float32 array) */
static void observe(const orc_batch* b, const orc_env* e, double* o) {
    double ptr[3];
    fk_pointer(&b->p, e->r, ptr);
    for (int i = 0; i < DOF; ++i) {
        const float r = e->r[i], dlo = r - b->r_lo[i], dhi = b->r_hi[i] - r;
        o[0 + i] = r;             o[6 + i] = cosf(r);             o[12 + i] = sinf(r);
        o[18 + i] = b->r_lo[i];   o[24 + i] = cosf(b->r_lo[i]);   o[30 + i] = sinf(b->r_lo[i]);
        o[36 + i] = b->r_hi[i];   o[42 + i] = cosf(b->r_hi[i]);   o[48 + i] = sinf(b->r_hi[i]);
        o[54 + i] = dlo;          o[60 + i] = cosf(dlo);          o[66 + i] = sinf(dlo);
        o[72 + i] = dhi;          o[78 + i] = cosf(dhi);          o[84 + i] = sinf(dhi);
        o[90 + i] = e->v[i];      o[96 + i] = cosf(e->v[i]);      o[102 + i] = sinf(e->v[i]);
        o[108 + i] = e->a[i];     o[114 + i] = cosf(e->a[i]);     o[120 + i] = sinf(e->a[i]);
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
orc_batch* orc_create(const orc_params* p, int64_t n, int64_t env_id_base, uint64_t seed) {
    orc_batch* b = (orc_batch*)calloc(1, sizeof(orc_batch));
    if (!b) return NULL;
    b->p = *p; b->n = n; b->env_id_base = env_id_base; b->seed = seed; b->tick = 0;
    for (int j = 0; j < DOF; ++j) {
        b->r_lo[j] = (float)p->lower[j]; b->r_hi[j] = (float)p->upper[j];
        const float span = b->r_hi[j] - b->r_lo[j];
        b->v_max[j] = (float)p->max_v_to_r * span;
        b->a_max[j] = (float)p->max_a_to_v * b->v_max[j];
    }
    b->dt = p->timestep * p->frame_skip;
    b->eps = 1e-5;
    b->envs = (orc_env*)calloc((size_t)n, sizeof(orc_env));
    if (!b->envs) { free(b); return NULL; }
    b->stats[4] = -INFINITY; b->stats[5] = INFINITY;
    for (int64_t i = 0; i < n; ++i) reset_env(b, i, NULL, NULL, b->tick);
    b->tick += 1;
    return b;
}

void orc_destroy(orc_batch* b) { if (b) { free(b->envs); free(b); } }

void orc_bounds(const orc_batch* b, float* r_lo, float* r_hi, float* v_max, float* a_max) {
    memcpy(r_lo, b->r_lo, sizeof b->r_lo); memcpy(r_hi, b->r_hi, sizeof b->r_hi);
    memcpy(v_max, b->v_max, sizeof b->v_max); memcpy(a_max, b->a_max, sizeof b->a_max);
}

/* idx NULL = all envs; q0 [n,6] / target [n,3] float32 or NULL = Philox; obs_out [n,137] float64 or NULL */
void orc_reset(orc_batch* b, const int64_t* idx, int64_t n, const float* q0, const float* target, double* obs_out) {
    for (int64_t k = 0; k < n; ++k) {
        const int64_t i = idx ? idx[k] : k;
        reset_env(b, i, q0 ? q0 + k * DOF : NULL, target ? target + k * 3 : NULL, b->tick);
        if (obs_out) observe(b, &b->envs[i], obs_out + k * OBS_DIM);
    }
    b->tick += 1;
}

void orc_observe(const orc_batch* b, double* obs_out) {
    for (int64_t i = 0; i < b->n; ++i) observe(b, &b->envs[i], obs_out + i * OBS_DIM);
}

/* one env step for every env: act() + observe() through TimeLimit, statistics, auto-reset.
 * actions float32 [n,6]; obs float64 [n,137] or NULL; reward float64 [n]; flags uint8 [n]. */
void orc_step(orc_batch* b, const float* actions, double* obs, double* reward, uint8_t* flags) {
    const orc_params* p = &b->p;
    for (int64_t i = 0; i < b->n; ++i) {
        orc_env* e = &b->envs[i];
        for (int j = 0; j < DOF; ++j) {
            float v1, r1;
            integrate_joint(b, j, e->a[j], e->v[j], e->r[j], &v1, &r1);
            e->v[j] = v1; e->r[j] = r1;
        }
        for (int j = 0; j < DOF; ++j) e->a[j] = actions[i * DOF + j];
        double ptr[3], d2 = 0;
        fk_pointer(p, e->r, ptr);
        for (int k = 0; k < 3; ++k) { const double diff = e->target[k] - ptr[k]; d2 += diff * diff; }
        const double distance = sqrt(d2);
        const double old_potential = e->potential;
        e->potential = potential_of(p, distance);
        int done = distance < p->done_distance;
        double rew = (e->potential - old_potential) + (-p->penalty_step) + (done ? p->award_done : 0.0);
        if (p->contact.n_obstacles > 0 && p->contact.contact_penalty != 0.0) {
            double qd_[DOF];
            for (int j = 0; j < DOF; ++j) qd_[j] = (double)e->r[j];
            const int rb = p->contact.random_box >= 0;
            rew -= p->contact.contact_penalty * orc_contact_depth(&p->contact, p->axis, p->origin_xyz, p->origin_rot, qd_,
                                                                  rb ? e->box_p : NULL, rb ? e->box_e : NULL);
        }
        if (obs) observe(b, e, obs + i * OBS_DIM);
        e->elapsed += 1;
        int truncated = 0;
        if (p->max_episode_steps > 0 && e->elapsed >= p->max_episode_steps) { truncated = !done; done = 1; }
        e->ep_return = e->ep_return + (float)rew;
        reward[i] = rew;
        flags[i] = (uint8_t)((done ? ORC_DONE : 0) | (truncated ? ORC_TRUNCATED : 0));
        b->stats[6] += 1;
        if (done) {
            const double ret = (double)e->ep_return, len = (double)e->elapsed;
            b->stats[0] += 1; b->stats[1] += ret; b->stats[2] += len; b->stats[3] += ret * ret;
            if (ret > b->stats[4]) b->stats[4] = ret;
            if (ret < b->stats[5]) b->stats[5] = ret;
            b->stats[7] += truncated ? 0 : 1;
            if (p->auto_reset) {
                reset_env(b, i, NULL, NULL, b->tick);
                if (obs && p->obs_autoreset) observe(b, e, obs + i * OBS_DIM);
            }
        }
    }
    b->tick += 1;
}

/* r, v, a float32 [n,6]; potential float64 [n]; target float64 [n,3]; t int32 [n]; ep_return float32 [n]; any NULL */
void orc_get_state(const orc_batch* b, float* r, float* v, float* a, double* potential, double* target, int32_t* t,
                   float* ep_return) {
    for (int64_t i = 0; i < b->n; ++i) {
        const orc_env* e = &b->envs[i];
        for (int j = 0; j < DOF; ++j) {
            if (r) r[i * DOF + j] = e->r[j];
            if (v) v[i * DOF + j] = e->v[j];
            if (a) a[i * DOF + j] = e->a[j];
        }
        if (potential) potential[i] = e->potential;
        if (target) for (int k = 0; k < 3; ++k) target[i * 3 + k] = e->target[k];
        if (t) t[i] = e->elapsed;
        if (ep_return) ep_return[i] = e->ep_return;
    }
}

/* box float64 [n,6]: centre and half extents of every env's random box */
void orc_get_boxes(const orc_batch* b, double* box) {
    for (int64_t i = 0; i < b->n; ++i)
        for (int k = 0; k < 3; ++k) { box[i * 6 + k] = b->envs[i].box_p[k]; box[i * 6 + 3 + k] = b->envs[i].box_e[k]; }
}

void orc_stats(const orc_batch* b, double* out8) { memcpy(out8, b->stats, sizeof b->stats); }
int64_t orc_sizeof_params(void) { return (int64_t)sizeof(orc_params); }
