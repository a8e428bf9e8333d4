// Dynamic (Tier-B) mode: Featherstone articulated-body forward dynamics of the 6-joint serial chain in float32,
// one thread per env, spatial 6x6 algebra written out in 3-vector / 3x3-block form and kept in registers.
//
// Spatial vectors are [angular; linear] in the coordinates of the body's own moving frame (origin on the joint
// axis).  An articulated inertia is the symmetric 6x6 [[I, H], [H^T, M]]: n = I w + H v, f = H^T w + M v.
// The frame of body i sits at p_i (origin_xyz) in its parent, rotated by R_i = O_i Rot(axis_i, q_i).
// Semantics of a substep: oracle/dynamics_oracle.py (the float64 6x6-matrix restatement these kernels are tested
// against).  The reference env never runs Bullet's dynamics with non-zero inputs: this mode
// is the north_star extension, parity unpinned vs PyBullet; the fact the reference does pin -- zero velocity,
// gravity and torque leave the state bit-unchanged -- holds exactly (every term is an exact zero).
#pragma once
#include "pnr_kernels.cuh"

struct V3 { float x, y, z; };
struct Sym3 { float xx, xy, xz, yy, yz, zz; };
struct Mat3 { float m[9]; };          // row-major

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ V3 symmul(const Sym3& s, V3 v) {
    return v3(fmaf(s.xx, v.x, fmaf(s.xy, v.y, s.xz * v.z)), fmaf(s.xy, v.x, fmaf(s.yy, v.y, s.yz * v.z)),
              fmaf(s.xz, v.x, fmaf(s.yz, v.y, s.zz * v.z)));
}
__device__ __forceinline__ V3 matmul(const Mat3& a, V3 v) {
    return v3(fmaf(a.m[0], v.x, fmaf(a.m[1], v.y, a.m[2] * v.z)), fmaf(a.m[3], v.x, fmaf(a.m[4], v.y, a.m[5] * v.z)),
              fmaf(a.m[6], v.x, fmaf(a.m[7], v.y, a.m[8] * v.z)));
}
__device__ __forceinline__ V3 matTmul(const Mat3& a, V3 v) {
    return v3(fmaf(a.m[0], v.x, fmaf(a.m[3], v.y, a.m[6] * v.z)), fmaf(a.m[1], v.x, fmaf(a.m[4], v.y, a.m[7] * v.z)),
              fmaf(a.m[2], v.x, fmaf(a.m[5], v.y, a.m[8] * v.z)));
}

// CHAIN selects how the joint axes are known.  PNR_CHAIN_GENERIC reads axis codes / origin rotations from the
// constant bank (warp-uniform branches, any serial 6-revolute chain).  PNR_CHAIN_PIONEER bakes in the shipped
// robot (axes Z Y Y X Y X, all positive, no origin rotation; assets/pioneer_reach_6dof.urdf = the reference's
// pioneer_knm_6dof.urdf:204-275): inside the fully unrolled joint loops every code is a compile-time constant, the
// branches and zero terms fold away and the substep body shrinks from ~12,000 to ~3,000 SASS instructions -- the
// generic body does not fit the instruction cache (71 % of warp stalls were instruction fetch, profiles/r01_dyn_v1).
#define PNR_CHAIN_GENERIC 0
#define PNR_CHAIN_PIONEER 1

template <int CHAIN>
__device__ __forceinline__ int pnr_code(const PnrParams& p, int j) {
    if (CHAIN == PNR_CHAIN_PIONEER) return j == 0 ? PNR_AXIS_Z : ((j == 3 || j == 5) ? PNR_AXIS_X : PNR_AXIS_Y);
    return p.axis_code[j];
}
template <int CHAIN>
__device__ __forceinline__ V3 pnr_axis(const PnrParams& p, int j) {
    if (CHAIN == PNR_CHAIN_PIONEER) {
        const int c = pnr_code<CHAIN>(p, j);
        return v3(c == PNR_AXIS_X ? 1.f : 0.f, c == PNR_AXIS_Y ? 1.f : 0.f, c == PNR_AXIS_Z ? 1.f : 0.f);
    }
    return v3(p.axis[j][0], p.axis[j][1], p.axis[j][2]);
}
// axis . v,  axis * s,  v x (axis * s),  I axis (a column),  H^T axis (a row): single components when the axis is known
template <int CHAIN>
__device__ __forceinline__ float pnr_axis_dot(const PnrParams& p, int j, V3 v) {
    if (CHAIN == PNR_CHAIN_PIONEER) {
        const int c = pnr_code<CHAIN>(p, j);
        return c == PNR_AXIS_X ? v.x : (c == PNR_AXIS_Y ? v.y : v.z);
    }
    return dot(pnr_axis<CHAIN>(p, j), v);
}
template <int CHAIN>
__device__ __forceinline__ V3 pnr_axis_scaled(const PnrParams& p, int j, float s) {
    if (CHAIN == PNR_CHAIN_PIONEER) {
        const int c = pnr_code<CHAIN>(p, j);
        return v3(c == PNR_AXIS_X ? s : 0.f, c == PNR_AXIS_Y ? s : 0.f, c == PNR_AXIS_Z ? s : 0.f);
    }
    return pnr_axis<CHAIN>(p, j) * s;
}
template <int CHAIN>
__device__ __forceinline__ V3 pnr_cross_axis(const PnrParams& p, int j, V3 a, float s) {     // a x (axis * s)
    if (CHAIN == PNR_CHAIN_PIONEER) {
        const int c = pnr_code<CHAIN>(p, j);
        if (c == PNR_AXIS_X) return v3(0.f, a.z * s, -a.y * s);
        if (c == PNR_AXIS_Y) return v3(-a.z * s, 0.f, a.x * s);
        return v3(a.y * s, -a.x * s, 0.f);
    }
    return cross(a, pnr_axis<CHAIN>(p, j) * s);
}
template <int CHAIN>
__device__ __forceinline__ V3 pnr_sym_axis(const PnrParams& p, int j, const Sym3& m) {        // m * axis
    if (CHAIN == PNR_CHAIN_PIONEER) {
        const int c = pnr_code<CHAIN>(p, j);
        if (c == PNR_AXIS_X) return v3(m.xx, m.xy, m.xz);
        if (c == PNR_AXIS_Y) return v3(m.xy, m.yy, m.yz);
        return v3(m.xz, m.yz, m.zz);
    }
    return symmul(m, pnr_axis<CHAIN>(p, j));
}
template <int CHAIN>
__device__ __forceinline__ V3 pnr_matT_axis(const PnrParams& p, int j, const Mat3& a) {       // a^T * axis
    if (CHAIN == PNR_CHAIN_PIONEER) {
        const int c = pnr_code<CHAIN>(p, j);
        return v3(a.m[3 * c + 0], a.m[3 * c + 1], a.m[3 * c + 2]);
    }
    return matTmul(a, pnr_axis<CHAIN>(p, j));
}

// origin_j x v.  The shipped robot's joint origins lie on one coordinate axis of the parent frame (joint 0 and 5: zero,
// 1 and 2: along z, 3: along y, 4: along x; pioneer_knm_6dof.urdf:209-264), so the cross product is two multiplies.
template <int CHAIN>
__device__ __forceinline__ V3 pnr_origin_cross(const PnrParams& p, int j, V3 v) {
    if (CHAIN == PNR_CHAIN_PIONEER) {
        if (j == 0 || j == 5) return v3(0.f, 0.f, 0.f);
        if (j == 1 || j == 2) { const float z = p.origin_xyz[j][2]; return v3(-z * v.y, z * v.x, 0.f); }
        if (j == 3) { const float y = p.origin_xyz[j][1]; return v3(y * v.z, 0.f, -y * v.x); }
        const float x = p.origin_xyz[j][0];
        return v3(0.f, -x * v.z, x * v.y);
    }
    return cross(v3(p.origin_xyz[j][0], p.origin_xyz[j][1], p.origin_xyz[j][2]), v);
}

// rotation about joint j's axis by the angle whose (sin, cos) are given; axis-aligned axes cost 4 FMA.
template <int CHAIN>
__device__ __forceinline__ V3 pnr_axis_rot(const PnrParams& p, int j, float s, float c, V3 v) {
    const int code = pnr_code<CHAIN>(p, j);
    if (CHAIN != PNR_CHAIN_PIONEER) s *= p.axis_sign[j];
    if (code == PNR_AXIS_X) return v3(v.x, fmaf(c, v.y, -s * v.z), fmaf(s, v.y, c * v.z));
    if (code == PNR_AXIS_Y) return v3(fmaf(c, v.x, s * v.z), v.y, fmaf(-s, v.x, c * v.z));
    if (code == PNR_AXIS_Z) return v3(fmaf(c, v.x, -s * v.y), fmaf(s, v.x, c * v.y), v.z);
    const V3 k = v3(p.axis[j][0], p.axis[j][1], p.axis[j][2]);          // Rodrigues
    const float kv = dot(k, v) * (1.f - c);
    const V3 kxv = cross(k, v);
    return v3(fmaf(v.x, c, fmaf(kxv.x, s, k.x * kv)), fmaf(v.y, c, fmaf(kxv.y, s, k.y * kv)),
              fmaf(v.z, c, fmaf(kxv.z, s, k.z * kv)));
}

// child -> parent coordinates: R_j v = O_j Rot(axis_j, q_j) v
template <int CHAIN>
__device__ __forceinline__ V3 pnr_rot(const PnrParams& p, int j, float s, float c, V3 v) {
    V3 r = pnr_axis_rot<CHAIN>(p, j, s, c, v);
    if (CHAIN != PNR_CHAIN_PIONEER && p.origin_has_rot[j]) {
        const float* O = p.origin_rot[j];
        r = v3(O[0] * r.x + O[1] * r.y + O[2] * r.z, O[3] * r.x + O[4] * r.y + O[5] * r.z, O[6] * r.x + O[7] * r.y + O[8] * r.z);
    }
    return r;
}

// parent -> child coordinates: R_j^T v
template <int CHAIN>
__device__ __forceinline__ V3 pnr_rot_t(const PnrParams& p, int j, float s, float c, V3 v) {
    if (CHAIN != PNR_CHAIN_PIONEER && p.origin_has_rot[j]) {
        const float* O = p.origin_rot[j];
        v = v3(O[0] * v.x + O[3] * v.y + O[6] * v.z, O[1] * v.x + O[4] * v.y + O[7] * v.z, O[2] * v.x + O[5] * v.y + O[8] * v.z);
    }
    return pnr_axis_rot<CHAIN>(p, j, -s, c, v);
}

// B' = R B R^T for a general 3x3 block: rotate the columns, then the rows
template <int CHAIN>
__device__ __forceinline__ Mat3 pnr_rot_block(const PnrParams& p, int j, float s, float c, const Mat3& b) {
    const V3 c0 = pnr_rot<CHAIN>(p, j, s, c, v3(b.m[0], b.m[3], b.m[6]));
    const V3 c1 = pnr_rot<CHAIN>(p, j, s, c, v3(b.m[1], b.m[4], b.m[7]));
    const V3 c2 = pnr_rot<CHAIN>(p, j, s, c, v3(b.m[2], b.m[5], b.m[8]));
    // rows of C = R B are (c0.x c1.x c2.x), ...; row i of B' = R * (row i of C)
    const V3 r0 = pnr_rot<CHAIN>(p, j, s, c, v3(c0.x, c1.x, c2.x));
    const V3 r1 = pnr_rot<CHAIN>(p, j, s, c, v3(c0.y, c1.y, c2.y));
    const V3 r2 = pnr_rot<CHAIN>(p, j, s, c, v3(c0.z, c1.z, c2.z));
    Mat3 o = {{r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z}};
    return o;
}

__device__ __forceinline__ Mat3 sym_to_mat(const Sym3& s) {
    Mat3 o = {{s.xx, s.xy, s.xz, s.xy, s.yy, s.yz, s.xz, s.yz, s.zz}};
    return o;
}
__device__ __forceinline__ Sym3 mat_to_sym(const Mat3& a) {     // a is symmetric up to rounding: average the pairs
    Sym3 o = {a.m[0], 0.5f * (a.m[1] + a.m[3]), 0.5f * (a.m[2] + a.m[6]), a.m[4], 0.5f * (a.m[5] + a.m[7]), a.m[8]};
    return o;
}

struct PnrDynWork {                 // per-joint quantities kept between the three passes
    float sn[PNR_DOF], cs[PNR_DOF];
    V3 c_ang[PNR_DOF], c_lin[PNR_DOF];      // velocity-product accelerations
    V3 p_ang[PNR_DOF], p_lin[PNR_DOF];      // bias forces
    V3 u_ang[PNR_DOF], u_lin[PNR_DOF];      // U = I^A S
    float dinv[PNR_DOF], u[PNR_DOF];
};

// qdd = ABA(q, qd, tau); gravity acts along -z of the base frame
template <int CHAIN>
__device__ __forceinline__ void pnr_aba(const PnrParams& p, const float (&q)[PNR_DOF], const float (&qd)[PNR_DOF],
                                        const float (&tau)[PNR_DOF], float (&qdd)[PNR_DOF]) {
    PnrDynWork w;
    // ---- pass 1 (base -> tip): velocities, velocity-product accelerations, bias forces
    V3 om = v3(0.f, 0.f, 0.f), vl = v3(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) {
        pnr_sincos_fast(q[i], w.sn[i], w.cs[i]);                       // q is inside the joint limits
        const V3 t = vl - pnr_origin_cross<CHAIN>(p, i, om);
        om = pnr_rot_t<CHAIN>(p, i, w.sn[i], w.cs[i], om);
        vl = pnr_rot_t<CHAIN>(p, i, w.sn[i], w.cs[i], t);
        om = om + pnr_axis_scaled<CHAIN>(p, i, qd[i]);
        w.c_ang[i] = pnr_cross_axis<CHAIN>(p, i, om, qd[i]);
        w.c_lin[i] = pnr_cross_axis<CHAIN>(p, i, vl, qd[i]);
        const Sym3 Io = {p.dyn_io[i][0], p.dyn_io[i][1], p.dyn_io[i][2], p.dyn_io[i][3], p.dyn_io[i][4], p.dyn_io[i][5]};
        const V3 mc = v3(p.dyn_mc[i][0], p.dyn_mc[i][1], p.dyn_mc[i][2]);
        const V3 n = symmul(Io, om) + cross(mc, vl);
        const V3 f = vl * p.dyn_mass[i] - cross(mc, om);
        w.p_ang[i] = cross(om, n) + cross(vl, f);
        w.p_lin[i] = cross(om, f);
    }
    // ---- pass 2 (tip -> base): articulated inertias and bias forces
    Sym3 I, M;
    Mat3 H;
    V3 pa_ang, pa_lin;
#pragma unroll
    for (int i = PNR_DOF - 1; i >= 0; --i) {
        const Sym3 Io = {p.dyn_io[i][0], p.dyn_io[i][1], p.dyn_io[i][2], p.dyn_io[i][3], p.dyn_io[i][4], p.dyn_io[i][5]};
        const V3 mc = v3(p.dyn_mc[i][0], p.dyn_mc[i][1], p.dyn_mc[i][2]);
        const float m = p.dyn_mass[i];
        if (i == PNR_DOF - 1) {
            I = Io;
            Mat3 h0 = {{0.f, -mc.z, mc.y, mc.z, 0.f, -mc.x, -mc.y, mc.x, 0.f}};
            H = h0;
            Sym3 m0 = {m, 0.f, 0.f, m, 0.f, m};
            M = m0;
            pa_ang = w.p_ang[i];
            pa_lin = w.p_lin[i];
        } else {                                                        // own rigid body + what the child handed up
            I.xx += Io.xx; I.xy += Io.xy; I.xz += Io.xz; I.yy += Io.yy; I.yz += Io.yz; I.zz += Io.zz;
            H.m[1] -= mc.z; H.m[2] += mc.y; H.m[3] += mc.z; H.m[5] -= mc.x; H.m[6] -= mc.y; H.m[7] += mc.x;
            M.xx += m; M.yy += m; M.zz += m;
            pa_ang = pa_ang + w.p_ang[i];
            pa_lin = pa_lin + w.p_lin[i];
        }
        const V3 Ua = pnr_sym_axis<CHAIN>(p, i, I), Ul = pnr_matT_axis<CHAIN>(p, i, H);
        const float dinv = 1.f / pnr_axis_dot<CHAIN>(p, i, Ua);
        const float u = tau[i] - pnr_axis_dot<CHAIN>(p, i, pa_ang);
        w.u_ang[i] = Ua; w.u_lin[i] = Ul; w.dinv[i] = dinv; w.u[i] = u;
        if (i > 0) {
            // I^a = I^A - U U^T / d
            const V3 Uad = Ua * dinv, Uld = Ul * dinv;
            I.xx -= Ua.x * Uad.x; I.xy -= Ua.x * Uad.y; I.xz -= Ua.x * Uad.z;
            I.yy -= Ua.y * Uad.y; I.yz -= Ua.y * Uad.z; I.zz -= Ua.z * Uad.z;
            H.m[0] -= Uad.x * Ul.x; H.m[1] -= Uad.x * Ul.y; H.m[2] -= Uad.x * Ul.z;
            H.m[3] -= Uad.y * Ul.x; H.m[4] -= Uad.y * Ul.y; H.m[5] -= Uad.y * Ul.z;
            H.m[6] -= Uad.z * Ul.x; H.m[7] -= Uad.z * Ul.y; H.m[8] -= Uad.z * Ul.z;
            M.xx -= Ul.x * Uld.x; M.xy -= Ul.x * Uld.y; M.xz -= Ul.x * Uld.z;
            M.yy -= Ul.y * Uld.y; M.yz -= Ul.y * Uld.z; M.zz -= Ul.z * Uld.z;
            // p^a = p^A + I^a c + U u / d
            const float ud = u * dinv;
            pa_ang = pa_ang + symmul(I, w.c_ang[i]) + matmul(H, w.c_lin[i]) + Ua * ud;
            pa_lin = pa_lin + matTmul(H, w.c_ang[i]) + symmul(M, w.c_lin[i]) + Ul * ud;
            // hand I^a, p^a up to the parent: rotate by R_i, then move the reference point by p_i
            const float s = w.sn[i], c = w.cs[i];
            const Mat3 Ir = pnr_rot_block<CHAIN>(p, i, s, c, sym_to_mat(I));
            const Mat3 Hr = pnr_rot_block<CHAIN>(p, i, s, c, H);
            const Mat3 Mr = pnr_rot_block<CHAIN>(p, i, s, c, sym_to_mat(M));
            // A = p x M' (column-wise); H'' = H' + A
            const V3 a0 = pnr_origin_cross<CHAIN>(p, i, v3(Mr.m[0], Mr.m[3], Mr.m[6]));
            const V3 a1 = pnr_origin_cross<CHAIN>(p, i, v3(Mr.m[1], Mr.m[4], Mr.m[7]));
            const V3 a2 = pnr_origin_cross<CHAIN>(p, i, v3(Mr.m[2], Mr.m[5], Mr.m[8]));
            const Mat3 A = {{a0.x, a1.x, a2.x, a0.y, a1.y, a2.y, a0.z, a1.z, a2.z}};
            // K = p x H'^T (columns of K = p x rows of H'); A P^T has rows p x (rows of A)
            const V3 k0 = pnr_origin_cross<CHAIN>(p, i, v3(Hr.m[0], Hr.m[1], Hr.m[2]));
            const V3 k1 = pnr_origin_cross<CHAIN>(p, i, v3(Hr.m[3], Hr.m[4], Hr.m[5]));
            const V3 k2 = pnr_origin_cross<CHAIN>(p, i, v3(Hr.m[6], Hr.m[7], Hr.m[8]));
            const Mat3 K = {{k0.x, k1.x, k2.x, k0.y, k1.y, k2.y, k0.z, k1.z, k2.z}};
            const V3 q0 = pnr_origin_cross<CHAIN>(p, i, v3(A.m[0], A.m[1], A.m[2]));
            const V3 q1 = pnr_origin_cross<CHAIN>(p, i, v3(A.m[3], A.m[4], A.m[5]));
            const V3 q2 = pnr_origin_cross<CHAIN>(p, i, v3(A.m[6], A.m[7], A.m[8]));
            Mat3 In;
            In.m[0] = Ir.m[0] + 2.f * K.m[0] + q0.x;
            In.m[1] = Ir.m[1] + K.m[1] + K.m[3] + q0.y;
            In.m[2] = Ir.m[2] + K.m[2] + K.m[6] + q0.z;
            In.m[3] = Ir.m[3] + K.m[3] + K.m[1] + q1.x;
            In.m[4] = Ir.m[4] + 2.f * K.m[4] + q1.y;
            In.m[5] = Ir.m[5] + K.m[5] + K.m[7] + q1.z;
            In.m[6] = Ir.m[6] + K.m[6] + K.m[2] + q2.x;
            In.m[7] = Ir.m[7] + K.m[7] + K.m[5] + q2.y;
            In.m[8] = Ir.m[8] + 2.f * K.m[8] + q2.z;
            I = mat_to_sym(In);
#pragma unroll
            for (int k = 0; k < 9; ++k) H.m[k] = Hr.m[k] + A.m[k];
            M = mat_to_sym(Mr);
            const V3 fp = pnr_rot<CHAIN>(p, i, s, c, pa_lin);
            pa_ang = pnr_rot<CHAIN>(p, i, s, c, pa_ang) + pnr_origin_cross<CHAIN>(p, i, fp);
            pa_lin = fp;
        }
    }
    // ---- pass 3 (base -> tip): accelerations.  The base "accelerates" by -g: a_lin = (0, 0, +gravity)
    V3 aa = v3(0.f, 0.f, 0.f), al = v3(0.f, 0.f, p.dyn_gravity);
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) {
        const V3 t = al - pnr_origin_cross<CHAIN>(p, i, aa);
        aa = pnr_rot_t<CHAIN>(p, i, w.sn[i], w.cs[i], aa) + w.c_ang[i];
        al = pnr_rot_t<CHAIN>(p, i, w.sn[i], w.cs[i], t) + w.c_lin[i];
        qdd[i] = (w.u[i] - dot(w.u_ang[i], aa) - dot(w.u_lin[i], al)) * w.dinv[i];
        aa = aa + pnr_axis_scaled<CHAIN>(p, i, qdd[i]);
    }
}

// joint PD / torque control with effort clamping, then viscous joint damping (oracle/dynamics_oracle.py::control_torque)
__device__ __forceinline__ float pnr_control_torque(const PnrParams& p, int i, float action, float q, float qd) {
    float tau = p.dyn_use_pd ? fmaf(p.dyn_kp, action - q, -p.dyn_kd * qd) : action;
    const float lim = p.dyn_tau_max[i];
    tau = fminf(fmaxf(tau, -lim), lim);
    return fmaf(-p.dyn_damping[i], qd, tau);
}

// frame_skip substeps of semi-implicit Euler: qd += qdd dt; q += qd dt; inelastic stops at the joint limits
template <int CHAIN>
__device__ __forceinline__ void pnr_dynamic_substeps(const PnrParams& p, float (&q)[PNR_DOF], float (&qd)[PNR_DOF],
                                                     const float (&action)[PNR_DOF]) {
#pragma unroll 1
    for (int sub = 0; sub < p.dyn_frame_skip; ++sub) {
        float tau[PNR_DOF], qdd[PNR_DOF];
#pragma unroll
        for (int i = 0; i < PNR_DOF; ++i) tau[i] = pnr_control_torque(p, i, action[i], q[i], qd[i]);
        pnr_aba<CHAIN>(p, q, qd, tau, qdd);
#pragma unroll
        for (int i = 0; i < PNR_DOF; ++i) {
            float v = fmaf(qdd[i], p.dyn_dt, qd[i]);
            float x = fmaf(v, p.dyn_dt, q[i]);
            if (x > p.r_hi[i]) { x = p.r_hi[i]; if (v > 0.f) v = 0.f; }
            if (x < p.r_lo[i]) { x = p.r_lo[i]; if (v < 0.f) v = 0.f; }
            q[i] = x; qd[i] = v;
        }
    }
}
