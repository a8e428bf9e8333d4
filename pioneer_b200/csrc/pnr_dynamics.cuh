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
#include <cmath>
#include "pnr_device.cuh"
#include "pnr_trig.cuh"

// host + device: tests/csrc/aba_check.cu compiles this header for the CPU and checks it against the float64 oracle
#define PNR_HD __host__ __device__ __forceinline__

struct V3 { float x, y, z; };
struct Sym3 { float xx, xy, xz, yy, yz, zz; };
struct Mat3 { float m[9]; };          // row-major

PNR_HD V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
PNR_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
PNR_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
PNR_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
PNR_HD float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
PNR_HD V3 cross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
PNR_HD V3 symmul(const Sym3& s, V3 v) {
    return v3(fmaf(s.xx, v.x, fmaf(s.xy, v.y, s.xz * v.z)), fmaf(s.xy, v.x, fmaf(s.yy, v.y, s.yz * v.z)),
              fmaf(s.xz, v.x, fmaf(s.yz, v.y, s.zz * v.z)));
}
PNR_HD V3 matmul(const Mat3& a, V3 v) {
    return v3(fmaf(a.m[0], v.x, fmaf(a.m[1], v.y, a.m[2] * v.z)), fmaf(a.m[3], v.x, fmaf(a.m[4], v.y, a.m[5] * v.z)),
              fmaf(a.m[6], v.x, fmaf(a.m[7], v.y, a.m[8] * v.z)));
}
PNR_HD V3 matTmul(const Mat3& a, V3 v) {
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
// PNR_CHAIN_PIONEER_ISO: additionally the inertials the shipped URDF gives every link but the tip (the stub
// <inertial> blocks of pioneer_knm_6dof.urdf: centre of mass on the frame origin, isotropic inertia iota * 1): for such
// a body  omega x (iota omega) = 0  and  v x (m v) = 0, so its bias force is (0, m omega x v) and it adds only three
// diagonal terms to I and M and nothing to H.  A URDF with other inertials takes PNR_CHAIN_PIONEER.
#define PNR_CHAIN_PIONEER_ISO 2

template <int CHAIN>
PNR_HD int pnr_code(const PnrParams& p, int j) {
    if (CHAIN != PNR_CHAIN_GENERIC) return j == 0 ? PNR_AXIS_Z : ((j == 3 || j == 5) ? PNR_AXIS_X : PNR_AXIS_Y);
    return p.axis_code[j];
}
template <int CHAIN>
PNR_HD V3 pnr_axis(const PnrParams& p, int j) {
    if (CHAIN != PNR_CHAIN_GENERIC) {
        const int c = pnr_code<CHAIN>(p, j);
        return v3(c == PNR_AXIS_X ? 1.f : 0.f, c == PNR_AXIS_Y ? 1.f : 0.f, c == PNR_AXIS_Z ? 1.f : 0.f);
    }
    return v3(p.axis[j][0], p.axis[j][1], p.axis[j][2]);
}
// axis . v,  axis * s,  v x (axis * s),  I axis (a column),  H^T axis (a row): single components when the axis is known
template <int CHAIN>
PNR_HD float pnr_axis_dot(const PnrParams& p, int j, V3 v) {
    if (CHAIN != PNR_CHAIN_GENERIC) {
        const int c = pnr_code<CHAIN>(p, j);
        return c == PNR_AXIS_X ? v.x : (c == PNR_AXIS_Y ? v.y : v.z);
    }
    return dot(pnr_axis<CHAIN>(p, j), v);
}
template <int CHAIN>
PNR_HD V3 pnr_axis_scaled(const PnrParams& p, int j, float s) {
    if (CHAIN != PNR_CHAIN_GENERIC) {
        const int c = pnr_code<CHAIN>(p, j);
        return v3(c == PNR_AXIS_X ? s : 0.f, c == PNR_AXIS_Y ? s : 0.f, c == PNR_AXIS_Z ? s : 0.f);
    }
    return pnr_axis<CHAIN>(p, j) * s;
}
template <int CHAIN>
PNR_HD V3 pnr_cross_axis(const PnrParams& p, int j, V3 a, float s) {     // a x (axis * s)
    if (CHAIN != PNR_CHAIN_GENERIC) {
        const int c = pnr_code<CHAIN>(p, j);
        if (c == PNR_AXIS_X) return v3(0.f, a.z * s, -a.y * s);
        if (c == PNR_AXIS_Y) return v3(-a.z * s, 0.f, a.x * s);
        return v3(a.y * s, -a.x * s, 0.f);
    }
    return cross(a, pnr_axis<CHAIN>(p, j) * s);
}
template <int CHAIN>
PNR_HD V3 pnr_sym_axis(const PnrParams& p, int j, const Sym3& m) {        // m * axis
    if (CHAIN != PNR_CHAIN_GENERIC) {
        const int c = pnr_code<CHAIN>(p, j);
        if (c == PNR_AXIS_X) return v3(m.xx, m.xy, m.xz);
        if (c == PNR_AXIS_Y) return v3(m.xy, m.yy, m.yz);
        return v3(m.xz, m.yz, m.zz);
    }
    return symmul(m, pnr_axis<CHAIN>(p, j));
}
template <int CHAIN>
PNR_HD V3 pnr_matT_axis(const PnrParams& p, int j, const Mat3& a) {       // a^T * axis
    if (CHAIN != PNR_CHAIN_GENERIC) {
        const int c = pnr_code<CHAIN>(p, j);
        return v3(a.m[3 * c + 0], a.m[3 * c + 1], a.m[3 * c + 2]);
    }
    return matTmul(a, pnr_axis<CHAIN>(p, j));
}

// origin_j x v.  The shipped robot's joint origins lie on one coordinate axis of the parent frame (joint 0 and 5: zero,
// 1 and 2: along z, 3: along y, 4: along x; pioneer_knm_6dof.urdf:209-264), so the cross product is two multiplies.
template <int CHAIN>
PNR_HD V3 pnr_origin_cross(const PnrParams& p, int j, V3 v) {
    if (CHAIN != PNR_CHAIN_GENERIC) {
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
PNR_HD V3 pnr_axis_rot(const PnrParams& p, int j, float s, float c, V3 v) {
    const int code = pnr_code<CHAIN>(p, j);
    if (CHAIN == PNR_CHAIN_GENERIC) s *= p.axis_sign[j];
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
PNR_HD V3 pnr_rot(const PnrParams& p, int j, float s, float c, V3 v) {
    V3 r = pnr_axis_rot<CHAIN>(p, j, s, c, v);
    if (CHAIN == PNR_CHAIN_GENERIC && p.origin_has_rot[j]) {
        const float* O = p.origin_rot[j];
        r = v3(O[0] * r.x + O[1] * r.y + O[2] * r.z, O[3] * r.x + O[4] * r.y + O[5] * r.z, O[6] * r.x + O[7] * r.y + O[8] * r.z);
    }
    return r;
}

// parent -> child coordinates: R_j^T v
template <int CHAIN>
PNR_HD V3 pnr_rot_t(const PnrParams& p, int j, float s, float c, V3 v) {
    if (CHAIN == PNR_CHAIN_GENERIC && p.origin_has_rot[j]) {
        const float* O = p.origin_rot[j];
        v = v3(O[0] * v.x + O[3] * v.y + O[6] * v.z, O[1] * v.x + O[4] * v.y + O[7] * v.z, O[2] * v.x + O[5] * v.y + O[8] * v.z);
    }
    return pnr_axis_rot<CHAIN>(p, j, -s, c, v);
}

// B' = R B R^T for a general 3x3 block: rotate the columns, then the rows
template <int CHAIN>
PNR_HD Mat3 pnr_rot_block(const PnrParams& p, int j, float s, float c, const Mat3& b) {
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

PNR_HD Mat3 sym_to_mat(const Sym3& s) {
    Mat3 o = {{s.xx, s.xy, s.xz, s.xy, s.yy, s.yz, s.xz, s.yz, s.zz}};
    return o;
}
PNR_HD Sym3 mat_to_sym(const Mat3& a) {     // a is symmetric up to rounding: average the pairs
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

// qdd = ABA(q, qd, tau); gravity acts along -z of the base frame.  Any serial 6-revolute chain (CHAIN = GENERIC reads the
// axis codes at run time); with CHAIN = PIONEER this is the first specialisation (kept as the A/B baseline of
// pnr_aba_pioneer below, tests/csrc/aba_check.cu).
// LINKDAMP adds Bullet's per-link damping as an external force on every body (PNR_STEPPING_BULLET): at the centre of mass,
// force -m v_com (k + k |v_com|) and torque -(I_com w) (k + k |w|), k = dyn_link_damping.
template <int CHAIN, bool LINKDAMP>
PNR_HD void pnr_aba_general_ext(const PnrParams& p, const float (&q)[PNR_DOF], const float (&qd)[PNR_DOF],
                                const float (&tau)[PNR_DOF], float gravity, float (&qdd)[PNR_DOF]) {
    PnrDynWork w;
    // ---- pass 1 (base -> tip): velocities, velocity-product accelerations, bias forces
    V3 om = v3(0.f, 0.f, 0.f), vl = v3(0.f, 0.f, 0.f);
#pragma unroll (CHAIN == 0 ? 1 : 6)   // GENERIC: rolled joint loops (the unrolled body does not fit the instruction cache)
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
        if (LINKDAMP) {
            const float kd = p.dyn_link_damping;
            const V3 com = v3(p.dyn_com[i][0], p.dyn_com[i][1], p.dyn_com[i][2]);
            const V3 vc = vl + cross(om, com);
            const float lin = fmaf(kd, sqrtf(dot(vc, vc)), kd), ang = fmaf(kd, sqrtf(dot(om, om)), kd);
            const Sym3 Ic = {p.dyn_icom[i][0], p.dyn_icom[i][1], p.dyn_icom[i][2], p.dyn_icom[i][3], p.dyn_icom[i][4], p.dyn_icom[i][5]};
            const V3 fe = vc * (-lin * p.dyn_mass[i]);
            const V3 ne = symmul(Ic, om) * (-ang) + cross(com, fe);
            w.p_ang[i] = w.p_ang[i] - ne;
            w.p_lin[i] = w.p_lin[i] - fe;
        }
    }
    // ---- pass 2 (tip -> base): articulated inertias and bias forces
    Sym3 I, M;
    Mat3 H;
    V3 pa_ang, pa_lin;
#pragma unroll (CHAIN == 0 ? 1 : 6)   // GENERIC: rolled joint loops (the unrolled body does not fit the instruction cache)
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
    V3 aa = v3(0.f, 0.f, 0.f), al = v3(0.f, 0.f, gravity);
#pragma unroll (CHAIN == 0 ? 1 : 6)   // GENERIC: rolled joint loops (the unrolled body does not fit the instruction cache)
    for (int i = 0; i < PNR_DOF; ++i) {
        const V3 t = al - pnr_origin_cross<CHAIN>(p, i, aa);
        aa = pnr_rot_t<CHAIN>(p, i, w.sn[i], w.cs[i], aa) + w.c_ang[i];
        al = pnr_rot_t<CHAIN>(p, i, w.sn[i], w.cs[i], t) + w.c_lin[i];
        qdd[i] = (w.u[i] - dot(w.u_ang[i], aa) - dot(w.u_lin[i], al)) * w.dinv[i];
        aa = aa + pnr_axis_scaled<CHAIN>(p, i, qdd[i]);
    }
}

template <int CHAIN>
PNR_HD void pnr_aba_general(const PnrParams& p, const float (&q)[PNR_DOF], const float (&qd)[PNR_DOF],
                            const float (&tau)[PNR_DOF], float (&qdd)[PNR_DOF]) {
    pnr_aba_general_ext<CHAIN, false>(p, q, qd, tau, p.dyn_gravity, qdd);
}

// ---------------------------------------------------------------------------------------------------------------
// pnr_aba_pioneer: the same algorithm written for the structure of the shipped robot (PNR_CHAIN_PIONEER: joint axes
// Z Y Y X Y X, positive, no origin rotations, every joint origin on ONE coordinate axis of its parent), with the
// sparsity that structure implies spelled out instead of left to the compiler (which may not fold x * 0 or x + 0):
//   * a rotation about coordinate axis k touches only the components a = k+1, b = k+2 (cyclic): vectors cost 4
//     operations, a SYMMETRIC block R S R^T is evaluated in closed form from cos^2, sin^2, sin cos (12 operations
//     instead of 24 + 6 to re-symmetrise);
//   * moving the reference point by p = p_k e_k:  H'' = H' + [p]x M',  I'' = I' + (H' [p]x^T)^T + H'' [p]x^T  has six
//     distinct entries of I and six of H that change -- 14 FMAs instead of three 3x3 products and a re-symmetrisation;
//   * the velocity-product terms c = v x (e_k qd) have a zero k-component: the 3x3 products with them skip a column;
//   * the tip joint's articulated inertia after its own rank-1 update is a constant of the robot: precomputed on the
//     host in float64 (PnrParams::dyn_tip_*, pnr_build_params);
//   * joint 0 sits on the fixed base: its velocity is e_k qd and its acceleration bias is zero.
// ---------------------------------------------------------------------------------------------------------------
// 1 / x for a well-scaled positive x: the hardware reciprocal (MUFU.RCP, <= 1 ulp) without the IEEE fix-up sequence and
// its slow-path call (5 per substep); the host build (tests/csrc/aba_check.cu) divides
PNR_HD float pnr_rcp_fast(float x) {
#ifdef __CUDA_ARCH__
    return __fdividef(1.f, x);
#else
    return 1.f / x;
#endif
}

PNR_HD int pnr_pio_code(int i) { return i == 0 ? PNR_AXIS_Z : ((i == 3 || i == 5) ? PNR_AXIS_X : PNR_AXIS_Y); }
PNR_HD int pnr_pio_ocode(int i) { return (i == 0 || i == 5) ? -1 : (i == 3 ? 1 : (i == 4 ? 0 : 2)); }   // axis the origin lies on
PNR_HD int pnr_nxt(int k) { return k == 2 ? 0 : k + 1; }       // a: the component after k, cyclic
PNR_HD int pnr_prv(int k) { return k == 0 ? 2 : k - 1; }       // b: the component before k

PNR_HD float vget(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
PNR_HD void vset(V3& v, int i, float f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else v.z = f; }
PNR_HD float sget(const Sym3& s, int i, int j) {
    const int lo = i < j ? i : j, hi = i < j ? j : i, k = 3 * lo + hi;
    return k == 0 ? s.xx : (k == 1 ? s.xy : (k == 2 ? s.xz : (k == 4 ? s.yy : (k == 5 ? s.yz : s.zz))));
}
PNR_HD void sset(Sym3& s, int i, int j, float f) {
    const int lo = i < j ? i : j, hi = i < j ? j : i, k = 3 * lo + hi;
    if (k == 0) s.xx = f; else if (k == 1) s.xy = f; else if (k == 2) s.xz = f; else if (k == 4) s.yy = f;
    else if (k == 5) s.yz = f; else s.zz = f;
}

// Rot(e_k, angle) v (child -> parent coordinates); the transpose takes -s
PNR_HD V3 pnr_rotk(int k, float s, float c, V3 v) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float va = vget(v, a), vb = vget(v, b);
    vset(v, a, fmaf(c, va, -s * vb));
    vset(v, b, fmaf(s, va, c * vb));
    return v;
}
// e_k s + v,  v x (e_k s),  (e_k s) x v
PNR_HD V3 pnr_add_ek(int k, V3 v, float s) { vset(v, k, vget(v, k) + s); return v; }
PNR_HD V3 pnr_cross_ek(int k, V3 v, float s) {                 // k-component is an exact zero
    const int a = pnr_nxt(k), b = pnr_prv(k);
    V3 r = v3(0.f, 0.f, 0.f);
    vset(r, a, vget(v, b) * s);
    vset(r, b, -vget(v, a) * s);
    return r;
}
PNR_HD V3 pnr_ek_cross(int k, float s, V3 v) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    V3 r = v3(0.f, 0.f, 0.f);
    vset(r, a, -s * vget(v, b));
    vset(r, b, s * vget(v, a));
    return r;
}
// v - (e_k pk) x w   and   v + (e_k pk) x w: only two components change
PNR_HD V3 pnr_sub_origin_cross(int k, float pk, V3 v, V3 w) {
    if (k < 0) return v;
    const int a = pnr_nxt(k), b = pnr_prv(k);
    vset(v, a, fmaf(pk, vget(w, b), vget(v, a)));
    vset(v, b, fmaf(-pk, vget(w, a), vget(v, b)));
    return v;
}
PNR_HD V3 pnr_add_origin_cross(int k, float pk, V3 v, V3 w) { return pnr_sub_origin_cross(k, -pk, v, w); }
// S v, A v, A^T v for a v whose k-component is an exact zero
PNR_HD V3 pnr_symmul_z(int k, const Sym3& s, V3 v) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float va = vget(v, a), vb = vget(v, b);
    return v3(fmaf(sget(s, 0, a), va, sget(s, 0, b) * vb), fmaf(sget(s, 1, a), va, sget(s, 1, b) * vb),
              fmaf(sget(s, 2, a), va, sget(s, 2, b) * vb));
}
PNR_HD V3 pnr_matmul_z(int k, const Mat3& m, V3 v) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float va = vget(v, a), vb = vget(v, b);
    return v3(fmaf(m.m[a], va, m.m[b] * vb), fmaf(m.m[3 + a], va, m.m[3 + b] * vb), fmaf(m.m[6 + a], va, m.m[6 + b] * vb));
}
PNR_HD V3 pnr_matTmul_z(int k, const Mat3& m, V3 v) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float va = vget(v, a), vb = vget(v, b);
    return v3(fmaf(m.m[3 * a], va, m.m[3 * b] * vb), fmaf(m.m[3 * a + 1], va, m.m[3 * b + 1] * vb),
              fmaf(m.m[3 * a + 2], va, m.m[3 * b + 2] * vb));
}
struct PnrRot2 { float s, c, cc, ss, sc, sc2, cms; };           // sin, cos and the products the closed forms need
PNR_HD PnrRot2 pnr_rot2(float s, float c) {
    PnrRot2 r;
    r.s = s; r.c = c; r.cc = c * c; r.ss = s * s; r.sc = s * c; r.sc2 = r.sc + r.sc; r.cms = r.cc - r.ss;
    return r;
}
// R S R^T for a symmetric block and R = Rot(e_k): closed form; the trace of the (a, b) sub-block is preserved
PNR_HD Sym3 pnr_rotk_sym(int k, const PnrRot2& r, const Sym3& S) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float saa = sget(S, a, a), sab = sget(S, a, b), sbb = sget(S, b, b), sak = sget(S, a, k), sbk = sget(S, b, k);
    Sym3 o = S;
    const float naa = fmaf(r.cc, saa, fmaf(-r.sc2, sab, r.ss * sbb));
    sset(o, a, a, naa);
    sset(o, b, b, (saa + sbb) - naa);
    sset(o, a, b, fmaf(r.sc, saa - sbb, r.cms * sab));
    sset(o, a, k, fmaf(r.c, sak, -r.s * sbk));
    sset(o, b, k, fmaf(r.s, sak, r.c * sbk));
    return o;
}
// R B R^T for a general block: rows a, b mix, then columns a, b
PNR_HD Mat3 pnr_rotk_block(int k, const PnrRot2& r, const Mat3& B) {
    const int a = pnr_nxt(k), b = pnr_prv(k);
    Mat3 t = B;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        t.m[3 * a + j] = fmaf(r.c, B.m[3 * a + j], -r.s * B.m[3 * b + j]);
        t.m[3 * b + j] = fmaf(r.s, B.m[3 * a + j], r.c * B.m[3 * b + j]);
    }
    Mat3 o = t;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o.m[3 * i + a] = fmaf(r.c, t.m[3 * i + a], -r.s * t.m[3 * i + b]);
        o.m[3 * i + b] = fmaf(r.s, t.m[3 * i + a], r.c * t.m[3 * i + b]);
    }
    return o;
}
// After joint k's rank-1 update the articulated inertia annihilates the joint axis: column / row k of I^a and row k of
// H^a are exact zeros (I^a S = 0).  They are not computed; the variants below skip them.
PNR_HD Sym3 pnr_rotk_sym_z(int k, const PnrRot2& r, const Sym3& S) {       // only the (a, b) block is populated
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float saa = sget(S, a, a), sab = sget(S, a, b), sbb = sget(S, b, b);
    Sym3 o = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float naa = fmaf(r.cc, saa, fmaf(-r.sc2, sab, r.ss * sbb));
    sset(o, a, a, naa);
    sset(o, b, b, (saa + sbb) - naa);
    sset(o, a, b, fmaf(r.sc, saa - sbb, r.cms * sab));
    return o;
}
PNR_HD Mat3 pnr_rotk_block_z(int k, const PnrRot2& r, const Mat3& B) {     // row k of B (and of the result) is zero
    const int a = pnr_nxt(k), b = pnr_prv(k);
    Mat3 t = {{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        t.m[3 * a + j] = fmaf(r.c, B.m[3 * a + j], -r.s * B.m[3 * b + j]);
        t.m[3 * b + j] = fmaf(r.s, B.m[3 * a + j], r.c * B.m[3 * b + j]);
    }
    Mat3 o = t;
    o.m[3 * a + a] = fmaf(r.c, t.m[3 * a + a], -r.s * t.m[3 * a + b]);
    o.m[3 * a + b] = fmaf(r.s, t.m[3 * a + a], r.c * t.m[3 * a + b]);
    o.m[3 * b + a] = fmaf(r.c, t.m[3 * b + a], -r.s * t.m[3 * b + b]);
    o.m[3 * b + b] = fmaf(r.s, t.m[3 * b + a], r.c * t.m[3 * b + b]);
    return o;
}
// move the reference point of the (rotated) articulated inertia by p = e_k pk (see the header comment); M is unchanged
PNR_HD void pnr_shift_ek(int k, float pk, Sym3& I, Mat3& H, const Sym3& M) {
    if (k < 0) return;
    const int a = pnr_nxt(k), b = pnr_prv(k);
    const float h_ab = H.m[3 * a + b], h_bb = H.m[3 * b + b], h_kb = H.m[3 * k + b];
    const float h_ba = H.m[3 * b + a], h_ka = H.m[3 * k + a];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        H.m[3 * a + j] = fmaf(-pk, sget(M, b, j), H.m[3 * a + j]);
        H.m[3 * b + j] = fmaf(pk, sget(M, a, j), H.m[3 * b + j]);
    }
    sset(I, a, a, fmaf(-pk, h_ab + H.m[3 * a + b], sget(I, a, a)));
    sset(I, a, b, fmaf(pk, H.m[3 * a + a], fmaf(-pk, h_bb, sget(I, a, b))));
    sset(I, a, k, fmaf(-pk, h_kb, sget(I, a, k)));
    sset(I, b, b, fmaf(pk, h_ba + H.m[3 * b + a], sget(I, b, b)));
    sset(I, b, k, fmaf(pk, h_ka, sget(I, b, k)));
}

template <bool ISO>
PNR_HD void pnr_aba_pioneer(const PnrParams& p, const float (&q)[PNR_DOF], const float (&qd)[PNR_DOF],
                            const float (&tau)[PNR_DOF], float (&qdd)[PNR_DOF]) {
    PnrDynWork w;
    // ---- pass 1 (base -> tip): velocities, velocity-product accelerations, bias forces of the rigid bodies
    V3 om = v3(0.f, 0.f, 0.f), vl = v3(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) {
        const int k = pnr_pio_code(i), ok = pnr_pio_ocode(i), a = pnr_nxt(k), b = pnr_prv(k);
        pnr_sincos_bounded(q[i], w.sn[i], w.cs[i]);                    // chain_kind 1 implies joint limits within +-64 rad
        const Sym3 Io = {p.dyn_io[i][0], p.dyn_io[i][1], p.dyn_io[i][2], p.dyn_io[i][3], p.dyn_io[i][4], p.dyn_io[i][5]};
        const V3 mc = v3(p.dyn_mc[i][0], p.dyn_mc[i][1], p.dyn_mc[i][2]);
        const bool iso = ISO && i < PNR_DOF - 1;                       // bias force (0, m om x vl)
        if (i == 0) {
            // fixed base: om = e_k qd, vl = 0, c = 0;  n = Io om,  f = -mc x om,  p = (om x n, om x f)
            const float wq = qd[0];
            if (!iso) {
                V3 n = v3(sget(Io, 0, k) * wq, sget(Io, 1, k) * wq, sget(Io, 2, k) * wq);
                V3 f = v3(0.f, 0.f, 0.f);
                vset(f, a, -vget(mc, b) * wq);
                vset(f, b, vget(mc, a) * wq);
                w.p_ang[0] = pnr_ek_cross(k, wq, n);
                w.p_lin[0] = pnr_ek_cross(k, wq, f);
            }
            om = v3(0.f, 0.f, 0.f);
            vset(om, k, wq);
        } else {
            const V3 t = pnr_sub_origin_cross(ok, ok >= 0 ? p.origin_xyz[i][ok] : 0.f, vl, om);
            om = pnr_add_ek(k, pnr_rotk(k, -w.sn[i], w.cs[i], om), qd[i]);
            vl = pnr_rotk(k, -w.sn[i], w.cs[i], t);
            w.c_ang[i] = pnr_cross_ek(k, om, qd[i]);
            w.c_lin[i] = pnr_cross_ek(k, vl, qd[i]);
            if (iso) {
                w.p_lin[i] = cross(om, vl);                             // the mass is applied where it is accumulated (pass 2)
            } else {
                const V3 n = symmul(Io, om) + cross(mc, vl);
                const V3 f = vl * p.dyn_mass[i] - cross(mc, om);
                w.p_ang[i] = cross(om, n) + cross(vl, f);
                w.p_lin[i] = cross(om, f);
            }
        }
    }
    // ---- pass 2 (tip -> base): articulated inertias and bias forces
    Sym3 I, M;
    Mat3 H;
    V3 pa_ang = w.p_ang[PNR_DOF - 1], pa_lin = w.p_lin[PNR_DOF - 1];
#pragma unroll
    for (int i = PNR_DOF - 1; i >= 0; --i) {
        const int k = pnr_pio_code(i), ok = pnr_pio_ocode(i);
        V3 Ua, Ul;
        float dinv;
        if (i == PNR_DOF - 1) {
            // the tip body's inertia after its own rank-1 update never changes: host-computed constants
            Sym3 i0 = {p.dyn_tip_I[0], p.dyn_tip_I[1], p.dyn_tip_I[2], p.dyn_tip_I[3], p.dyn_tip_I[4], p.dyn_tip_I[5]};
            Sym3 m0 = {p.dyn_tip_M[0], p.dyn_tip_M[1], p.dyn_tip_M[2], p.dyn_tip_M[3], p.dyn_tip_M[4], p.dyn_tip_M[5]};
            I = i0; M = m0;
#pragma unroll
            for (int e = 0; e < 9; ++e) H.m[e] = p.dyn_tip_H[e];
            Ua = v3(p.dyn_tip_ua[0], p.dyn_tip_ua[1], p.dyn_tip_ua[2]);
            Ul = v3(p.dyn_tip_ul[0], p.dyn_tip_ul[1], p.dyn_tip_ul[2]);
            dinv = p.dyn_tip_dinv;
        } else {                                                        // own rigid body + what the child handed up
            const float m = p.dyn_mass[i];
            if (ISO) {                                                  // centre of mass on the origin, inertia iota * 1
                const float iota = p.dyn_io[i][0];
                I.xx += iota; I.yy += iota; I.zz += iota;
                if (i > 0) {                                            // + m om x vl; the base joint's own bias force is zero
                    pa_lin = v3(fmaf(m, w.p_lin[i].x, pa_lin.x), fmaf(m, w.p_lin[i].y, pa_lin.y), fmaf(m, w.p_lin[i].z, pa_lin.z));
                }
            } else {
                const V3 mc = v3(p.dyn_mc[i][0], p.dyn_mc[i][1], p.dyn_mc[i][2]);
                I.xx += p.dyn_io[i][0]; I.xy += p.dyn_io[i][1]; I.xz += p.dyn_io[i][2];
                I.yy += p.dyn_io[i][3]; I.yz += p.dyn_io[i][4]; I.zz += p.dyn_io[i][5];
                H.m[1] -= mc.z; H.m[2] += mc.y; H.m[3] += mc.z; H.m[5] -= mc.x; H.m[6] -= mc.y; H.m[7] += mc.x;
                pa_ang = pa_ang + w.p_ang[i];
                pa_lin = pa_lin + w.p_lin[i];
            }
            M.xx += m; M.yy += m; M.zz += m;
            Ua = v3(sget(I, 0, k), sget(I, 1, k), sget(I, 2, k));       // I e_k
            Ul = v3(H.m[3 * k], H.m[3 * k + 1], H.m[3 * k + 2]);        // H^T e_k
            dinv = pnr_rcp_fast(vget(Ua, k));                           // d = S^T I^A S >= the body's own inertia about the axis
        }
        const float u = tau[i] - vget(pa_ang, k);
        w.u_ang[i] = Ua; w.u_lin[i] = Ul; w.dinv[i] = dinv; w.u[i] = u;
        if (i > 0) {
            const int a = pnr_nxt(k), b = pnr_prv(k);
            if (i != PNR_DOF - 1) {                                     // I^a = I^A - U U^T / d; entries with index k vanish
                const float uaa = vget(Ua, a), uab = vget(Ua, b);
                const float da = uaa * dinv, db = uab * dinv;
                const V3 Uld = Ul * dinv;
                const float iaa = fmaf(-uaa, da, sget(I, a, a)), iab = fmaf(-uaa, db, sget(I, a, b));
                const float ibb = fmaf(-uab, db, sget(I, b, b));
                Sym3 z = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                I = z;
                sset(I, a, a, iaa); sset(I, a, b, iab); sset(I, b, b, ibb);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    H.m[3 * a + j] = fmaf(-da, vget(Ul, j), H.m[3 * a + j]);
                    H.m[3 * b + j] = fmaf(-db, vget(Ul, j), H.m[3 * b + j]);
                    H.m[3 * k + j] = 0.f;
                }
                M.xx -= Ul.x * Uld.x; M.xy -= Ul.x * Uld.y; M.xz -= Ul.x * Uld.z;
                M.yy -= Ul.y * Uld.y; M.yz -= Ul.y * Uld.z; M.zz -= Ul.z * Uld.z;
            }
            // p^a = p^A + I^a c + U u / d   (c has a zero k-component; so have I^a c_ang and H^a c_lin)
            const float ud = u * dinv;
            const float ca_a = vget(w.c_ang[i], a), ca_b = vget(w.c_ang[i], b);
            const float cl_a = vget(w.c_lin[i], a), cl_b = vget(w.c_lin[i], b);
            // every term is folded into one FMA chain per component that ends in the accumulator (no separate adds)
            vset(pa_ang, a, fmaf(sget(I, a, a), ca_a, fmaf(sget(I, a, b), ca_b, fmaf(H.m[3 * a + a], cl_a,
                            fmaf(H.m[3 * a + b], cl_b, fmaf(vget(Ua, a), ud, vget(pa_ang, a)))))));
            vset(pa_ang, b, fmaf(sget(I, a, b), ca_a, fmaf(sget(I, b, b), ca_b, fmaf(H.m[3 * b + a], cl_a,
                            fmaf(H.m[3 * b + b], cl_b, fmaf(vget(Ua, b), ud, vget(pa_ang, b)))))));
            vset(pa_ang, k, fmaf(vget(Ua, k), ud, vget(pa_ang, k)));
#pragma unroll
            for (int j = 0; j < 3; ++j)                                 // H^T c_ang + M c_lin + Ul u / d
                vset(pa_lin, j, fmaf(H.m[3 * a + j], ca_a, fmaf(H.m[3 * b + j], ca_b, fmaf(sget(M, j, a), cl_a,
                                fmaf(sget(M, j, b), cl_b, fmaf(vget(Ul, j), ud, vget(pa_lin, j)))))));
            // hand I^a, p^a up to the parent: rotate by Rot(e_k, q_i), then move the reference point by p_i
            const PnrRot2 r = pnr_rot2(w.sn[i], w.cs[i]);
            const float pk = ok >= 0 ? p.origin_xyz[i][ok] : 0.f;
            I = pnr_rotk_sym_z(k, r, I);
            M = pnr_rotk_sym(k, r, M);
            H = pnr_rotk_block_z(k, r, H);
            pnr_shift_ek(ok, pk, I, H, M);
            const V3 fp = pnr_rotk(k, r.s, r.c, pa_lin);
            pa_ang = pnr_add_origin_cross(ok, pk, pnr_rotk(k, r.s, r.c, pa_ang), fp);
            pa_lin = fp;
        }
    }
    // ---- pass 3 (base -> tip): accelerations.  The base "accelerates" by -g: a_lin = (0, 0, +gravity)
    V3 aa, al;
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) {
        const int k = pnr_pio_code(i), ok = pnr_pio_ocode(i);
        if (i == 0) {                                                   // joint 0 turns about z: gravity stays (0, 0, g)
            al = v3(0.f, 0.f, p.dyn_gravity);
            qdd[0] = (w.u[0] - w.u_lin[0].z * p.dyn_gravity) * w.dinv[0];
            aa = v3(0.f, 0.f, qdd[0]);
        } else {
            const V3 t = pnr_sub_origin_cross(ok, ok >= 0 ? p.origin_xyz[i][ok] : 0.f, al, aa);
            aa = pnr_rotk(k, -w.sn[i], w.cs[i], aa);
            al = pnr_rotk(k, -w.sn[i], w.cs[i], t);
            const int a = pnr_nxt(k), b = pnr_prv(k);                   // c_ang, c_lin have a zero k-component
            vset(aa, a, vget(aa, a) + vget(w.c_ang[i], a)); vset(aa, b, vget(aa, b) + vget(w.c_ang[i], b));
            vset(al, a, vget(al, a) + vget(w.c_lin[i], a)); vset(al, b, vget(al, b) + vget(w.c_lin[i], b));
            float acc = w.u[i];                                         // u - U . a as one FMA chain
            acc = fmaf(-w.u_ang[i].x, aa.x, acc); acc = fmaf(-w.u_ang[i].y, aa.y, acc); acc = fmaf(-w.u_ang[i].z, aa.z, acc);
            acc = fmaf(-w.u_lin[i].x, al.x, acc); acc = fmaf(-w.u_lin[i].y, al.y, acc); acc = fmaf(-w.u_lin[i].z, al.z, acc);
            qdd[i] = acc * w.dinv[i];
            aa = pnr_add_ek(k, aa, qdd[i]);
        }
    }
}

template <int CHAIN>
PNR_HD void pnr_aba(const PnrParams& p, const float (&q)[PNR_DOF], const float (&qd)[PNR_DOF],
                    const float (&tau)[PNR_DOF], float (&qdd)[PNR_DOF]) {
#ifdef PNR_ABA_BASELINE                                 // A/B builds: the first specialisation
    pnr_aba_general<CHAIN>(p, q, qd, tau, qdd);
#else
    if (CHAIN == PNR_CHAIN_PIONEER_ISO) pnr_aba_pioneer<true>(p, q, qd, tau, qdd);
    else if (CHAIN != PNR_CHAIN_GENERIC) pnr_aba_pioneer<false>(p, q, qd, tau, qdd);
    else pnr_aba_general<CHAIN>(p, q, qd, tau, qdd);
#endif
}

// joint PD / torque control with effort clamping, then viscous joint damping (oracle/dynamics_oracle.py::control_torque)
PNR_HD float pnr_control_torque(const PnrParams& p, int i, float action, float q, float qd) {
    float tau = p.dyn_use_pd ? fmaf(p.dyn_kp, action - q, -p.dyn_kd * qd) : action;
    const float lim = p.dyn_tau_max[i];
    tau = fminf(fmaxf(tau, -lim), lim);
    return fmaf(-p.dyn_damping[i], qd, tau);
}

// inelastic stops at the joint limits
PNR_HD void pnr_joint_stop(const PnrParams& p, int i, float& x, float& v) {
    if (x > p.r_hi[i]) { x = p.r_hi[i]; if (v > 0.f) v = 0.f; }
    if (x < p.r_lo[i]) { x = p.r_lo[i]; if (v < 0.f) v = 0.f; }
}

// PNR_STEPPING_BULLET: one Bullet-like substep, float32 twin of oracle/dynamics_oracle.c::dyn_substep_bullet (restated from
// memory of btMultiBody / btMultiBodyJointMotor, unpinned): per-link damping + URDF joint damping -> unconstrained velocity;
// POSITION_CONTROL motors as velocity-level constraints (target positionGain (u - q) / dt + (1 - velocityGain) v*, impulse
// limit force * dt) solved by PNR_BULLET_ITERATIONS sweeps of projected Gauss-Seidel on A = M(q)^-1, whose columns are the
// responses to unit joint impulses (one ABA call each: qd = 0, g = 0, tau = e_j); +-max_velocity clamp; stops.
// Opt-in and seven ABA evaluations per substep: correctness first, the joint loop over the unit impulses is rolled.
#define PNR_BULLET_ITERATIONS 10
template <int CHAIN>
PNR_HD void pnr_bullet_substep(const PnrParams& p, float (&q)[PNR_DOF], float (&qd)[PNR_DOF], const float (&action)[PNR_DOF]) {
    float tau[PNR_DOF], qdd[PNR_DOF], v[PNR_DOF];
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) tau[i] = -p.dyn_damping[i] * qd[i];
    pnr_aba_general_ext<CHAIN, true>(p, q, qd, tau, p.dyn_gravity, qdd);
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) v[i] = fmaf(qdd[i], p.dyn_dt, qd[i]);
    if (p.dyn_motor_impulse > 0.f) {
        float A[PNR_DOF][PNR_DOF], rhs[PNR_DOF], lam[PNR_DOF];
        const float zero[PNR_DOF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int j = 0; j < PNR_DOF; ++j) {
            float e[PNR_DOF], col[PNR_DOF];
#pragma unroll
            for (int i = 0; i < PNR_DOF; ++i) e[i] = i == j ? 1.f : 0.f;
            pnr_aba_general_ext<CHAIN, false>(p, q, zero, e, 0.f, col);
#pragma unroll
            for (int i = 0; i < PNR_DOF; ++i) A[i][j] = col[i];
            rhs[j] = fmaf(p.dyn_motor_kp / p.dyn_dt, action[j] - q[j], (1.f - p.dyn_motor_kd) * v[j]);
            lam[j] = 0.f;
        }
        const float max_imp = p.dyn_motor_impulse;
#pragma unroll 1
        for (int it = 0; it < PNR_BULLET_ITERATIONS; ++it)
#pragma unroll 1
            for (int j = 0; j < PNR_DOF; ++j) {
                float dl = (rhs[j] - v[j]) / A[j][j];
                const float nl = fminf(fmaxf(lam[j] + dl, -max_imp), max_imp);
                dl = nl - lam[j];
                lam[j] = nl;
#pragma unroll
                for (int i = 0; i < PNR_DOF; ++i) v[i] = fmaf(A[i][j], dl, v[i]);
            }
    }
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) {
        float vi = fminf(fmaxf(v[i], -p.dyn_max_velocity), p.dyn_max_velocity);
        float x = fmaf(vi, p.dyn_dt, q[i]);
        pnr_joint_stop(p, i, x, vi);
        q[i] = x; qd[i] = vi;
    }
}

// frame_skip substeps.  STEPPING = PNR_STEPPING_EXPLICIT: semi-implicit Euler, qd += qdd dt; q += qd dt, stops at the limits
template <int CHAIN, int STEPPING = PNR_STEPPING_EXPLICIT>
PNR_HD void pnr_dynamic_substeps(const PnrParams& p, float (&q)[PNR_DOF], float (&qd)[PNR_DOF],
                                                     const float (&action)[PNR_DOF]) {
#pragma unroll 1
    for (int sub = 0; sub < p.dyn_frame_skip; ++sub) {
        if (STEPPING == PNR_STEPPING_BULLET) {
            pnr_bullet_substep<CHAIN == PNR_CHAIN_PIONEER_ISO ? PNR_CHAIN_PIONEER : CHAIN>(p, q, qd, action);
        } else {
            float tau[PNR_DOF], qdd[PNR_DOF];
#pragma unroll
            for (int i = 0; i < PNR_DOF; ++i) tau[i] = pnr_control_torque(p, i, action[i], q[i], qd[i]);
            pnr_aba<CHAIN>(p, q, qd, tau, qdd);
#pragma unroll
            for (int i = 0; i < PNR_DOF; ++i) {
                float v = fmaf(qdd[i], p.dyn_dt, qd[i]);
                float x = fmaf(v, p.dyn_dt, q[i]);
                // (spelled out, not pnr_joint_stop: through the helper's references ptxas emits branches instead of selects)
                if (x > p.r_hi[i]) { x = p.r_hi[i]; if (v > 0.f) v = 0.f; }
                if (x < p.r_lo[i]) { x = p.r_lo[i]; if (v < 0.f) v = 0.f; }
                q[i] = x; qd[i] = v;
            }
        }
    }
}
