// Branch-free float32 sine/cosine pairs for the observation columns (observe(), pioneer_knm_env.py:194-211).
//
// 30 (sin, cos) pairs per env step dominate the FP32 work of the fused kernel.  The library sincosf()
// carries a data-dependent branch into a Payne-Hanek slow path after every range reduction, which keeps the
// compiler from interleaving independent evaluations; these versions are straight-line code, so the 24
// bounded pairs of one env are scheduled as one block of independent FMAs.
//
//   pnr_sincos_bounded : |x| <= 64   (joint angles, limit distances, joint rates: |x| <= 4 pi by construction)
//   pnr_sincos_fast    : |x| <= 105615 (stored actions inside any sane range); the caller falls back to
//                        sincosf() for larger / non-finite arguments, which are legal (actions are unclipped,
//                        pioneer_knm_env.py:144) but never produced by a policy bounded by the action space.
//
// Method: Cody-Waite reduction by pi/2 with the round-to-nearest "magic number" trick, then the classic
// degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4] (Cephes single precision, public domain
// coefficients).  Measured maximum error against float64 sin/cos over each range: < 1.2e-7 absolute
// (tests/test_trig_host.py compiles this header for the host and checks it).
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#ifdef __CUDA_ARCH__
#define PNR_F2I_BITS(x) __float_as_int(x)
#define PNR_I2F_BITS(x) __int_as_float(x)
#else
static inline int32_t pnr_host_f2i(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
static inline float pnr_host_i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
#define PNR_F2I_BITS(x) pnr_host_f2i(x)
#define PNR_I2F_BITS(x) pnr_host_i2f(x)
#endif

// t in [-pi/4, pi/4] (a little beyond is fine), q = quadrant count: x = q * pi/2 + t
__host__ __device__ __forceinline__ void pnr_sincos_quadrant(float t, int32_t q, float& s_out, float& c_out) {
    const float z = t * t;
    // sin t = t + t z (S1 + z (S2 + z S3))
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float sn = fmaf(t * z, ps, t);
    // cos t = 1 - z/2 + z^2 (C1 + z (C2 + z C3))
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    const float cs = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    // quadrant: q&1 swaps, bit 1 of q negates sin, bit 1 of (q+1) negates cos
    const bool swap = (q & 1) != 0;
    const float s_sel = swap ? cs : sn;
    const float c_sel = swap ? sn : cs;
    const int32_t s_sign = (q & 2) << 30;
    const int32_t c_sign = ((q + 1) & 2) << 30;
    s_out = PNR_I2F_BITS(PNR_F2I_BITS(s_sel) ^ s_sign);
    c_out = PNR_I2F_BITS(PNR_F2I_BITS(c_sel) ^ c_sign);
}

#define PNR_TRIG_MAGIC 12582912.0f          // 1.5 * 2^23: adding it rounds to the nearest integer in the low mantissa bits
#define PNR_TWO_OVER_PI 0.636619772367581343f
#define PNR_PIO2_HI 1.5707962512969971e+0f  // pi/2 split into three float32 pieces (Cody-Waite)
#define PNR_PIO2_MID 7.5497894158615964e-8f
#define PNR_PIO2_LO 5.3903029534742384e-15f

__host__ __device__ __forceinline__ void pnr_sincos_bounded(float x, float& s, float& c) {
    const float j = fmaf(x, PNR_TWO_OVER_PI, PNR_TRIG_MAGIC);
    const int32_t q = PNR_F2I_BITS(j);
    const float k = j - PNR_TRIG_MAGIC;
    float t = fmaf(k, -PNR_PIO2_HI, x);
    t = fmaf(k, -PNR_PIO2_MID, t);
    pnr_sincos_quadrant(t, q, s, c);
}

__host__ __device__ __forceinline__ void pnr_sincos_fast(float x, float& s, float& c) {
    const float j = fmaf(x, PNR_TWO_OVER_PI, PNR_TRIG_MAGIC);
    const int32_t q = PNR_F2I_BITS(j);
    const float k = j - PNR_TRIG_MAGIC;
    float t = fmaf(k, -PNR_PIO2_HI, x);
    t = fmaf(k, -PNR_PIO2_MID, t);
    t = fmaf(k, -PNR_PIO2_LO, t);
    pnr_sincos_quadrant(t, q, s, c);
}

#define PNR_TRIG_FAST_LIMIT 105615.0f
