// hdp_device.cuh — device-side building blocks shared by all kernels of
// libhdpde_b200.so (sm_100a).  Counter-based Philox draws generated in
// registers, reference-ordered Euler–Maruyama arithmetic (never contracted),
// per-coordinate Neumann reflection, and the enumerated user-function kinds.
//
// Reference arithmetic being reproduced (HighDimPDE.jl v2.6.0):
//   src/MLP.jl:415-417          x2 = x + (μ*dt + (σ*sqrt(dt))*ξ)
//   src/DeepSplitting.jl:159-160 y1 = (y0 + μ*dt) + (σ*sqrt(dt))*dW
//   src/reflect.jl:8-55          billiard reflection, per-coordinate equivalent
//   src/MCSample.jl:28-33,63-71  sampler affine maps
// The whole library is compiled with -fmad=false; fused multiply-adds appear
// only where written explicitly (inside the normal generator).
#pragma once
#ifdef HDP_HOST_EMU
// test-only hook: lets the tests compile this device code for the host; never defined in the product build
#include "hdp_host_emu_shim.h"
#define HDP_NOINLINE __attribute__((noinline))
#else
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define HDP_NOINLINE __noinline__
#endif
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif
#include "../../include/hdpde.h"

// Bodies of the *_USER kinds: CUDA-C expressions that the NVRTC build of the K3 kernels
// defines before including this file (hdp_user_exprs, include/hdpde.h).  In the static
// library they are inert.
#ifndef HDP_USER_MU_EXPR
#define HDP_USER_MU_EXPR (0)
#endif
#ifndef HDP_USER_SIGMA_EXPR
#define HDP_USER_SIGMA_EXPR (0)
#endif
#ifndef HDP_USER_G_EXPR
#define HDP_USER_G_EXPR (0)
#endif
#ifndef HDP_USER_F_EXPR
#define HDP_USER_F_EXPR (0)
#endif

namespace hdp {

typedef unsigned long long u64;

// ---------------------------------------------------------------------
// element-type traits: one Philox4x32 call = 128 random bits = VEC draws
// = one 128-bit load/store of path state.
// ---------------------------------------------------------------------
template <typename T> struct Tr;
template <> struct Tr<double> {
    static constexpr int VEC = 2;
    typedef double2 V;
};
template <> struct Tr<float> {
    static constexpr int VEC = 4;
    typedef float4 V;
};

template <typename T> struct Vec { T v[Tr<T>::VEC]; };

template <typename T> __device__ __forceinline__ Vec<T> ldv(const T* p)
{
    Vec<T> r;
    typename Tr<T>::V q = *reinterpret_cast<const typename Tr<T>::V*>(p);
    if constexpr (Tr<T>::VEC == 2) { r.v[0] = q.x; r.v[1] = q.y; }
    else { r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w; }
    return r;
}
template <typename T> __device__ __forceinline__ void stv(T* p, const Vec<T>& r)
{
    typename Tr<T>::V q;
    if constexpr (Tr<T>::VEC == 2) { q.x = r.v[0]; q.y = r.v[1]; }
    else { q.x = r.v[0]; q.y = r.v[1]; q.z = r.v[2]; q.w = r.v[3]; }
    *reinterpret_cast<typename Tr<T>::V*>(p) = q;
}

// ---------------------------------------------------------------------
// IEEE round-to-nearest primitives that the compiler may never contract
// ---------------------------------------------------------------------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }

// ---------------------------------------------------------------------
// Philox4x32-10.  ctr = (ordinal_lo, ordinal_hi, group, stream),
// key = (seed_lo, seed_hi).  Addressing by consumption ordinal — never by
// thread id — makes every draw independent of launch geometry and GPU count.
// ---------------------------------------------------------------------
struct Philox4 { uint32_t w[4]; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#ifndef HDP_HOST_EMU
    // keep the round keys as immediates added to the two seed registers: without this the
    // compiler hoists the 18 bumped keys out of the caller's coordinate loop and spills them
    asm volatile("" : "+r"(k0), "+r"(k1));
#endif
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// ---------------------------------------------------------------------
// bit-level helpers (host emulation provides the same functions)
// ---------------------------------------------------------------------
#ifndef HDP_HOST_EMU
__device__ __forceinline__ double hilo2d(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint32_t d2hi(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ uint32_t d2lo(double x) { return (uint32_t)__double2loint(x); }
__device__ __forceinline__ float u2f(uint32_t b) { return __uint_as_float(b); }
__device__ __forceinline__ double rcp_approx(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double rsqrt_approx(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double xfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return fmaf(a, b, c); }
// MUFU approximations with flush-to-zero: the arguments are never denormal here, and the
// non-ftz forms carry denormal fix-up code (extra FSETP/FMUL per call)
__device__ __forceinline__ float rsq_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x)
{
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x)
{
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// ---------------------------------------------------------------------
// Uniform and normal draws.
// f64: 52-bit uniforms.  Words (w_hi, w_lo) -> k = (w_hi >> 12) : w_lo in [0, 2^52);
//      the double with exponent field 0x3ff and mantissa k is 1 + k*2^-52, so
//      u = that - 1 in [0,1)                       (one exact DADD, no I2F)
//      u1 = that - (1 - 2^-53) = (k + 1/2)*2^-52 in (0,1)   (Box–Muller radius)
// Box–Muller: z0 = R cos(T), z1 = R sin(T), R = sqrt(-2 ln u1), T = 2π k2 / 2^52.
//      -2 ln u1: exponent/mantissa split, ln(m) = 2 atanh((m-1)/(m+1)), MUFU.RCP64H seed
//      + 2 Newton steps for the quotient, degree-8 series — no special cases since
//      u1 is a normal number in (0,1).
//      sqrt: MUFU.RSQ64H seed + 1 Newton step + Heron correction.
//      sin/cos: octant from the top 3 bits of k2 (exact integer range reduction, odd
//      octants mirrored in integers), Taylor polynomials on [0, π/4].
// f32: 23-bit uniforms by the same trick; Box–Muller on the MUFU pipe.
// ---------------------------------------------------------------------
#ifndef HDP_HOST_EMU
#define HDP_CONST static __constant__
#else
#define HDP_CONST static const
#endif
// atanh series 1/3 .. 1/19, Taylor sin (-1/3! .. -1/15!) and cos (-1/2! .. 1/16!)
HDP_CONST double kAtanhC[9] = {1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0,
                               1.0 / 17.0, 1.0 / 19.0};
HDP_CONST double kSinC[7] = {-1.6666666666666666667e-01, 8.3333333333333333333e-03, -1.9841269841269841270e-04,
                             2.7557319223985890653e-06, -2.5052108385441718775e-08, 1.6059043836821614599e-10,
                             -7.6471637318198164759e-13};
HDP_CONST double kCosC[8] = {-0.5, 4.1666666666666666667e-02, -1.3888888888888888889e-03, 2.4801587301587301587e-05,
                             -2.7557319223985890653e-07, 2.0876756987868098979e-09, -1.1470745597729724714e-11,
                             4.7794773323873852974e-14};
HDP_CONST double kMisc[6] = {-4.638093627692599e-17, -1.3862943611198906, 4503601774854144.0, 4503599627370496.0,
                             1.3951473992034527e-15, 0.99999999999999988898};

__device__ __forceinline__ double neg2log_unit(double u)
{
    uint32_t hi = d2hi(u);
    const uint32_t lo = d2lo(u);
    int e = (int)(hi >> 20) - 1023;                // u = 2^e * 1.f
    hi = (hi & 0x000fffffu) | 0x3ff00000u;          // m in [1,2)
    const uint32_t adj = hi >= 0x3ff6a09fu ? 1u : 0u; // m >= ~sqrt(2): halve
    hi -= adj << 20;
    e += (int)adj;
    const double m = hilo2d(hi, lo);                // m in [~0.7071, ~1.4142)
    const double f = m - 1.0;
    const double dn = m + 1.0;
    double r = rcp_approx(dn);
    double t = xfma(-dn, r, 1.0);
    r = xfma(r, t, r);
    t = xfma(-dn, r, 1.0);
    r = xfma(r, t, r);
    const double sv = f * r;                        // s = (m-1)/(m+1), |s| <= 0.1716
    const double s2 = sv * sv;
    // atanh(s)/s - 1 = s2/3 + s2^2/5 + ... + s2^9/19
    // even/odd split: two half-length Horner chains in s4 = s2^2 (shorter dependent latency)
    const double s4 = s2 * s2;
    double pe = kAtanhC[8];
    pe = xfma(pe, s4, kAtanhC[6]);
    pe = xfma(pe, s4, kAtanhC[4]);
    pe = xfma(pe, s4, kAtanhC[2]);
    pe = xfma(pe, s4, kAtanhC[0]);
    double po = kAtanhC[7];
    po = xfma(po, s4, kAtanhC[5]);
    po = xfma(po, s4, kAtanhC[3]);
    po = xfma(po, s4, kAtanhC[1]);
    const double pl = xfma(po, s2, pe);
    const double lm2 = xfma(sv * s2, pl, sv);       // atanh(s) = ln(m)/2
    // e as a double without I2F: (2^52 + 2^31 + e) - (2^52 + 2^31)
    const double ed = hilo2d(0x43300000u, (uint32_t)e ^ 0x80000000u) - kMisc[2];
    // -2 ln u = -2 e ln2 - 4 atanh(s)
    double res = xfma(ed, kMisc[0], -4.0 * lm2); // -2 ln2, low part
    return xfma(ed, kMisc[1], res);              // -2 ln2, high part
}

__device__ __forceinline__ double sqrt_pos(double x)
{
    double y = rsqrt_approx(x);
    const double h = 0.5 * x;
    double t = y * y;
    t = xfma(-h, t, 1.5);
    y = y * t;
    double sq = x * y;
    const double er = xfma(-sq, sq, x);
    return xfma(er, 0.5 * y, sq);
}

// (cos, sin) of 2π k / 2^52 for the 52-bit integer k = (kh : kl), kh < 2^20
__device__ __forceinline__ void sincos_turn52(uint32_t kh, uint32_t kl, double& sn, double& cs)
{
    const uint32_t q = kh >> 17;                    // octant 0..7
    u64 f = ((u64)(kh & 0x1ffffu) << 32) | (u64)kl; // 49-bit fraction within the octant
    if (q & 1u) f = ((u64)1 << 49) - f;             // mirror odd octants (exact)
    const double t = hilo2d(0x43300000u | (uint32_t)(f >> 32), (uint32_t)f) - kMisc[3]; // f as double
    const double a = t * kMisc[4];                  // (π/4) * 2^-49, a in [0, π/4]
    const double a2 = a * a;
    const double a4 = a2 * a2;
    double se = kSinC[6];
    se = xfma(se, a4, kSinC[4]);
    se = xfma(se, a4, kSinC[2]);
    se = xfma(se, a4, kSinC[0]);
    double so = kSinC[5];
    so = xfma(so, a4, kSinC[3]);
    so = xfma(so, a4, kSinC[1]);
    const double ps = xfma(so, a2, se);
    const double sa = xfma(a * a2, ps, a);
    double ce = kCosC[6];
    ce = xfma(ce, a4, kCosC[4]);
    ce = xfma(ce, a4, kCosC[2]);
    ce = xfma(ce, a4, kCosC[0]);
    double co = kCosC[7];
    co = xfma(co, a4, kCosC[5]);
    co = xfma(co, a4, kCosC[3]);
    co = xfma(co, a4, kCosC[1]);
    const double pc = xfma(co, a2, ce);
    const double ca = xfma(a2, pc, 1.0);
    const bool swap = ((q + 1u) & 2u) != 0;          // octants 1,2,5,6
    double s0 = swap ? ca : sa;
    double c0 = swap ? sa : ca;
    // signs: sin < 0 in octants 4..7, cos < 0 in octants 2..5
    const uint32_t ssign = (q & 4u) << 29;
    const uint32_t csign = ((q + 2u) & 4u) << 29;
    sn = hilo2d(d2hi(s0) ^ ssign, d2lo(s0));
    cs = hilo2d(d2hi(c0) ^ csign, d2lo(c0));
}

template <typename T>
__device__ __forceinline__ Vec<T> draw_vec(u64 seed, int stream, u64 ord, uint32_t group, bool normal);

template <>
__device__ __forceinline__ Vec<double> draw_vec<double>(u64 seed, int stream, u64 ord, uint32_t group, bool normal)
{
    const Philox4 p = philox4x32_10((uint32_t)ord, (uint32_t)(ord >> 32), group, (uint32_t)stream,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    Vec<double> r;
    const uint32_t h1 = p.w[0] >> 12, h2 = p.w[2] >> 12;
    if (!normal) {
        r.v[0] = hilo2d(0x3ff00000u | h1, p.w[1]) - 1.0;
        r.v[1] = hilo2d(0x3ff00000u | h2, p.w[3]) - 1.0;
    } else {
        const double u1 = hilo2d(0x3ff00000u | h1, p.w[1]) - kMisc[5]; // 1 - 2^-53
        const double rad = sqrt_pos(neg2log_unit(u1));
        double sn, cs;
        sincos_turn52(h2, p.w[3], sn, cs);
        r.v[0] = rad * cs;
        r.v[1] = rad * sn;
    }
    return r;
}

template <>
__device__ __forceinline__ Vec<float> draw_vec<float>(u64 seed, int stream, u64 ord, uint32_t group, bool normal)
{
    const Philox4 p = philox4x32_10((uint32_t)ord, (uint32_t)(ord >> 32), group, (uint32_t)stream,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    Vec<float> r;
    if (!normal) {
#pragma unroll
        for (int e = 0; e < 4; ++e) r.v[e] = u2f(0x3f800000u | (p.w[e] >> 9)) - 1.0f;
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float u1 = u2f(0x3f800000u | (p.w[2 * h] >> 9)) - 0.99999994f; // (k + 1/2) 2^-23 in (0,1)
            const float u2 = u2f(0x3f800000u | (p.w[2 * h + 1] >> 9)) - 1.0f;
            // -2 ln u1 = -2 ln2 * log2(u1)
            const float tt = -1.3862943611198906f * lg2_approx(u1); // > 0: u1 < 1
            const float rad = tt * rsq_approx(tt);
            const float th = 6.283185307179586f * u2;
            r.v[2 * h] = rad * cos_approx(th);
            r.v[2 * h + 1] = rad * sin_approx(th);
        }
    }
    return r;
}

// ---------------------------------------------------------------------
// device copies of the POD descriptors (double parameters → T once, on host)
// ---------------------------------------------------------------------
template <typename T> struct DevProblem {
    int d;
    int mu_kind, sigma_kind, g_kind, f_kind;
    int has_bc, bc_is_scalar;
    T t0, t1;
    T mu_c, sigma_c;  // p[0] of μ, σ
    T mu_p[2], sigma_p[2];
    T g_p[4];
    T f_p[8];
    double f_pd[2];   // replicator–mutator weights precomputed in fp64
    double g_pd[2];   // gaussian-density prefactor
    T lo_s, hi_s;
    const T* lo;      // device, length d (when !bc_is_scalar)
    const T* hi;
};
template <typename T> struct DevSampler {
    int kind, shifted, scalar;
    T a_s, b_s, sigma;
    const T* a;       // device, length d
    const T* b;
};

// μ(x,p,t)[i], σ(x,p,t)[i]
template <typename T> __device__ __forceinline__ T coef(int kind, T c, T xi)
{
    return kind == HDP_COEF_LINEAR ? mul_rn(c, xi) : c;
}

// user-defined μ(x,p,t)[i], σ(x,p,t)[i] (HDP_COEF_USER)
template <typename T> __device__ __forceinline__ T user_mu(T x, int i, T t, int d, const T* p)
{
    (void)x; (void)i; (void)t; (void)d; (void)p;
    return (T)(HDP_USER_MU_EXPR);
}
template <typename T> __device__ __forceinline__ T user_sigma(T x, int i, T t, int d, const T* p)
{
    (void)x; (void)i; (void)t; (void)d; (void)p;
    return (T)(HDP_USER_SIGMA_EXPR);
}

// MLP step, src/MLP.jl:417:  x + ((μ*dt) + ((σ*sq)*ξ))
template <typename T> __device__ __forceinline__ T em_mlp(T x, T mu, T sg, T dt, T sq, T xi)
{
    return add_rn(x, add_rn(mul_rn(mu, dt), mul_rn(mul_rn(sg, sq), xi)));
}
// DeepSplitting step, src/DeepSplitting.jl:160:  (y0 + μ*dt) + ((σ*sq)*dW)
template <typename T> __device__ __forceinline__ T em_ds(T y0, T mu, T sg, T dt, T sq, T dw)
{
    return add_rn(add_rn(y0, mul_rn(mu, dt)), mul_rn(mul_rn(sg, sq), dw));
}

// ---------------------------------------------------------------------
// Neumann reflection of one coordinate against [lo, hi].
// In an axis-aligned box the reference's sequential billiard
// (src/reflect.jl:31-53) decouples per coordinate: coordinate i is mirrored
// at each of its own wall hits, b_i <- b_i - 2*(b_i - c_i) with c_i the hit
// point (the wall).  We iterate exactly that mirror; nb counts wall hits.
// A far excursion (> 64 box widths, or a degenerate box) is folded in closed
// form by a cold out-of-line routine.
// ---------------------------------------------------------------------
template <typename T> struct FarResult {
    T b;
    int hits;
};
// (returns the hit count by value: a pointer argument would force the caller's counter into memory)
template <typename T> __device__ HDP_NOINLINE FarResult<T> reflect_far(T b, T lo, T hi)
{
    FarResult<T> res;
    res.hits = 0;
    const T w = sub_rn(hi, lo);
    if (!(w > (T)0)) { res.b = b != b ? b : lo; return res; } // degenerate box
    const T y = sub_rn(b, lo);
    const T per = (T)2 * w;
    T z = fmod(y, per);
    if (z < (T)0) z += per;
    const T q = y > (T)0 ? ceil(y / w) - (T)1 : floor(-y / w) + (T)1; // wall hits, closed form
    res.hits = (int)fmin(q, (T)1.0e9);
    b = add_rn(lo, z <= w ? z : sub_rn(per, z));
    if (b < lo) b = lo;
    if (b > hi) b = hi;
    res.b = b;
    return res;
}

template <typename T> __device__ __forceinline__ T mirror_once(T b, T wall)
{
    return sub_rn(b, mul_rn((T)2, sub_rn(b, wall))); // b - 2*(b - c), src/reflect.jl:34 for the hit coordinate
}

template <typename T> __device__ __forceinline__ T reflect_coord(T b, T lo, T hi, int& nb)
{
    bool below = b < lo, above = b > hi;
    if (below | above) {
        if (!(fabs(sub_rn(b, lo)) <= (T)64 * sub_rn(hi, lo))) {
            const FarResult<T> fr = reflect_far<T>(b, lo, hi);
            nb += fr.hits;
            return fr.b;
        }
        do {
            b = mirror_once<T>(b, below ? lo : hi);
            ++nb;
            below = b < lo;
            above = b > hi;
        } while (below | above);
    }
    return b;
}

// VEC coordinates of one lane at once.  The first wall hit is applied branch-free
// (selects) to every coordinate; only coordinates that are still outside afterwards — a
// second hit in the same step — enter the per-coordinate loop.  No collective inside.
template <typename T, int G, bool COUNT = true>
__device__ __forceinline__ void reflect_vec(Vec<T>& x, const Vec<T>& lo, const Vec<T>& hi, unsigned gmask, int& nb)
{
    constexpr int VEC = Tr<T>::VEC;
    (void)gmask;
    bool again = false;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const bool below = x.v[e] < lo.v[e];
        const bool above = x.v[e] > hi.v[e];
        const bool o = below | above;
        const T t = mirror_once<T>(x.v[e], below ? lo.v[e] : hi.v[e]);
        x.v[e] = o ? t : x.v[e];
        if (COUNT) nb += o ? 1 : 0;
        again |= o && ((x.v[e] < lo.v[e]) | (x.v[e] > hi.v[e]));
    }
    if (again) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) x.v[e] = reflect_coord<T>(x.v[e], lo.v[e], hi.v[e], nb);
    }
}

// ---------------------------------------------------------------------
// samplers, src/MCSample.jl:28-33 and :63-71
// ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T uniform_map(T u, T a, T b)
{
    T m = div_rn(add_rn(b, a), (T)2);                          // :31
    return add_rn(mul_rn(sub_rn(u, (T)0.5), sub_rn(b, a)), m); // :32
}
template <typename T> __device__ __forceinline__ T normal_map(T xi, T sigma, bool shifted, T xref)
{
    T v = mul_rn(xi, sigma);            // :65
    return shifted ? add_rn(v, xref) : v; // :70
}

// ---------------------------------------------------------------------
// g and f kinds.  Points enter only through sum(x^2) and min(x).
// ---------------------------------------------------------------------
template <typename T> struct Feat {
    T ssq;
    T mn;
};

__device__ __forceinline__ double xexp(double x) { return exp(x); }
__device__ __forceinline__ float xexp(float x) { return expf(x); }

template <typename T> __device__ __forceinline__ T g_eval(const DevProblem<T>& pb, const Feat<T>& f)
{
    switch (pb.g_kind) {
    case HDP_G_SUMSQ: return f.ssq + pb.g_p[0];
    case HDP_G_INV_SUMSQ: return (T)1 / (pb.g_p[0] + pb.g_p[1] * f.ssq);
    case HDP_G_EXP_SUMSQ: return xexp(-pb.g_p[0] * f.ssq);
    case HDP_G_MIN: return f.mn;
    case HDP_G_GAUSS: return (T)(pb.g_pd[0] * exp(-0.5 * (double)f.ssq / pb.g_pd[1]));
    case HDP_G_USER: {
        const T ssq = f.ssq, mn = f.mn;
        const int d = pb.d;
        const T* p = pb.g_p;
        (void)ssq; (void)mn; (void)d; (void)p;
        return (T)(HDP_USER_G_EXPR);
    }
    }
    return (T)0;
}

template <typename T> __device__ __forceinline__ T maxz(T u) { return u > (T)0 ? u : (T)0; }
template <typename T> __device__ __forceinline__ T allen_cahn(T u) { return u - u * u * u; }

template <typename T>
__device__ __forceinline__ T f_eval(const DevProblem<T>& pb, T ssq_y, T ssq_z, T vy, T vz)
{
    const T* p = pb.f_p;
    switch (pb.f_kind) {
    case HDP_F_ZERO: return (T)0 * vy;
    case HDP_F_LINEAR: return p[0] * vy;
    case HDP_F_ALLEN_CAHN: return p[0] * allen_cahn(vy);
    case HDP_F_DEFAULT_RISK: {
        T delta = p[0], R = p[1], vh = p[2], vl = p[3], gh = p[4], gl = p[5];
        T u = vy;
        T Q = (T)(u < vh) * gh + (T)(u >= vl) * gl +
              (T)((u >= vh) && (u < vl)) * (((gh - gl) / (vh - vl)) * (u - vh) + gh);
        return -((T)1 - delta) * Q * vy - R * vy;
    }
    case HDP_F_FISHER_KPP: return maxz(vy) * ((T)1 - maxz(vy));
    case HDP_F_NL_ALLEN_CAHN_DIFF: return allen_cahn(vy) - allen_cahn(vz);
    case HDP_F_NL_FISHER: return maxz(vy) * ((T)1 - maxz(vz));
    case HDP_F_NL_REPLICATOR: {
        // f_pd[0] = (2π)^(d/2) σs^d, f_pd[1] = 0.5/σs²
        double my = -0.5 * (double)ssq_y, mz = -0.5 * (double)ssq_z;
        double w = pb.f_pd[0] * exp(pb.f_pd[1] * (double)ssq_z);
        return (T)((double)maxz(vy) * (my - (double)maxz(vz) * mz * w));
    }
    case HDP_F_USER:
    case HDP_F_NL_USER: {
        const T t = pb.t1; // f is always called with the outer end time (src/MLP.jl:140-146, quirk Q2)
        const int d = pb.d;
        (void)t; (void)d; (void)ssq_y; (void)ssq_z; (void)vz;
        return (T)(HDP_USER_F_EXPR);
    }
    }
    return (T)0;
}

} // namespace hdp
