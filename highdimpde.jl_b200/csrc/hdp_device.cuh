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
// Philox4x32-R with a FIXED key (R = HDP_PHILOX_ROUNDS, include/hdpde.h: 7, the fewest rounds
// Salmon et al. (SC'11) report as Crush-resistant; their default of 10 buys margin the Monte-Carlo
// estimators here do not need and costs 12 of the ~147 instructions of a vec-group); the 64-bit
// seed travels in the counter:
//   ctr = (w0, w1, seed_lo, seed_hi),   (w0, w1) = draw_address(stream, ordinal, group)
// Addressing by consumption ordinal — never by thread id — makes every draw independent of
// launch geometry and GPU count.  The fixed key makes the ten round keys immediates of the
// two XORs of a round (2 IMAD.WIDE + 2 LOP3 per round, no key registers), whatever seed each
// lane group of a warp works on.  The word that varies fastest inside a loop (the group
// inside a path-step, the step inside `sde_loop!`) sits in w1, which enters the first round
// only through an XOR: the compiler hoists the loop-invariant multiplies of rounds 1–3.
// ---------------------------------------------------------------------
struct Philox4 { uint32_t w[4]; };

constexpr uint32_t kPhiloxKey0 = 0xA4093822u, kPhiloxKey1 = 0x299F31D0u; // fractional digits of pi

__device__ __forceinline__ Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
#pragma unroll
    for (int r = 0; r < HDP_PHILOX_ROUNDS; ++r) {
        const uint32_t k0 = kPhiloxKey0 + (uint32_t)r * 0x9E3779B9u, k1 = kPhiloxKey1 + (uint32_t)r * 0xBB67AE85u;
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// (w0, w1) of a draw.  Streams of the MLP tree (TIME, STEP, SAMPLER): ordinal < 2^40 (the
// reference's consumption ordinal), group < 2^21 (vec-group inside a point).  Streams of the
// batched loop (SDE, X0, MC): ordinal < 2^24 (step / sample index), group < 2^37 (vec-group
// of the whole d×B array).  The host checks the limits (draw_address_ok).
__device__ __forceinline__ void draw_address(int stream, u64 ord, u64 group, uint32_t& w0, uint32_t& w1)
{
    if (stream <= HDP_STREAM_SAMPLER) {
        w0 = (uint32_t)ord;
        w1 = (uint32_t)(ord >> 32) | ((uint32_t)group << 8) | ((uint32_t)stream << 29);
    } else {
        w0 = (uint32_t)group;
        w1 = (uint32_t)(group >> 32) | ((uint32_t)ord << 5) | ((uint32_t)stream << 29);
    }
}
__host__ __device__ inline bool draw_address_ok(int stream, u64 n_ord, u64 n_group)
{
    if (stream <= HDP_STREAM_SAMPLER) return n_ord <= ((u64)1 << 40) && n_group <= ((u64)1 << 21);
    return n_ord <= ((u64)1 << 24) && n_group <= ((u64)1 << 37);
}

// ---------------------------------------------------------------------
// bit-level helpers (host emulation provides the same functions)
// ---------------------------------------------------------------------
#ifndef HDP_HOST_EMU
__device__ __forceinline__ double hilo2d(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint32_t d2hi(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ uint32_t d2lo(double x) { return (uint32_t)__double2loint(x); }
__device__ __forceinline__ float u2f(uint32_t b) { return __uint_as_float(b); }
__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
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
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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
//      u1 = (k | 1)*2^-52 in (0,1)                 (Box–Muller radius; k | 1 converted by one I2F.F64.U64)
// Box–Muller: z0 = R cos(T), z1 = R sin(T), R = sqrt(-2 ln u1), T = 2π k2 / 2^52.
//   Both transcendentals are table driven; the tables (12 KB, hdp_tables.h) are staged in
//   shared memory by tables_init() at kernel start.
//   -2 ln u1: u1 = 2^e * m, m in [1,2).  idx = m rounded to 8 fraction bits (0..256) selects
//      rc ≈ 1/(1 + idx/256) and lm2 = -2 ln(1/rc) (shifted by ln 2 for idx >= 107, so the
//      entries next to u1 = 1 are exactly zero and nothing cancels there); r = fma(m, rc, -1),
//      |r| <= 2^-9, log1p(r) by its degree-6 Taylor polynomial.  No special cases: u1 is a
//      normal number in (0,1).  Agrees with glibc's log to 2 ulp.
//   sqrt: MUFU.RSQ64H seed + 1 Goldschmidt step + Heron correction.
//   sin/cos: the top 9 bits of k2 select (sin, cos) at the centre of a 1/512 turn, the other
//      43 bits give the offset |δ| <= π/512 exactly; sin δ by its degree-5, cos δ − 1 by its degree-4
//      Taylor polynomial, and the angle-sum formulas.  No octant logic.
//   32 fp64 instructions and two I2F per pair of normals (round 1: 38).
// f32: 23-bit uniforms by the same trick; Box–Muller on the MUFU pipe.
// ---------------------------------------------------------------------
#ifndef HDP_HOST_EMU
#define HDP_TABLE static __device__ const
#else
#define HDP_TABLE static const
#endif
#include "hdp_tables.h"

// Constants of the fp64 transform.  The five that need all 53 bits are handed to the draw
// functions in a DrawK, which hot loops build once (draw_k_init) as opaque registers — left
// to itself the compiler re-materialises each of them with two UMOVs per loop iteration.
// The rest multiply small corrections: they are rounded to 21 significant bits so that they
// encode as 32-bit immediates of DFMA.
#define HDP_C5T 0x1.9999ap-3    /* 1/5   */
#define HDP_C6T 0x1.55555p-3    /* 1/6   */
#define HDP_C120T 0x1.11111p-7  /* 1/120 */

template <typename T> struct DrawK {};
template <> struct DrawK<double> {
    double c3, two_ln2, turn, c6, c24; // 1/3, 2 ln 2 (high part), 2π/2^52, 1/6, 1/24
    uint32_t log_tab, sincos_tab;      // shared-window addresses of the tables
};

#ifndef HDP_HOST_EMU
__shared__ double2 s_log_tab[HDP_LOG_N + 1];
__shared__ double2 s_sincos_tab[HDP_SINCOS_N];
// every kernel that draws fp64 normals calls this first, with all threads of the block
__device__ __forceinline__ void tables_init()
{
    for (int i = threadIdx.x; i < HDP_LOG_N + 1; i += blockDim.x) s_log_tab[i] = make_double2(kLogTab[i][0], kLogTab[i][1]);
    for (int i = threadIdx.x; i < HDP_SINCOS_N; i += blockDim.x)
        s_sincos_tab[i] = make_double2(kSinCosTab[i][0], kSinCosTab[i][1]);
    __syncthreads();
}
__device__ __forceinline__ void tab_entry(uint32_t base, uint32_t idx, double& a, double& b)
{
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(base + idx * 16u));
}
// `zero` is 0.  Hot loops pass a zero the compiler cannot see through (a kernel parameter),
// which pins the five constants and the two table addresses in registers for the whole loop;
// cold call sites pass the literal and everything folds.
__device__ __forceinline__ double opaque_const(double c, int zero)
{
    return hilo2d(d2hi(c) + (uint32_t)zero, d2lo(c) + (uint32_t)zero);
}
template <typename T> __device__ __forceinline__ DrawK<T> draw_k_init(int zero)
{
    DrawK<T> k;
    if constexpr (sizeof(T) == 8) {
        k.c3 = opaque_const(1.0 / 3.0, zero);
        k.two_ln2 = opaque_const(HDP_TWO_LN2_HI, zero);
        k.turn = opaque_const(HDP_TURN52, zero);
        k.c6 = opaque_const(1.0 / 6.0, zero);
        k.c24 = opaque_const(1.0 / 24.0, zero);
        k.log_tab = (uint32_t)__cvta_generic_to_shared(s_log_tab) + (uint32_t)zero;
        k.sincos_tab = (uint32_t)__cvta_generic_to_shared(s_sincos_tab) + (uint32_t)zero;
    }
    return k;
}
#else
static inline void tables_init() {}
static inline void tab_entry(uint32_t base, uint32_t idx, double& a, double& b)
{
    const double(*t)[2] = base ? kSinCosTab : kLogTab;
    a = t[idx][0]; b = t[idx][1];
}
template <typename T> static inline DrawK<T> draw_k_init(int)
{
    DrawK<T> k;
    if constexpr (sizeof(T) == 8) {
        k.c3 = 1.0 / 3.0; k.two_ln2 = HDP_TWO_LN2_HI; k.turn = HDP_TURN52; k.c6 = 1.0 / 6.0; k.c24 = 1.0 / 24.0;
        k.log_tab = 0; k.sincos_tab = 1;
    }
    return k;
}
#endif

// -2 ln(u1) for u1 = k 2^-52, k = (kh : kl) an integer in [1, 2^52)
// (k is converted exactly by one I2F.F64.U64 — the conversion pipe, one dispatch cycle — and only the mantissa
// and the exponent of the result are used, so the scaling by 2^-52 is a constant in the exponent arithmetic;
// building u1 = (1 + k 2^-52) - (1 - 2^-53) from the bits took an OR and a DADD)
__device__ __forceinline__ double neg2log_k52(uint32_t kh, uint32_t kl, const DrawK<double>& k)
{
#ifndef HDP_GEN_I2F
#define HDP_GEN_I2F 1 // 0: the bit-built conversions (A/B: profiles/r2t_generator_ab.txt)
#endif
#if HDP_GEN_I2F
    const double u = (double)(((u64)kh << 32) | (u64)kl); // k, exactly
#else
    const double u = (hilo2d(0x3ff00000u | kh, kl) - 1.0) * 0x1p+52;
#endif
    const uint32_t hi = d2hi(u), lo = d2lo(u);
    const uint32_t mh = (hi & 0x000fffffu) | 0x3ff00000u;   // m = 1.f in [1,2)
    static_assert(HDP_LOG_N == 256 && HDP_SINCOS_N == 512, "the index arithmetic below is written for these table sizes");
    const uint32_t idx = (mh - 0x3feff800u) >> 12;          // f rounded to 8 bits (to nearest): 0..256
    double rc, lm2;
    tab_entry(k.log_tab, idx, rc, lm2);
    const double r = xfma(hilo2d(mh, lo), rc, -1.0);        // |r| <= 2^-9
    // ne = -(e + [idx >= HDP_LOG_SPLIT]) >= 0
    const uint32_t ne = (1023u + 52u) - (hi >> 20) - ((idx + (256u - HDP_LOG_SPLIT)) >> 8);
#if HDP_GEN_I2F
    const double ned = (double)ne; // I2F.F64.U32 on the conversion pipe: one dispatch cycle, where (2^52 + ne) - 2^52 was a DADD
#else
    const double ned = hilo2d(0x43300000u, ne) - 0x1p+52;
#endif
    // the large terms first (off the critical path: they do not need the polynomial)
    // (the low part of 2 ln 2, 2^-54, times ne <= 52 is at most 0.3 ulp of the result anywhere: not worth an fp64 issue slot)
    const double big = xfma(ned, k.two_ln2, lm2);
    // log1p(r) = r + r^2 q(r), q = (-1/2 + r/3) + r^2 ((-1/4 + r/5) - r^2/6); the next term, r^7/7, is below
    // 2^-56 of the result for |r| <= 2^-9.  Two independent first-level terms: dependent depth 4.
    const double r2 = r * r;
    const double qa = xfma(r, k.c3, -0.5);
    const double qb = xfma(r, HDP_C5T, -0.25);
    const double q = xfma(r2, xfma(r2, -HDP_C6T, qb), qa);
    const double p = xfma(r2, q, r);
    return xfma(p, -2.0, big);
}

// sqrt of a positive normal double: MUFU.RSQ64H seed (≈2^-22), one Goldschmidt step on g ≈ sqrt x with
// h ≈ 1/(2 sqrt x) and a Heron correction; 6 operations, dependent depth 5; within 1 ulp of sqrt
__device__ __forceinline__ double sqrt_pos(double x)
{
    const double y = rsqrt_approx(x);
    const double g = x * y;
    // h = y / 2 by one integer subtraction from the exponent field (y is a normal number in [0.1, 2^27]: x is
    // -2 ln u1 in [2^-52, 73]): an ALU instruction holds the dispatch port for one cycle, a DMUL for two
    const double h = hilo2d(d2hi(y) - 0x00100000u, d2lo(y));
    const double r = xfma(-g, h, 0.5);
    const double g1 = xfma(g, r, g);       // relative error 2^-43
    const double e = xfma(-g1, g1, x);     // exact residual
    return xfma(e, h, g1);                 // h is good to 2^-22: the correction term is off by 2^-65 of the result
                                           // (refining h as well, as a full Goldschmidt step would, buys nothing)
}

// (sin, cos) of 2π k / 2^52 for the 52-bit integer k = (kh : kl), kh < 2^20
__device__ __forceinline__ void sincos_turn52(uint32_t kh, uint32_t kl, double& sn, double& cs, const DrawK<double>& k)
{
    double sj, cj;
    tab_entry(k.sincos_tab, kh >> 11, sj, cj);
    // offset from the centre of the 1/512 turn: (f - 2^42) * 2π/2^52, f = low 43 bits of k
    const double dl = (hilo2d(0x43300000u | (kh & 0x7ffu), kl) - 0x1.004p+52 /* 2^52 + 2^42 */) * k.turn;
    const double t = dl * dl;
    const double ps = xfma(t, HDP_C120T, -k.c6);
    const double sd = xfma(dl * t, ps, dl);                 // sin δ
    const double pc = xfma(t, k.c24, -0.5);                 // (δ^6/720 <= 7e-17 for |δ| <= π/512)
    const double cd1 = t * pc;                              // cos δ − 1
    // sin(a + δ) = sj + (cj sin δ + sj (cos δ - 1)) as two chained fmas (one rounding more than summing the two
    // corrections first, one fp64 instruction less per output: <= 1 ulp)
    sn = xfma(sj, cd1, xfma(cj, sd, sj));
    cs = xfma(cj, cd1, xfma(-sj, sd, cj));
}

template <typename T>
__device__ __forceinline__ Vec<T> draw_vec(u64 seed, int stream, u64 ord, u64 group, bool normal, const DrawK<T>& k);

// the 128 random bits of a draw, and the VEC normals made from them (two stages, so that a loop can
// start the integer work of its next pass while the floating-point work of this one is in flight)
__device__ __forceinline__ Philox4 draw_words(u64 seed, int stream, u64 ord, u64 group)
{
    uint32_t w0, w1;
    draw_address(stream, ord, group, w0, w1);
    return philox4x32(w0, w1, (uint32_t)seed, (uint32_t)(seed >> 32));
}
template <typename T> __device__ __forceinline__ Vec<T> normals_from_words(const Philox4& p, const DrawK<T>& k);
template <> __device__ __forceinline__ Vec<double> normals_from_words<double>(const Philox4& p, const DrawK<double>& k)
{
    Vec<double> r;
    const uint32_t h1 = p.w[0] >> 12, h2 = p.w[2] >> 12;
    const double rad = sqrt_pos(neg2log_k52(h1, p.w[1] | 1u, k)); // u1 = (k1 | 1) 2^-52 in (0,1)
    double sn, cs;
    sincos_turn52(h2, p.w[3], sn, cs, k);
    r.v[0] = rad * cs;
    r.v[1] = rad * sn;
    return r;
}
template <> __device__ __forceinline__ Vec<float> normals_from_words<float>(const Philox4& p, const DrawK<float>&)
{
    Vec<float> r;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = u2f(0x3f800000u | (p.w[2 * h] >> 9)) - 0.99999994f; // (k + 1/2) 2^-23 in (0,1)
        // 2π u2 with u2 = m - 1, m in [1,2) built from the bits: the product 2π (m - 1) is formed exactly inside the fma
        // and rounded once — the same value as fl(2π * u2), one instruction less
        const float th = xfma(u2f(0x3f800000u | (p.w[2 * h + 1] >> 9)), 6.283185307179586f, -6.283185307179586f);
        // -2 ln u1 = -2 ln2 * log2(u1)
        const float tt = -1.3862943611198906f * lg2_approx(u1); // > 0: u1 < 1
        const float rad = sqrt_approx(tt);
        r.v[2 * h] = rad * cos_approx(th);
        r.v[2 * h + 1] = rad * sin_approx(th);
    }
    return r;
}

template <>
__device__ __forceinline__ Vec<double> draw_vec<double>(u64 seed, int stream, u64 ord, u64 group, bool normal,
                                                        const DrawK<double>& k)
{
    const Philox4 p = draw_words(seed, stream, ord, group);
    if (normal) return normals_from_words<double>(p, k);
    Vec<double> r;
    r.v[0] = hilo2d(0x3ff00000u | (p.w[0] >> 12), p.w[1]) - 1.0;
    r.v[1] = hilo2d(0x3ff00000u | (p.w[2] >> 12), p.w[3]) - 1.0;
    return r;
}

template <>
__device__ __forceinline__ Vec<float> draw_vec<float>(u64 seed, int stream, u64 ord, u64 group, bool normal,
                                                      const DrawK<float>& k)
{
    const Philox4 p = draw_words(seed, stream, ord, group);
    if (normal) return normals_from_words<float>(p, k);
    Vec<float> r;
#pragma unroll
    for (int e = 0; e < 4; ++e) r.v[e] = u2f(0x3f800000u | (p.w[e] >> 9)) - 1.0f;
    return r;
}
// convenience form for cold call sites
template <typename T>
__device__ __forceinline__ Vec<T> draw_vec(u64 seed, int stream, u64 ord, u64 group, bool normal)
{
    return draw_vec<T>(seed, stream, ord, group, normal, draw_k_init<T>(0));
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

// m with its sign flipped where s is negative
__device__ __forceinline__ double xor_sign(double m, double s) { return hilo2d(d2hi(m) ^ (d2hi(s) & 0x80000000u), d2lo(m)); }
__device__ __forceinline__ float xor_sign(float m, float s) { return u2f(f2u(m) ^ (f2u(s) & 0x80000000u)); }

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

// VEC coordinates of one lane at once.  The first wall hit is applied branch-free (selects)
// to every coordinate; the return value says whether some coordinate is still outside
// afterwards — a second hit in the same step — and needs reflect_coord.  No collective inside.
template <typename T, bool COUNT = true>
__device__ __forceinline__ bool reflect_first(Vec<T>& x, T lo, T hi, int& nb)
{
    constexpr int VEC = Tr<T>::VEC;
    bool again = false;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const bool below = x.v[e] < lo;
        const bool above = x.v[e] > hi;
        const bool o = below | above;
        const T t = mirror_once<T>(x.v[e], below ? lo : hi);
        x.v[e] = o ? t : x.v[e];
        if (COUNT) nb += o ? 1 : 0;
        again |= o && ((t < lo) | (t > hi));
    }
    return again;
}
// Symmetric scalar box [-h, h] (every BASELINE configuration with a boundary): |x| > h is the
// same decision as (x < -h) | (x > h), and the wall that was hit is copysign(h, x).
// With m = h - |x| (negative exactly when the coordinate is outside: x - x rounds to +0, so the
// sign bit of m IS the decision) the mirror b - 2*(b - wall) of src/reflect.jl:34 is
// sign(b) * (|b| + 2 m): |b| - h = -m exactly (rounding is symmetric), 2*(.) is exact, so the one
// fma below rounds like the reference's subtract, double, subtract.  Everything that follows the
// two fp64 operations is integer work on the high word of m: the select predicate, the hit count
// (one LEA.HI) and the test for a second hit — fp64 compares hold the dispatch port twice as long.
// `far` = sym_far_bits(h): a coordinate whose first mirror image is still outside has
// m <= -2h (up to one rounding), i.e. high word of m >= high word of -2h as unsigned integers;
// the test below is that comparison with a margin, so it may report a few coordinates that are
// in fact inside — its callers re-check with reflect_coord, which decides exactly.
__device__ __forceinline__ uint32_t top_bits(double x) { return d2hi(x); }
__device__ __forceinline__ uint32_t top_bits(float x) { return f2u(x); }
template <typename T> __device__ __forceinline__ uint32_t sym_far_bits(T h)
{
    return top_bits(mul_rn((T)-2, h)) - (sizeof(T) == 8 ? 1u : 16u);
}
// First wall hit of VEC coordinates, branch-free.  Returns the largest top_bits(m) of the group:
// >= `far` (sym_far_bits) means some coordinate may still be outside after its first mirror.
// (Predicated fmas written in PTX come back from ptxas as fma + select, so the select is written here.)
template <typename T, bool COUNT = true>
__device__ __forceinline__ uint32_t reflect_first_sym(Vec<T>& x, T h, int& nb)
{
    constexpr int VEC = Tr<T>::VEC;
    uint32_t top = 0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const T v = x.v[e];
        const T a = fabs(v);
        const T m = sub_rn(h, a);
        const T ta = xfma((T)2, m, a);
        const uint32_t mb = top_bits(m);
        x.v[e] = (int32_t)mb < 0 ? xor_sign(ta, v) : v;
        if (COUNT) nb += (int)(mb >> 31);
        top = mb > top ? mb : top;
    }
    return top;
}
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
template <typename T, bool COUNT = true>
__device__ __forceinline__ void reflect_vec_sym(Vec<T>& x, T h, uint32_t far, int& nb)
{
    if (reflect_first_sym<T, COUNT>(x, h, nb) >= far) {
#pragma unroll
        for (int e = 0; e < Tr<T>::VEC; ++e) x.v[e] = reflect_coord<T>(x.v[e], -h, h, nb);
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
