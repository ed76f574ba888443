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
#include <cuda_runtime.h>
#define HDP_NOINLINE __noinline__
#endif
#include <stdint.h>
#include "../../include/hdpde.h"

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
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

__device__ __forceinline__ double bits53(uint32_t hi, uint32_t lo)
{ // integer k in [0, 2^53) as a double
    return (double)(((u64)(hi >> 5) << 26) | (u64)(lo >> 6));
}

// VEC draws (uniform in [0,1) or standard normal) for (stream, ordinal, group).
// f64: 53-bit uniforms; Box–Muller with fp64 log / sincospi (u1 in (0,1]).
template <typename T>
__device__ __forceinline__ Vec<T> draw_vec(u64 seed, int stream, u64 ord, uint32_t group, bool normal);

template <>
__device__ __forceinline__ Vec<double> draw_vec<double>(u64 seed, int stream, u64 ord, uint32_t group, bool normal)
{
    Philox4 p = philox4x32_10((uint32_t)ord, (uint32_t)(ord >> 32), group, (uint32_t)stream,
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    Vec<double> r;
    double k1 = bits53(p.w[0], p.w[1]);
    double k2 = bits53(p.w[2], p.w[3]);
    if (!normal) {
        r.v[0] = k1 * 0x1.0p-53;
        r.v[1] = k2 * 0x1.0p-53;
    } else {
        double u1 = (k1 + 1.0) * 0x1.0p-53;
        double u2 = k2 * 0x1.0p-53;
        double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        r.v[0] = rad * c;
        r.v[1] = rad * s;
    }
    return r;
}

// f32: 24-bit uniforms; Box–Muller on the MUFU pipe (lg2 / sin / cos approx).
template <>
__device__ __forceinline__ Vec<float> draw_vec<float>(u64 seed, int stream, u64 ord, uint32_t group, bool normal)
{
    Philox4 p = philox4x32_10((uint32_t)ord, (uint32_t)(ord >> 32), group, (uint32_t)stream,
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    Vec<float> r;
    if (!normal) {
#pragma unroll
        for (int e = 0; e < 4; ++e) r.v[e] = (float)(p.w[e] >> 8) * 0x1.0p-24f;
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float u1 = ((float)(p.w[2 * h] >> 8) + 1.0f) * 0x1.0p-24f;
            float u2 = (float)(p.w[2 * h + 1] >> 8) * 0x1.0p-24f;
            // -2 ln u1 = -2 ln2 * log2(u1)
            float rad = sqrtf(-1.3862943611198906f * __log2f(u1));
            float th = 6.283185307179586f * u2;
            r.v[2 * h] = rad * __cosf(th);
            r.v[2 * h + 1] = rad * __sinf(th);
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
    T mu_c, sigma_c;
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
// A far excursion (> 64 box widths) is first folded in closed form.
// ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T reflect_coord(T b, T lo, T hi, int& nb)
{
    const T w = sub_rn(hi, lo);
    if (!(w > (T)0)) return b != b ? b : lo; // degenerate box
    T y = sub_rn(b, lo);
    if (fabs(y) > (T)64 * w) {
        T per = (T)2 * w;
        T z = fmod(y, per);
        if (z < (T)0) z += per;
        T q = y > (T)0 ? ceil(y / w) - (T)1 : floor(-y / w) + (T)1; // wall hits, closed form
        nb += (int)fmin(q, (T)1.0e9);
        b = add_rn(lo, z <= w ? z : sub_rn(per, z));
        if (b < lo) b = lo;
        if (b > hi) b = hi;
        return b;
    }
    for (int it = 0; it < 256; ++it) {
        if (b < lo) b = sub_rn(b, mul_rn((T)2, sub_rn(b, lo)));
        else if (b > hi) b = sub_rn(b, mul_rn((T)2, sub_rn(b, hi)));
        else break;
        ++nb;
    }
    return b;
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
    }
    return (T)0;
}

} // namespace hdp
