// hdp_host_emu_shim.h — TEST INFRASTRUCTURE ONLY.
// Lets g++ compile the product's device code (hdp_device.cuh, k_mlp.cuh, k_sde.cuh) for the
// host with a single-lane group (G = 1) / a single thread, so the tree planner, the
// consumption-ordinal arithmetic, the combine logic and the K1 step loops can be checked
// against the oracle on a machine without a GPU.  Never compiled into libhdpde_b200.so.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)

struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };

// launch geometry of the emulated K1 kernels (k_sde.cuh): one block of one thread walks the whole grid-stride loop
struct hdp_emu_dim3 { unsigned x, y, z; };
static const hdp_emu_dim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline int __reduce_add_sync(unsigned, int v) { return v; }

// IEEE round-to-nearest primitives (the emulation is built with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
#define __log2f(x) log2f(x)
#define __cosf(x) cosf(x)
#define __sinf(x) sinf(x)
static inline double hilo2d(uint32_t hi, uint32_t lo)
{
    uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline uint32_t d2hi(double x) { uint64_t b; memcpy(&b, &x, 8); return (uint32_t)(b >> 32); }
static inline uint32_t d2lo(double x) { uint64_t b; memcpy(&b, &x, 8); return (uint32_t)b; }
static inline float u2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t f2u(float x) { uint32_t b; memcpy(&b, &x, 4); return b; }
static inline double rsqrt_approx(double x) { return (double)(float)(1.0 / sqrt(x)); }
static inline float rsq_approx(float x) { return 1.0f / sqrtf(x); }
static inline float sqrt_approx(float x) { return sqrtf(x); }
static inline float lg2_approx(float x) { return log2f(x); }
static inline float sin_approx(float x) { return sinf(x); }
static inline float cos_approx(float x) { return cosf(x); }
static inline double xfma(double a, double b, double c) { return fma(a, b, c); }
static inline float xfma(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
    unsigned long long o = *p;
    *p += v;
    return o;
}
static inline int atomicExch(int* p, int v)
{
    int o = *p;
    *p = v;
    return o;
}
