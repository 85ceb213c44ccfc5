// qs_math.cuh -- scalar helpers shared by device code and the host test harness.
//
// Everything the per-env code needs is expressed through these wrappers so the very
// same env/dynamics source compiles (a) into the sm_100a kernels and (b) with plain g++
// into tests/host_harness (test infrastructure: lets the closed form be checked against
// the oracle on the CPU-only dev box).  There is no CPU product path.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define QS_HD __host__ __device__ __forceinline__
#define QS_D __device__ __forceinline__
#else
#define QS_HD inline
#define QS_D inline
#endif

namespace qs {

QS_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// a constant in the lane type R (float here; the two-lane f2 specialisation lives in qs_pack2.cuh)
template <class R> QS_HD R splat_(float c);
template <> QS_HD float splat_<float>(float c) { return c; }

QS_HD float abs_(float x) { return fabsf(x); }

QS_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
    float r;                                // one MUFU.RSQ (<= 2 ulp); callers pass normal-range arguments
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}

QS_HD float sqrt_(float x) { return sqrtf(x); }

// accurate sin/cos (state-carrying quantities).  |x| <= 0.5 -- every call on the hot path: the quaternion step
// angle |w| dt / 2 and the half Euler angles of a reset -- takes a Taylor pair (sin to x^11, cos to x^10: truncation
// < 2e-11, i.e. below float32 rounding) instead of libm's range reduction; identical code on host and device.
QS_HD void sincos_(float x, float* s, float* c) {
    if (fabsf(x) <= 0.5f) {
        const float z = x * x;
        float ps = -2.5052108385441720e-08f;
        ps = fma_(ps, z, 2.7557319223985893e-06f);
        ps = fma_(ps, z, -1.9841269841269841e-04f);
        ps = fma_(ps, z, 8.3333333333333332e-03f);
        ps = fma_(ps, z, -1.6666666666666666e-01f);
        float pc = -2.7557319223985888e-07f;
        pc = fma_(pc, z, 2.4801587301587302e-05f);
        pc = fma_(pc, z, -1.3888888888888889e-03f);
        pc = fma_(pc, z, 4.1666666666666664e-02f);
        pc = fma_(pc, z, -0.5f);
        *s = fma_(x * z, ps, x);
        *c = fma_(z, pc, 1.0f);
        return;
    }
#if defined(__CUDA_ARCH__)
    sincosf(x, s, c);
#else
    *s = sinf(x); *c = cosf(x);
#endif
}

// fast sin/cos for quantities whose error budget is >= 1e-5 absolute (rotor drag phase)
QS_HD void sincos_fast_(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    __sincosf(x, s, c);
#else
    *s = sinf(x); *c = cosf(x);
#endif
}

QS_HD float exp_(float x) {
#if defined(__CUDA_ARCH__)
    return __expf(x);                       // ex2.approx(x * log2e): <= 2 ulp + 2^-21 relative, far inside 1e-5
#else
    return expf(x);
#endif
}

QS_HD float rcp_(float x) {
#if defined(__CUDA_ARCH__)
    float r;                                // one MUFU.RCP (<= 1 ulp); only feeds the atan2 ratio
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// odd minimax polynomial of degree 15 for atan(t) on [-1, 1] (max error 3.1e-7 rad over float32 inputs, measured against
// float64)
// (R = float, or the two-lane f2 of qs_pack2.cuh: the same polynomial on both lanes with packed FP32 instructions)
template <class R>
QS_HD R atan_unit_(R t) {
    const R s = t * t;
    R p = splat_<R>(-0.004054448804439777f);
    p = fma_(p, s, 0.021862509027492236f);
    p = fma_(p, s, -0.05591164661595514f);
    p = fma_(p, s, 0.09642144994501779f);
    p = fma_(p, s, -0.13908608182486404f);
    p = fma_(p, s, 0.1994656129881479f);
    p = fma_(p, s, -0.3332986043366184f);
    p = fma_(p, s, 0.999999335547872f);
    return p * t;
}

// atan2 without branches or IEEE division: atan_unit_ on min/max, then octant fix-ups.
// atan2(0, 0) = 0 like libm; a NaN argument gives NaN.
QS_HD float atan2_(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 1e-30f ? mn * rcp_(mx) : 0.f;     // (also keeps flushed denormals away from MUFU.RCP)
    float r = atan_unit_(t);
    r = ay > ax ? 1.5707963267948966f - r : r;
    r = x < 0.f ? 3.141592653589793f - r : r;
    r = copysignf(r, y);
    const float chk = x + y;
    return chk != chk ? chk : r;
}

// asin on [-1, 1] through the half-angle identity asin(x) = 2 atan(x / (1 + sqrt(1 - x^2))): the atan argument never
// leaves [-1, 1], so the same polynomial serves, branch-free, with relative accuracy near 0 (max error 6e-7 rad).
QS_HD float asin_unit_(float x) {
#if defined(__CUDA_ARCH__)
    float c;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(fma_(-x, x, 1.0f)));
#else
    const float c = sqrtf(fma_(-x, x, 1.0f));
#endif
    return 2.0f * atan_unit_(x * rcp_(1.0f + c));
}

QS_HD float clamp_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

QS_HD bool finite_(float x) {
    // true for every non-NaN, non-Inf value; written on the bit pattern so that
    // fast-math style optimisations can never fold it away
    union { float f; uint32_t u; } v; v.f = x;
    return (v.u & 0x7f800000u) != 0x7f800000u;
}

}  // namespace qs

#include "qs_pack2.cuh"
