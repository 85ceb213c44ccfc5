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

QS_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);                       // MUFU.RSQ + 0 NR steps, <= 2 ulp
#else
    return 1.0f / sqrtf(x);
#endif
}

QS_HD float sqrt_(float x) { return sqrtf(x); }

// accurate sin/cos (state-carrying quantities)
QS_HD void sincos_(float x, float* s, float* c) {
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
    return expf(x);
#else
    return expf(x);
#endif
}

QS_HD float clamp_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

QS_HD bool finite_(float x) {
    // true for every non-NaN, non-Inf value; written on the bit pattern so that
    // fast-math style optimisations can never fold it away
    union { float f; uint32_t u; } v; v.f = x;
    return (v.u & 0x7f800000u) != 0x7f800000u;
}

}  // namespace qs
