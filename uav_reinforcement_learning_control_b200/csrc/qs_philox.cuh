// qs_philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11).
//
// Replaces the reference's reset RNGs (jax.random threefry, train_brax_ppo.py:259-261;
// NumPy PCG64 via Gymnasium's np_random, envs/hover_env.py:220,227) as the north star asks:
// a reset is a pure function of (seed, global env id, episode index), so shards on
// different GPUs reproduce the single-GPU run.  oracle/philox.py is the NumPy statement
// of exactly this arithmetic; tests compare raw uint32 draws bit for bit.
#pragma once

#include "qs_math.cuh"

namespace qs {

enum : uint32_t { STREAM_RESET = 0u, STREAM_POLICY = 1u, STREAM_ACTION = 2u };

struct U4 { uint32_t x, y, z, w; };

QS_HD void mulhilo_(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p; *hi = (uint32_t)(p >> 32);
#endif
}

QS_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo_(0xD2511F53u, c.x, &hi0, &lo0);
        mulhilo_(0xCD9E8D57u, c.z, &hi1, &lo1);
        U4 n;
        n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

// the same with the ten round keys precomputed (QsParams.philox_key): two IMAD.WIDE + two LOP3 per round, the keys
// come straight from the constant bank
QS_HD U4 philox4x32_10(U4 c, const uint32_t* __restrict__ key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo_(0xD2511F53u, c.x, &hi0, &lo0);
        mulhilo_(0xCD9E8D57u, c.z, &hi1, &lo1);
        U4 n;
        n.x = hi1 ^ c.y ^ key[2 * r]; n.y = lo1; n.z = hi0 ^ c.w ^ key[2 * r + 1]; n.w = lo0;
        c = n;
    }
    return c;
}

// 24-bit uniform in [0, 1)
QS_HD float u01_(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

// fl32(fl32(u * (hi - lo)) + lo): two separately rounded ops so NumPy reproduces it exactly
QS_HD float uniform_(uint32_t x, float lo, float hi) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(__fmul_rn(u01_(x), hi - lo), lo);
#else
    volatile float t = u01_(x) * (hi - lo);
    return t + lo;
#endif
}

}  // namespace qs
