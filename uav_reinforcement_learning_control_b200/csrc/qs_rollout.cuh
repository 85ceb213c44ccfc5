// qs_rollout.cuh -- state-resident policy rollout (env step + 2x128 ReLU actor-critic) and GAE.
//
// Replaces the hot loops around the env: brax acting.generate_unroll inside ppo_train.train
// (train_brax_ppo.py:589-620: policy(obs) -> env.step, scanned unroll_length times) and SB3
// PPO.collect_rollouts (train.py:133-137 with MlpPolicy net_arch [128,128] ReLU, train.py:61-64),
// plus the GAE scans (brax compute_gae; SB3 RolloutBuffer.compute_returns_and_advantage).
//
// Kernel shape (FMA path): one CTA owns a tile of E envs for all T steps.  Env state lives in the
// registers of the tile's first E threads; the whole actor-critic (fp32, ~150 KB) lives in shared
// memory for the lifetime of the launch; activations ping-pong between two [128][E+4] shared
// buffers.  Each layer is a register-tiled [E x K] x [K x 128] product: a thread owns 4 envs x 8
// columns, a warp 16 envs x 64 columns, operands arrive as conflict-free LDS.128 broadcasts
// (3 wavefronts per 32 FFMA per thread), so the layer is FP32-pipe bound, not smem bound.
// Trajectories are written once, time-major, with coalesced stores (float4 for actions).
#pragma once

#include <cuda_runtime.h>
#include "qs_kernels.cuh"

namespace qs {

constexpr int kH = 128;          // hidden width (train.py:61-64, train_brax_ppo.py:453-455)
constexpr int kA = 4;            // action dim

struct RolloutBuffers {
    float* last_obs;             // [B][D]  (optional)
    float* obs; float* act; float* logp; float* value; float* reward; float* done; float* trunc;
    float* last_value;           // [B]
};

struct PolicyLayout {            // offsets (in floats) into the packed parameter vector
    int aW1, ab1, aW2, ab2, aW3, ab3, cW1, cb1, cW2, cb2, cW3, cb3, log_std, mean, inv_std, total;
};

__host__ __device__ inline PolicyLayout policy_layout(int D, int dist) {
    PolicyLayout L;
    const int Ao = dist == 1 ? 2 * kA : kA;
    int o = 0;
    L.aW1 = o; o += D * kH; L.ab1 = o; o += kH; L.aW2 = o; o += kH * kH; L.ab2 = o; o += kH;
    L.aW3 = o; o += kH * Ao; L.ab3 = o; o += Ao;
    L.cW1 = o; o += D * kH; L.cb1 = o; o += kH; L.cW2 = o; o += kH * kH; L.cb2 = o; o += kH;
    L.cW3 = o; o += kH; L.cb3 = o; o += 1;
    L.log_std = o; if (dist == 0) o += kA;
    L.mean = o; o += D; L.inv_std = o; o += D;
    L.total = o;
    return L;
}

inline int policy_param_count(const QsPolicyDesc& d) { return policy_layout(d.obs_dim, d.dist).total; }

#ifndef QS_FMA_PACKED
#define QS_FMA_PACKED 1
#endif
__device__ __forceinline__ unsigned long long pack2_(float lo, float hi) {
    unsigned long long p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
    return p;
}
__device__ __forceinline__ void unpack2_(unsigned long long p, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ unsigned long long fma2_(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ------------------------------------------------------------------------------------------
// [E x K] x [K x 128] register-tiled layer.  in: sIn[K][ES] (transposed activations), W[K][128],
// out: sOut[128][ES] = relu(in W + b).  ES = E + 4 (16 B aligned rows, conflict-light stores).
// Thread tile: envs e0..e0+3, columns {c0..c0+3} U {c0+32..c0+35}.
// ------------------------------------------------------------------------------------------
template <int E>
__device__ __forceinline__ void dense_relu(const float* __restrict__ sIn, int K, const float* __restrict__ W,
                                           const float* __restrict__ bias, float* __restrict__ sOut) {
    constexpr int ES = E + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarpsE = E / 16;                       // warps along the env axis
    const int we = warp % kWarpsE, wn = warp / kWarpsE;   // wn in {0,1}: 64-column half
    const int e0 = we * 16 + (lane & 3) * 4;
    const int c0 = wn * 64 + (lane >> 2) * 4;
#if QS_FMA_PACKED && defined(__CUDA_ARCH__)
    // Blackwell packed fp32 (fma.rn.f32x2, SASS FFMA2): one instruction = two FMAs on adjacent output columns.  The
    // weight pairs come packed out of the 128-bit shared loads; only the activation is duplicated into both halves.
    unsigned long long acc2[4][4];
    {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c0);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + c0 + 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc2[i][0] = pack2_(b0.x, b0.y); acc2[i][1] = pack2_(b0.z, b0.w);
            acc2[i][2] = pack2_(b1.x, b1.y); acc2[i][3] = pack2_(b1.z, b1.w);
        }
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(sIn + k * ES + e0);
        const float4 w0 = *reinterpret_cast<const float4*>(W + k * kH + c0);
        const float4 w1 = *reinterpret_cast<const float4*>(W + k * kH + c0 + 32);
        const unsigned long long wp[4] = {pack2_(w0.x, w0.y), pack2_(w0.z, w0.w), pack2_(w1.x, w1.y), pack2_(w1.z, w1.w)};
        const unsigned long long ap[4] = {pack2_(a.x, a.x), pack2_(a.y, a.y), pack2_(a.z, a.z), pack2_(a.w, a.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = fma2_(ap[i], wp[j], acc2[i][j]);
    }
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack2_(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
#else
    float acc[4][8];
    {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c0);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + c0 + 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
            acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
        }
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(sIn + k * ES + e0);
        const float4 w0 = *reinterpret_cast<const float4*>(W + k * kH + c0);
        const float4 w1 = *reinterpret_cast<const float4*>(W + k * kH + c0 + 32);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + (j & 3) + (j >> 2) * 32;
        float4 o;
        o.x = fmaxf(acc[0][j], 0.f); o.y = fmaxf(acc[1][j], 0.f);
        o.z = fmaxf(acc[2][j], 0.f); o.w = fmaxf(acc[3][j], 0.f);
        *reinterpret_cast<float4*>(sOut + c * ES + e0) = o;
    }
}

// small head: out[e][j] = sum_k sIn[k][e] W[k][NO + j] + b[j], j < NO (NO in {1,4,8})
template <int E, int NO>
__device__ __forceinline__ void head(const float* __restrict__ sIn, const float* __restrict__ W,
                                     const float* __restrict__ bias, float* __restrict__ sOut /*[E][NO]*/,
                                     float* __restrict__ sPart /*[NT/E][E][NO]*/, int nthreads) {
    constexpr int ES = E + 4;
    const int e = threadIdx.x % E, part = threadIdx.x / E;
    const int nparts = nthreads / E;
    const int kper = kH / nparts;
    float acc[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) acc[j] = 0.f;
    for (int k = part * kper; k < (part + 1) * kper; ++k) {
        const float h = sIn[k * ES + e];
#pragma unroll
        for (int j = 0; j < NO; ++j) acc[j] = fmaf(h, W[k * NO + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NO; ++j) sPart[(part * E + e) * NO + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < E) {
#pragma unroll
        for (int j = 0; j < NO; ++j) {
            float s = bias[j];
            for (int p = 0; p < nparts; ++p) s += sPart[(p * E + e) * NO + j];
            sOut[e * NO + j] = s;
        }
    }
}

QS_HD float softplus_(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// ------------------------------------------------------------------------------------------
// rollout kernel
// ------------------------------------------------------------------------------------------
template <int MODE, int DIST, int E>
__global__ void __launch_bounds__(E * 4)
rollout_policy_kernel(const __grid_constant__ QsParams P, Tables T, int n, float* __restrict__ state,
                      const float* __restrict__ params, int steps, uint32_t t0, int deterministic,
                      float bootstrap_gamma, RolloutBuffers rb, const float* __restrict__ first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    constexpr int ES = E + 4;
    constexpr int NT = E * 4;
    constexpr int Ao = DIST == 1 ? 2 * kA : kA;
    extern __shared__ __align__(16) float smem[];
    const PolicyLayout L = policy_layout(D, DIST);
    const int wfloats = (L.total + 3) & ~3;
    float* sW = smem;                        // packed policy parameters
    float* sX = sW + wfloats;                // [D][ES]  normalised obs, transposed
    float* sH1 = sX + ((D * ES + 3) & ~3);   // [128][ES]
    float* sH2 = sH1 + kH * ES;              // [128][ES]
    float* sObs = sH2 + kH * ES;             // [E][D]   raw obs rows (trajectory + last_obs)
    float* sOut = sObs + ((E * D + 3) & ~3); // [E][Ao]
    float* sVal = sOut + E * Ao;             // [E]
    float* sPart = sVal + E;                 // [4][E][Ao]

    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * E;
    const int rows = min(E, n - b0);
    for (int idx = tid; idx < L.total; idx += NT) sW[idx] = params[idx];

    const bool owner = tid < E && (b0 + tid) < n;
    const uint32_t gid = P.env_id_offset + (uint32_t)(b0 + tid);
    Env e;
    float obs_[D];
    if (owner) {
        load_env<MODE>(P, state, n, b0 + tid, e);
        float rpy[3] = {0.f, 0.f, 0.f};
        if constexpr (ModeTraits<MODE>::kGym) quat_to_rpy(e.b.q, rpy);
        compute_obs<MODE>(P, e, rpy, obs_);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) obs_[k] = 0.f;
    }
    // Speculative resets (gym modes with Philox re-sampling).  During the env step only the first E threads of the CTA
    // have work, so threads E..2E-1 compute, for env (tid - E), the state it WOULD be reset to for its next episode
    // -- all lanes busy, off the owners' critical path -- into the idle sH1 buffer; a finished env picks its candidate
    // up after the barrier that follows the env step.  The partner tracks the env's episode counter itself and learns
    // of a reset through a flag word next to the candidates.
    constexpr bool kGymMode = ModeTraits<MODE>::kGym;
    const bool spec_reset = kGymMode && P.auto_reset == QS_RESET_RESAMPLE && !P.waypoint_mode;    // CTA-uniform
    const bool partner = spec_reset && tid >= E && tid < 2 * E && (b0 + tid - E) < n;
    constexpr int kCandF = 28;                                   // p3 q4 v3 w3 target3 obs12
    float* sCand = sH1;                                          // [E][kCandF]
    uint32_t* sFlag = reinterpret_cast<uint32_t*>(sH1 + E * kCandF);   // [E]
    uint32_t p_epi = 0u;
    if (partner) p_epi = f2u_(state[26 * (size_t)n + b0 + tid - E]);
    __syncthreads();

    auto publish_obs = [&](const float* o) {
        if (tid < E) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                sObs[tid * D + k] = o[k];
                sX[k * ES + tid] = (o[k] - sW[L.mean + k]) * sW[L.inv_std + k];
            }
        }
    };
    auto critic = [&]() {   // sX -> sVal ; leaves with a barrier
        dense_relu<E>(sX, D, sW + L.cW1, sW + L.cb1, sH1);
        __syncthreads();
        dense_relu<E>(sH1, kH, sW + L.cW2, sW + L.cb2, sH2);
        __syncthreads();
        head<E, 1>(sH2, sW + L.cW3, sW + L.cb3, sVal, sPart, NT);
        __syncthreads();
    };

    for (int t = 0; t < steps; ++t) {
        publish_obs(obs_);
        __syncthreads();
        // trajectory: raw observations, one contiguous span per tile
        if (rb.obs) {
            float* dst = rb.obs + ((size_t)t * n + b0) * D;
            for (int idx = tid; idx < rows * D; idx += NT) dst[idx] = sObs[idx];
        }
        // actor
        dense_relu<E>(sX, D, sW + L.aW1, sW + L.ab1, sH1);
        __syncthreads();
        dense_relu<E>(sH1, kH, sW + L.aW2, sW + L.ab2, sH2);
        __syncthreads();
        head<E, Ao>(sH2, sW + L.aW3, sW + L.ab3, sOut, sPart, NT);
        __syncthreads();
        // critic
        critic();

        StepOut so;
        so.reward = 0.f; so.done = 0.f; so.truncated = 0.f; so.finished = false;
        float tobs[D];
        bool need_boot = false;
        if (owner) {
            const U4 r = philox4x32_10(U4{gid, t0 + (uint32_t)t, 0u, STREAM_POLICY}, P.philox_key);
            // Box-Muller: two pairs
            float eps[4];
            {
                const float u0 = ((float)(r.x >> 8) + 1.0f) * 5.9604644775390625e-8f;
                const float u1 = (float)(r.y >> 8) * 5.9604644775390625e-8f;
                const float u2 = ((float)(r.z >> 8) + 1.0f) * 5.9604644775390625e-8f;
                const float u3 = (float)(r.w >> 8) * 5.9604644775390625e-8f;
                // fast intrinsics: the noise only has to be N(0,1) to ~1e-6, it is not a state variable
                const float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
                float s0, c0, s1, c1;
                __sincosf(6.283185307179586f * u1, &s0, &c0);
                __sincosf(6.283185307179586f * u3, &s1, &c1);
                eps[0] = r0 * c0; eps[1] = r0 * s0; eps[2] = r1 * c1; eps[3] = r1 * s1;
            }
            float raw[4], act[4], logp = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (DIST == 0) {
                    // SB3 DiagGaussianDistribution: a ~ N(mean, exp(log_std)); clipped only for the env
                    const float mean = sOut[tid * Ao + j];
                    const float ls = sW[L.log_std + j];
                    const float z = deterministic ? 0.f : eps[j];
                    raw[j] = fmaf(expf(ls), z, mean);
                    logp += -0.5f * z * z - ls - 0.9189385332046727f;
                    act[j] = clamp_(raw[j], -1.0f, 1.0f);
                } else {
                    // brax NormalTanhDistribution(min_std = 0.001)
                    const float loc = sOut[tid * Ao + j];
                    const float scale = softplus_(sOut[tid * Ao + kA + j]) + 0.001f;
                    const float z = deterministic ? 0.f : eps[j];
                    raw[j] = fmaf(scale, z, loc);
                    const float ldj = 2.0f * (0.6931471805599453f - raw[j] - softplus_(-2.0f * raw[j]));
                    logp += -0.5f * z * z - logf(scale) - 0.9189385332046727f - ldj;
                    act[j] = tanhf(raw[j]);
                }
            }
            const size_t o = (size_t)t * n + b0 + tid;
            if (rb.act) reinterpret_cast<float4*>(rb.act)[o] = make_float4(raw[0], raw[1], raw[2], raw[3]);
            if (rb.logp) rb.logp[o] = logp;
            if (rb.value) rb.value[o] = sVal[tid];
            so.needs_reset = false;
            env_step<MODE, kGymMode>(P, T, gid, e, act, obs_, tobs, first ? first + b0 + tid : nullptr, n, so);
            need_boot = bootstrap_gamma > 0.f && so.finished && so.truncated != 0.f && so.done == 0.f;
            if constexpr (ModeTraits<MODE>::kBrax) need_boot = bootstrap_gamma > 0.f && so.truncated != 0.f;
            if constexpr (kGymMode) {
                if (spec_reset) {
                    sFlag[tid] = so.needs_reset ? 1u : 0u;
                } else if (so.needs_reset) {                  // waypoint resets draw nothing: inline
                    float rpy[3];
                    reset_env<MODE>(P, T, gid, e, rpy);
                    compute_obs<MODE>(P, e, rpy, obs_);
                }
            }
        } else if (partner) {
            if constexpr (kGymMode) {
                Env r;
                r.episode = p_epi + 1u;
                r.wp_idx = 0; r.wp_reached = 0; r.laps = 0;
                float rpy[3], oc[D];
                reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)(b0 + tid - E), r, rpy);
                compute_obs<MODE>(P, r, rpy, oc);
                float4* d = reinterpret_cast<float4*>(sCand + (tid - E) * kCandF);
                d[0] = make_float4(r.b.p[0], r.b.p[1], r.b.p[2], r.b.q[0]);
                d[1] = make_float4(r.b.q[1], r.b.q[2], r.b.q[3], r.b.v[0]);
                d[2] = make_float4(r.b.v[1], r.b.v[2], r.b.w[0], r.b.w[1]);
                d[3] = make_float4(r.b.w[2], r.target[0], r.target[1], r.target[2]);
                d[4] = make_float4(oc[0], oc[1], oc[2], oc[3]);
                d[5] = make_float4(oc[4], oc[5], oc[6], oc[7]);
                d[6] = make_float4(oc[8], oc[9], oc[10], oc[11]);
            }
        }
        const int boot_any = __syncthreads_or(need_boot ? 1 : 0);
        if constexpr (kGymMode) {
            if (spec_reset) {
                if (owner && so.needs_reset) {                // e.episode was advanced by env_step
                    const float4* d = reinterpret_cast<const float4*>(sCand + tid * kCandF);
                    const float4 c0 = d[0], c1 = d[1], c2 = d[2], c3 = d[3], c4 = d[4], c5 = d[5], c6 = d[6];
                    e.b.p[0] = c0.x; e.b.p[1] = c0.y; e.b.p[2] = c0.z; e.b.q[0] = c0.w;
                    e.b.q[1] = c1.x; e.b.q[2] = c1.y; e.b.q[3] = c1.z; e.b.v[0] = c1.w;
                    e.b.v[1] = c2.x; e.b.v[2] = c2.y; e.b.w[0] = c2.z; e.b.w[1] = c2.w;
                    e.b.w[2] = c3.x; e.target[0] = c3.y; e.target[1] = c3.z; e.target[2] = c3.w;
                    obs_[0] = c4.x; obs_[1] = c4.y; obs_[2] = c4.z; obs_[3] = c4.w;
                    obs_[4] = c5.x; obs_[5] = c5.y; obs_[6] = c5.z; obs_[7] = c5.w;
                    obs_[8] = c6.x; obs_[9] = c6.y; obs_[10] = c6.z; obs_[11] = c6.w;
#pragma unroll
                    for (int k = 0; k < 4; ++k) { e.b.th[k] = 0.f; e.b.s[k] = 0.f; e.prev_action[k] = 0.f; }
#pragma unroll
                    for (int k = 0; k < 3; ++k) e.rate_int[k] = 0.f;
                    e.step_count = 0; e.ep_steps = 0; e.done_prev = 0.f; e.voltage = P.v_nominal;
                }
                if (partner && sFlag[tid - E]) p_epi += 1u;
            }
        }
        // SB3 timeout bootstrap: reward += gamma * V(terminal_obs) for truncated-not-terminated episodes
        if (boot_any) {
            publish_obs(need_boot ? tobs : obs_);
            __syncthreads();
            critic();
            if (need_boot) so.reward = fmaf(bootstrap_gamma, sVal[tid], so.reward);
            __syncthreads();
        }
        if (owner) {
            const size_t o = (size_t)t * n + b0 + tid;
            if (rb.reward) rb.reward[o] = so.reward;
            if (rb.done) rb.done[o] = so.done;
            if (rb.trunc) rb.trunc[o] = so.truncated;
        }
    }

    // value of the final observation (GAE bootstrap) and final obs
    publish_obs(obs_);
    __syncthreads();
    if (rb.last_obs) {
        float* dst = rb.last_obs + (size_t)b0 * D;
        for (int idx = tid; idx < rows * D; idx += NT) dst[idx] = sObs[idx];
    }
    critic();
    if (owner) {
        if (rb.last_value) rb.last_value[b0 + tid] = sVal[tid];
        store_env<MODE>(P, state, n, b0 + tid, e);
    }
}

template <int MODE, int DIST, int E>
inline size_t rollout_smem_bytes() {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    constexpr int ES = E + 4;
    constexpr int Ao = DIST == 1 ? 2 * kA : kA;
    const PolicyLayout L = policy_layout(D, DIST);
    size_t f = (size_t)((L.total + 3) & ~3) + ((D * ES + 3) & ~3) + 2 * (size_t)kH * ES + ((E * D + 3) & ~3) +
               (size_t)E * Ao + E + 4 * (size_t)E * Ao;
    return f * sizeof(float);
}

struct RolloutOpts { int deterministic; float bootstrap_gamma; };

template <int MODE, int DIST, int E>
inline int launch_rollout_t(const QsParams& P, const Tables& T, int n, float* state, const float* params, int steps,
                            uint32_t t0, const RolloutOpts& opt, const RolloutBuffers& rb, const float* first,
                            cudaStream_t s) {
    auto kern = rollout_policy_kernel<MODE, DIST, E>;
    const size_t smem = rollout_smem_bytes<MODE, DIST, E>();
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
    kern<<<(n + E - 1) / E, E * 4, smem, s>>>(P, T, n, state, params, steps, t0, opt.deterministic,
                                              opt.bootstrap_gamma, rb, first);
    return 0;
}

inline int launch_rollout_policy(const QsParams& P, const Tables& T, int n, float* state, const QsPolicyDesc& d,
                                 const float* params, int steps, uint32_t t0, const RolloutBuffers& rb,
                                 const float* first, cudaStream_t s) {
    RolloutOpts opt{d.deterministic, d.bootstrap_gamma};
    // tile size: 64 envs per CTA unless that leaves most SMs idle
    const bool small = n <= 148 * 32;
#define QS_RL(MODE_, DIST_)                                                                                   \
    return small ? launch_rollout_t<MODE_, DIST_, 32>(P, T, n, state, params, steps, t0, opt, rb, first, s)   \
                 : launch_rollout_t<MODE_, DIST_, 64>(P, T, n, state, params, steps, t0, opt, rb, first, s)
    if (P.mode == QS_MODE_HOVER_GYM && d.dist == 0) { QS_RL(QS_MODE_HOVER_GYM, 0); }
    if (P.mode == QS_MODE_HOVER_GYM && d.dist == 1) { QS_RL(QS_MODE_HOVER_GYM, 1); }
    if (P.mode == QS_MODE_TRAJ_GYM && d.dist == 0) { QS_RL(QS_MODE_TRAJ_GYM, 0); }
#undef QS_RL
    // 21-D observations: the fp32 weights alone take 159 KB of shared memory, so only the 32-env tile
    // (2 x [128][36] activation buffers) fits in the 227 KB a CTA may use
    if (P.mode == QS_MODE_MJX_BRAX && d.dist == 1)
        return launch_rollout_t<QS_MODE_MJX_BRAX, 1, 32>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    if (P.mode == QS_MODE_MJX_BRAX && d.dist == 0)
        return launch_rollout_t<QS_MODE_MJX_BRAX, 0, 32>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    if (P.mode == QS_MODE_HOVER_BRAX && d.dist == 1)          // QuadHoverBraxEnv under the Brax trainer (train_brax_ppo.py:39-176)
        return launch_rollout_t<QS_MODE_HOVER_BRAX, 1, 32>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    return -100;
}

// ------------------------------------------------------------------------------------------
// GAE reverse-time scan, one thread per env, [T][B] time-major buffers (coalesced per step).
// ------------------------------------------------------------------------------------------
// The recurrence is sequential in t but its loads are not: each thread fetches kGaeBatch time steps at once (all
// loads in flight together), then runs the recurrence over them, so the HBM latency is paid once per batch.
constexpr int kGaeBatch = 8;
constexpr int kGaeBlock = 32;      // small CTAs: 8192 envs still spread over every SM

__global__ void __launch_bounds__(kGaeBlock)
gae_kernel(int T, int B, const float* __restrict__ reward, const float* __restrict__ value,
           const float* __restrict__ done, const float* __restrict__ trunc, const float* __restrict__ last_value,
           float gamma, float lam, int brax_form, float* __restrict__ adv, float* __restrict__ ret) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float next_v = last_value[b], vs_next = next_v, a = 0.f;
    for (int t1 = T; t1 > 0; t1 -= kGaeBatch) {
        // time steps t1-1, t1-2, ... (k = 0 is the latest); steps below 0 are masked out
        float r_[kGaeBatch], v_[kGaeBatch], d_[kGaeBatch], tr_[kGaeBatch];
#pragma unroll
        for (int k = 0; k < kGaeBatch; ++k) {
            const int t = t1 - 1 - k;
            const size_t o = (size_t)(t < 0 ? 0 : t) * B + b;
            r_[k] = __ldcs(reward + o); v_[k] = __ldcs(value + o); d_[k] = __ldcs(done + o);
            tr_[k] = trunc ? __ldcs(trunc + o) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kGaeBatch; ++k) {
            const int t = t1 - 1 - k;
            if (t < 0) break;
            const size_t o = (size_t)t * B + b;
            const float v = v_[k], r = r_[k];
            if (!brax_form) {
                // SB3 RolloutBuffer.compute_returns_and_advantage: an episode boundary after step t
                // (terminated or truncated) cuts both the bootstrap and the recursion
                const float nnt = 1.0f - fmaxf(d_[k], tr_[k]);
                const float delta = fmaf(gamma * next_v, nnt, r) - v;
                a = fmaf(gamma * lam * nnt, a, delta);
                __stcs(adv + o, a);
                __stcs(ret + o, a + v);
                next_v = v;
            } else {
                // brax.training.agents.ppo.losses.compute_gae with termination = done * (1 - truncation)
                const float tr = tr_[k];
                const float term = d_[k] * (1.0f - tr);
                const float mask = 1.0f - tr;
                const float disc = gamma * (1.0f - term);
                const float delta = (fmaf(disc, next_v, r) - v) * mask;
                a = fmaf(disc * mask * lam, a, delta);
                const float vs = a + v;
                __stcs(adv + o, (fmaf(disc, vs_next, r) - v) * mask);
                __stcs(ret + o, vs);
                next_v = v; vs_next = vs;
            }
        }
    }
}

}  // namespace qs
