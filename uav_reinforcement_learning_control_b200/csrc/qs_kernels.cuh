// qs_kernels.cuh -- plane load/store helpers and the per-launch env kernels (sm_100a).
//
// HBM layout: planar SoA float32 state[QS_NPLANES][num_envs] (include/quadsim_abi.h).  One
// thread owns one env; every plane access is a fully coalesced 128 B line per warp.  A kernel
// only touches the planes its MODE / flags need, so the algorithmic traffic of a hover step is
// 27 words in + 27 words out + action 16 B + obs 48 B + reward/done 8 B = 288 B per env-step.
#pragma once

#include "qs_env.cuh"

namespace qs {

constexpr int kBlock = 128;

// Block-cooperative, fully coalesced store of per-thread rows obs_local[D] to out[n][D]:
// rows are staged in shared memory (row stride D+1 when D is even to avoid bank conflicts)
// and written back as one contiguous span per block.
template <int D>
__device__ __forceinline__ void store_rows(float* __restrict__ out, int n, int block_first, const float* row,
                                           bool valid, float* smem) {
    constexpr int S = (D % 2 == 0) ? D + 1 : D;
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < D; ++k) smem[t * S + k] = row[k];
    __syncthreads();
    const int rows = min(kBlock, n - block_first);
    const int total = rows * D;
    float* base = out + (size_t)block_first * D;
    for (int idx = t; idx < total; idx += kBlock) {
        const int r = idx / D, c = idx - r * D;
        base[idx] = smem[r * S + c];
    }
    (void)valid;
    __syncthreads();
}

// ------------------------------------------------------------------------------ step
template <int MODE>
__global__ void __launch_bounds__(kBlock)
step_kernel(const __grid_constant__ QsParams P, Tables T, int n, float* __restrict__ state,
            const float4* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
            float* __restrict__ done, float* __restrict__ trunc, float* __restrict__ metrics,
            float* __restrict__ term_obs, const float* __restrict__ first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    __shared__ float sm[kBlock * (D + 1)];
    const int block_first = blockIdx.x * kBlock;
    const int i = block_first + threadIdx.x;
    const bool valid = i < n;
    float o_[D];
    if (valid) {
        Env e;
        load_env<MODE>(P, state, n, i, e);
        const float4 a4 = action[i];
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        StepOut so;
        float tobs[D];
        env_step<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, a, o_, term_obs ? tobs : nullptr,
                       first ? first + i : nullptr, n, so);
        store_env<MODE>(P, state, n, i, e);
        reward[i] = so.reward;
        done[i] = so.done;
        if (trunc) trunc[i] = so.truncated;
        if (metrics) {
            metrics[i] = so.pos_error; metrics[(size_t)n + i] = so.reward_hover;
            metrics[2 * (size_t)n + i] = so.reward_action; metrics[3 * (size_t)n + i] = so.reward;
        }
        if (term_obs && so.finished) {
#pragma unroll
            for (int k = 0; k < D; ++k) term_obs[(size_t)i * D + k] = tobs[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) o_[k] = 0.f;
    }
    store_rows<D>(obs, n, block_first, o_, valid, sm);
}

// ------------------------------------------------------------------------------ reset
template <int MODE>
__global__ void __launch_bounds__(kBlock)
reset_kernel(const __grid_constant__ QsParams P, Tables T, int n, float* __restrict__ state,
             const uint8_t* __restrict__ mask, float* __restrict__ obs, float* __restrict__ first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    Env e;
    load_env<MODE>(P, state, n, i, e, true);      // keeps episode / lifetime counters
    reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)i, e);
    store_env<MODE>(P, state, n, i, e);
    if (first) {
        const float x[21] = {e.b.p[0], e.b.p[1], e.b.p[2], e.b.q[0], e.b.q[1], e.b.q[2], e.b.q[3],
                             e.b.th[0], e.b.th[1], e.b.th[2], e.b.th[3], e.b.v[0], e.b.v[1], e.b.v[2],
                             e.b.w[0], e.b.w[1], e.b.w[2], e.b.s[0], e.b.s[1], e.b.s[2], e.b.s[3]};
#pragma unroll
        for (int k = 0; k < 21; ++k) first[(size_t)k * n + i] = x[k];
    }
    if (obs) {
        float rpy[3] = {0.f, 0.f, 0.f};
        if constexpr (ModeTraits<MODE>::kGym) quat_to_rpy(e.b.q, rpy);
        float o_[D];
        compute_obs<MODE>(P, e, rpy, o_);
#pragma unroll
        for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = o_[k];
    }
}

// ------------------------------------------------------------------------------ observe
template <int MODE>
__global__ void __launch_bounds__(kBlock)
observe_kernel(const __grid_constant__ QsParams P, Tables T, int n, const float* __restrict__ state,
               const float4* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
               float* __restrict__ done) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    Env e;
    load_env<MODE>(P, state, n, i, e);
    float rpy[3] = {0.f, 0.f, 0.f};
    if constexpr (ModeTraits<MODE>::kGym) quat_to_rpy(e.b.q, rpy);
    StepOut so;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (action) { const float4 a4 = action[i]; a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w; }
    evaluate<MODE>(P, T, e, action ? a : nullptr, rpy, so);
    if (obs) {
        float o_[D];
        compute_obs<MODE>(P, e, rpy, o_);
#pragma unroll
        for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = o_[k];
    }
    if (reward) reward[i] = so.reward;
    if (done) done[i] = so.done;
}

// ------------------------------------------------------------------------------ bare physics
__global__ void __launch_bounds__(kBlock)
physics_kernel(const __grid_constant__ QsParams P, int n, float* __restrict__ state, const float4* __restrict__ ctrl) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    Env e;
    load_env<QS_MODE_HOVER_BRAX>(P, state, n, i, e);     // planes 0..20 only
    const float4 c4 = ctrl[i];
    const float c[4] = {c4.x, c4.y, c4.z, c4.w};
    physics_step(P, e.b, c);
    store_env<QS_MODE_HOVER_BRAX>(P, state, n, i, e);
}

// ------------------------------------------------------------------------------ resident dyn-only rollout
// T env steps per launch with state in registers; actions U(-1,1)^4 from Philox stream 2 keyed by
// (global env id, global step index).  stats[4][n] += (sum reward, episodes finished, obs checksum, steps).
template <int MODE>
__global__ void __launch_bounds__(kBlock)
rollout_random_kernel(const __grid_constant__ QsParams P, Tables T, int n, float* __restrict__ state, int steps,
                      uint32_t t0, float* __restrict__ stats, const float* __restrict__ first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    Env e;
    load_env<MODE>(P, state, n, i, e);
    const uint32_t gid = P.env_id_offset + (uint32_t)i;
    float sum_r = 0.f, fin = 0.f, chk = 0.f;
    for (int t = 0; t < steps; ++t) {
        const U4 r = philox4x32_10(U4{gid, t0 + (uint32_t)t, 0u, STREAM_ACTION}, P.seed_lo, P.seed_hi);
        const float a[4] = {uniform_(r.x, -1.f, 1.f), uniform_(r.y, -1.f, 1.f), uniform_(r.z, -1.f, 1.f),
                            uniform_(r.w, -1.f, 1.f)};
        float o_[D];
        StepOut so;
        env_step<MODE>(P, T, gid, e, a, o_, nullptr, first ? first + i : nullptr, n, so);
        sum_r += so.reward;
        fin += so.finished ? 1.f : 0.f;
        float c = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) c += o_[k];
        chk += c;
    }
    store_env<MODE>(P, state, n, i, e);
    if (stats) {
        stats[i] += sum_r; stats[(size_t)n + i] += fin; stats[2 * (size_t)n + i] += chk;
        stats[3 * (size_t)n + i] += (float)steps;
    }
}

}  // namespace qs
