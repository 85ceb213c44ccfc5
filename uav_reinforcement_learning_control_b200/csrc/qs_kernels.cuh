// qs_kernels.cuh -- plane load/store helpers and the per-launch env kernels (sm_100a).
//
// HBM layout: planar SoA float32 state[QS_NPLANES][num_envs] (include/quadsim_abi.h).  One
// thread owns one env; every plane access is a fully coalesced 128 B line per warp.  A kernel
// only touches the planes its MODE / flags need, so the algorithmic traffic of a hover step is
// 27 words in + 27 words out + action 16 B + obs 48 B + reward/done 8 B = 288 B per env-step.
#pragma once

#include "qs_env.cuh"

namespace qs {

#ifndef QS_STEP_BLOCK
#define QS_STEP_BLOCK 128
#endif
#ifndef QS_STEP_MIN_BLOCKS
#define QS_STEP_MIN_BLOCKS 7   /* <= 72 registers: 28 warps/SM; measured best on B200 (profiles/README.md) */
#endif
#ifndef QS_USE_PDL
#define QS_USE_PDL 1          /* step kernel: programmatic dependent launch (griddepcontrol) */
#endif
#ifndef QS_PREFETCH_AHEAD
#define QS_PREFETCH_AHEAD 518  /* step kernel: each CTA pulls the state planes / actions of the tile this many CTAs ahead into L2
                                   (148 SMs x 7 resident CTAs = 1036 per wave; half a wave measured best); 0 = off */
#endif
#ifndef QS_STREAM_HINTS
#define QS_STREAM_HINTS 1      /* step kernel: actions are read and obs / reward / done written with evict-first (.cs) policy so the
                                   state planes keep the L2 between steps */
#endif
constexpr int kBlock = QS_STEP_BLOCK;

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

// ------------------------------------------------------------------------------ warp-cooperative auto-reset
// Under random (or early-training) policies ~10 % of the envs finish an episode every step, so nearly every warp
// would run the ~570-instruction Philox reset path with 3 of 32 lanes active.  The shipped scheme is a
// warp-cooperative reset through a small per-warp shared scratch, no block barrier and no __fns (the block-level
// compaction, the shuffle-based warp variant and the inline reset it replaced are in the tuning log of
// profiles/README.md; their source is in the history up to commit 776b737).
// A finished lane with rank r (among the warp's finished lanes) posts (lane, episode) in slot r; lanes 4r..4r+3 each
// compute ONE Philox block of that reset and already map their four words to the target ranges (the bounds are
// indexed by the block id, a constant-bank load), post the four floats, and the owner picks up its 16 values with
// four 128-bit shared loads.  Only Euler -> quaternion and the observation run at low lane efficiency.  Up to 8
// resets per pass; more (rare: the mean is ~3.5 per warp under random policies) loop.  Same arithmetic as reset_env.
struct WarpResetScratch {
    uint32_t src[8];
    uint32_t epi[8];
    float4 val[32];
};

// The env of lane L has global id gid_warp_first + L * LANE_STRIDE (LANE_STRIDE = 2 in the packed step kernel, where a
// lane owns two adjacent envs and the routine runs once per half with gid_warp_first advanced by the half).
template <int MODE, int LANE_STRIDE = 1>
__device__ __forceinline__ void warp_autoreset_smem(const QsParams& P, uint32_t gid_warp_first, Env& e, float* obs,
                                                    bool need, WarpResetScratch& S) {
    static_assert(ModeTraits<MODE>::kGym, "Philox re-sampling exists in the gym modes only");
    const int lane = threadIdx.x & 31;
    unsigned pending = __ballot_sync(0xffffffffu, need);
    while (pending) {                                   // warp-uniform
        const int rank = __popc(pending & ((1u << lane) - 1u));
        const bool mine = ((pending >> lane) & 1u) && rank < 8;
        if (mine) { S.src[rank] = (uint32_t)lane; S.epi[rank] = e.episode; }
        __syncwarp();
        const int npass = min(__popc(pending), 8);
        const int r = lane >> 2, blk = lane & 3;
        constexpr int kBlocks = (MODE == QS_MODE_HOVER_GYM) ? 4 : 3;
        if (r < npass && blk < kBlocks) {
            const U4 rnd = philox4x32_10(U4{gid_warp_first + S.src[r] * (uint32_t)LANE_STRIDE, S.epi[r], (uint32_t)blk, STREAM_RESET}, P.philox_key);
            // words 0..11 -> state12 ranges, 12..14 -> target ranges: both live in one table of 16 (lo, hi) pairs
            const float* lo = blk < 3 ? &P.init_lo[4 * blk] : &P.target_lo[0];
            const float* hi = blk < 3 ? &P.init_hi[4 * blk] : &P.target_hi[0];
            float4 v;
            v.x = uniform_(rnd.x, lo[0], hi[0]);
            v.y = uniform_(rnd.y, lo[1], hi[1]);
            v.z = uniform_(rnd.z, lo[2], hi[2]);
            v.w = blk < 3 ? uniform_(rnd.w, lo[3], hi[3]) : 0.f;
            S.val[lane] = v;
        }
        __syncwarp();
        if (mine) {
            const float4 a = S.val[4 * rank], b = S.val[4 * rank + 1], c = S.val[4 * rank + 2];
            e.b.p[0] = a.x; e.b.p[1] = a.y; e.b.p[2] = a.z;
            float rpy[3] = {a.w, b.x, b.y};
            rpy_to_quat(rpy, e.b.q);
            e.b.v[0] = b.z; e.b.v[1] = b.w; e.b.v[2] = c.x;
            e.b.w[0] = c.y; e.b.w[1] = c.z; e.b.w[2] = c.w;
            if constexpr (MODE == QS_MODE_HOVER_GYM) {
                const float4 d = S.val[4 * rank + 3];
                e.target[0] = d.x; e.target[1] = d.y; e.target[2] = d.z;
            } else {
                e.target[0] = e.b.p[0]; e.target[1] = e.b.p[1]; e.target[2] = e.b.p[2];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { e.b.th[k] = 0.f; e.b.s[k] = 0.f; e.prev_action[k] = 0.f; }
#pragma unroll
            for (int k = 0; k < 3; ++k) e.rate_int[k] = 0.f;
            e.step_count = 0; e.ep_steps = 0; e.done_prev = 0.f; e.voltage = P.v_nominal;
            compute_obs<MODE>(P, e, rpy, obs);
        }
        pending &= ~__ballot_sync(0xffffffffu, mine);   // also orders this pass's shared reads before the next pass's writes
    }
}

// whether this launch re-samples finished envs through the warp-cooperative path (block-uniform)
template <int MODE>
__device__ __forceinline__ bool use_compaction(const QsParams& P) {
    return ModeTraits<MODE>::kGym && P.auto_reset == QS_RESET_RESAMPLE && !P.waypoint_mode;
}

// 12-float observation row -> three 128-bit stores
__device__ __forceinline__ void store_obs12(float* __restrict__ obs, int i, const float* o) {
    float4* d = reinterpret_cast<float4*>(obs + (size_t)i * 12);
#if QS_STREAM_HINTS
    __stcs(d, make_float4(o[0], o[1], o[2], o[3]));
    __stcs(d + 1, make_float4(o[4], o[5], o[6], o[7]));
    __stcs(d + 2, make_float4(o[8], o[9], o[10], o[11]));
#else
    d[0] = make_float4(o[0], o[1], o[2], o[3]);
    d[1] = make_float4(o[4], o[5], o[6], o[7]);
    d[2] = make_float4(o[8], o[9], o[10], o[11]);
#endif
}

// ------------------------------------------------------------------------------ step
// F = FeatLean: the plain configuration (see qs_env.cuh) with metrics / terminal_obs not requested
template <int MODE, class F = FeatAll>
__global__ void __launch_bounds__(kBlock, QS_STEP_MIN_BLOCKS)
step_kernel(const __grid_constant__ QsParams P, Tables T, int n, int lo, int count, float* __restrict__ state,
            const float4* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
            float* __restrict__ done, float* __restrict__ trunc, float* __restrict__ metrics,
            float* __restrict__ term_obs, const float* __restrict__ first) {
    // n = plane stride (all envs of the handle); this launch steps envs [lo, lo + count)
    constexpr int D = ModeTraits<MODE>::kObsDim;
    constexpr bool kGym = ModeTraits<MODE>::kGym;
    constexpr bool kLean = !F::kWaypoint;           // lean launches never carry metrics / terminal_obs
    // gym modes: reset scratch; brax modes: staging tile for the 21-float observation rows
    constexpr size_t kGymScratch = sizeof(WarpResetScratch) * (kBlock / 32);
    __shared__ __align__(16) unsigned char smem_raw[kGym ? kGymScratch : sizeof(float) * kBlock * (D + 1)];
    const int block_first = lo + blockIdx.x * kBlock;
    const int i = block_first + threadIdx.x;
    const bool valid = i < lo + count;
#if QS_USE_PDL
    // programmatic dependent launch (quadsim.cu: launch_step): let the next kernel on the stream start its ramp-up
    // now, and do not touch global memory before everything launched ahead of us has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
#if QS_PREFETCH_AHEAD > 0
    if constexpr (kGym) {
        // L2 prefetch of a tile one wave ahead: 27 planes x 4 lines + 16 action lines, one line per thread and pass.
        // It decouples the DRAM reads from the few warps that are in their load phase at any moment.
        const int pf_first = block_first + QS_PREFETCH_AHEAD * kBlock;
        if (pf_first + kBlock <= lo + count) {
            constexpr int kLines = kBlock * 4 / 128;                 // 128-byte lines per plane per tile
            for (int l = threadIdx.x; l < 27 * kLines; l += kBlock) {
                const int p = l / kLines, c = l - p * kLines;
                if (p != 25 || (F::kBattery && P.battery))
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(state + (size_t)p * n + pf_first + c * 32));
            }
            if (threadIdx.x < kBlock / 8)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(action + pf_first + threadIdx.x * 8));
        }
    }
#endif
    float o_[D];
    Env e;
    StepOut so;
    so.needs_reset = false;
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        load_env<MODE, F>(P, state, n, i, e);
#if QS_STREAM_HINTS
        a4 = __ldcs(action + i);
#else
        a4 = action[i];
#endif
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float tobs[D];
        env_step<MODE, kGym, F>(P, T, P.env_id_offset + (uint32_t)i, e, a, o_, (!kLean && term_obs) ? tobs : nullptr,
                                first ? first + i : nullptr, n, so);
#if QS_STREAM_HINTS
        __stcs(reward + i, so.reward);
        __stcs(done + i, so.done);
        if (trunc) __stcs(trunc + i, so.truncated);
#else
        reward[i] = so.reward;
        done[i] = so.done;
        if (trunc) trunc[i] = so.truncated;
#endif
        if (!kLean && metrics) {
            metrics[i] = so.pos_error; metrics[(size_t)n + i] = so.reward_hover;
            metrics[2 * (size_t)n + i] = so.reward_action; metrics[3 * (size_t)n + i] = so.reward;
        }
        if (!kLean && term_obs && so.finished) {
#pragma unroll
            for (int k = 0; k < D; ++k) term_obs[(size_t)i * D + k] = tobs[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) o_[k] = 0.f;
    }
    if constexpr (kGym) {
        if (P.auto_reset == QS_RESET_RESAMPLE) {
            if (F::kWaypoint && P.waypoint_mode) {
                // waypoint resets draw nothing: do them inline
                if (so.needs_reset) {
                    float rpy[3];
                    reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, rpy);
                    compute_obs<MODE>(P, e, rpy, o_);
                }
            } else {
                warp_autoreset_smem<MODE>(P, P.env_id_offset + (uint32_t)(i - (int)(threadIdx.x & 31)), e, o_, so.needs_reset,
                                          reinterpret_cast<WarpResetScratch*>(smem_raw)[threadIdx.x >> 5]);
            }
        }
        // one store sequence for all lanes (splitting it by needs_reset makes nearly every warp run it twice)
        if (valid) store_env<MODE, F>(P, state, n, i, e);
        if (valid) store_obs12(obs, i, o_);          // per-thread 128-bit stores (3 % faster than a coalescing smem tile: two barriers fewer)
    } else {
        if (valid) store_env<MODE, F>(P, state, n, i, e);
        store_rows<D>(obs, lo + count, block_first, o_, valid, reinterpret_cast<float*>(smem_raw));
    }
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
    float rpy[3];
    reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, rpy);
    store_env<MODE>(P, state, n, i, e);
    if (first) {
        const float x[21] = {e.b.p[0], e.b.p[1], e.b.p[2], e.b.q[0], e.b.q[1], e.b.q[2], e.b.q[3],
                             e.b.th[0], e.b.th[1], e.b.th[2], e.b.th[3], e.b.v[0], e.b.v[1], e.b.v[2],
                             e.b.w[0], e.b.w[1], e.b.w[2], e.b.s[0], e.b.s[1], e.b.s[2], e.b.s[3]};
#pragma unroll
        for (int k = 0; k < 21; ++k) first[(size_t)k * n + i] = x[k];
    }
    if (obs) {
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
__global__ void __launch_bounds__(kBlock, QS_STEP_MIN_BLOCKS)
rollout_random_kernel(const __grid_constant__ QsParams P, Tables T, int n, float* __restrict__ state, int steps,
                      uint32_t t0, float* __restrict__ stats, const float* __restrict__ first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    constexpr bool kGym = ModeTraits<MODE>::kGym;
    __shared__ __align__(16) unsigned char smem_raw[kGym ? sizeof(WarpResetScratch) * (kBlock / 32) : 16];
    const int block_first = blockIdx.x * kBlock;
    const int i = block_first + threadIdx.x;
    const bool valid = i < n;
    const bool compact = use_compaction<MODE>(P);
    Env e;
    if (valid) load_env<MODE>(P, state, n, i, e);
    const uint32_t gid = P.env_id_offset + (uint32_t)i;
    float sum_r = 0.f, fin = 0.f, chk = 0.f;
    for (int t = 0; t < steps; ++t) {
        float o_[D];
        StepOut so;
        so.needs_reset = false;
        if (valid) {
            const U4 r = philox4x32_10(U4{gid, t0 + (uint32_t)t, 0u, STREAM_ACTION}, P.philox_key);
            const float a[4] = {uniform_(r.x, -1.f, 1.f), uniform_(r.y, -1.f, 1.f), uniform_(r.z, -1.f, 1.f),
                                uniform_(r.w, -1.f, 1.f)};
            env_step<MODE, kGym>(P, T, gid, e, a, o_, nullptr, first ? first + i : nullptr, n, so);
            sum_r += so.reward;
            fin += so.finished ? 1.f : 0.f;
        }
        if constexpr (kGym) {
            if (compact) {
                warp_autoreset_smem<MODE>(P, gid - (uint32_t)(threadIdx.x & 31), e, o_, so.needs_reset,
                                          reinterpret_cast<WarpResetScratch*>(smem_raw)[threadIdx.x >> 5]);
            } else if (so.needs_reset) {
                float rpy[3];
                reset_env<MODE>(P, T, gid, e, rpy);
                compute_obs<MODE>(P, e, rpy, o_);
            }
        }
        if (valid) {
            float c = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) c += o_[k];
            chk += c;
        }
    }
    if (valid) {
        store_env<MODE>(P, state, n, i, e);
        if (stats) {
            stats[i] += sum_r; stats[(size_t)n + i] += fin; stats[2 * (size_t)n + i] += chk;
            stats[3 * (size_t)n + i] += (float)steps;
        }
    }
}

}  // namespace qs
