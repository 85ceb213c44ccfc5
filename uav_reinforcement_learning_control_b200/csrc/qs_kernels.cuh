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
#ifndef QS_RESET_STRATEGY
#define QS_RESET_STRATEGY 3   /* step kernel: 0 = inline per-lane reset, 1 = block-level compaction, 2 = warp-cooperative Philox
                                 (shuffles + __fns), 3 = warp-cooperative Philox through a per-warp shared scratch */
#endif
#ifndef QS_USE_PDL
#define QS_USE_PDL 1          /* step kernel: programmatic dependent launch (griddepcontrol) */
#endif
#ifndef QS_PREFETCH_AHEAD
#define QS_PREFETCH_AHEAD 518  /* step kernel: each CTA pulls the state planes / actions of the tile this many CTAs ahead into L2
                                   (148 SMs x 7 resident CTAs = 1036 per wave; half a wave measured best); 0 = off */
#endif
#ifndef QS_PREFETCH_L1
#define QS_PREFETCH_L1 0       /* step kernel: 1 = one L1 prefetch per thread for the CTA's own tile at the top of the kernel
                                   (as good as the L2 prefetch-ahead, not additive: 56.7 us either way) */
#endif
#ifndef QS_STREAM_HINTS
#define QS_STREAM_HINTS 1      /* step kernel: actions are read and obs / reward / done written with evict-first (.cs) policy so the
                                   state planes keep the L2 between steps */
#endif
#ifndef QS_TMA_STAGE
#define QS_TMA_STAGE 0         /* step kernel (gym modes): 1 = state planes + actions of a tile arrive through cp.async.bulk into smem
                                   (measured slower than direct loads + L2 prefetch: 62.7 vs 56.9 us; profiles/README.md) */
#endif
#ifndef QS_OBS_DIRECT
#define QS_OBS_DIRECT 1   /* 1: per-thread 128-bit obs stores (measured 3 % faster: two barriers fewer);
                             0: coalesced through a shared-memory tile */
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

// ------------------------------------------------------------------------------ compacted auto-reset
// Under random (or early-training) policies ~10 % of the envs finish an episode every step, so nearly
// every warp would run the ~570-instruction Philox reset path with 3 of 32 lanes active.  Instead the
// finished lanes enqueue themselves in shared memory, the first `count` threads of the block perform
// the resets densely (one env per lane), and the owners read their new state back.  Results are
// identical to the inline path (same reset_env / compute_obs source).
constexpr int kSlotF = 21 + 3 + 12;     // qpos/qvel, target, obs (gym modes only)

// Shared scratch of the compacted reset, sized for NT threads.
template <int NT>
struct ResetScratch {
    int wcnt[2][NT / 32];               // finished lanes per warp, double-buffered across steps
    unsigned short owner[NT];           // [warp][rank]: thread id of the finished lane
    uint32_t epi[NT];                   // its (already incremented) episode index
    float slot[NT * kSlotF];            // new state + obs, one row per reset
};

// Atomic-free: every warp ranks its finished lanes with a ballot and writes them into its own 32-entry
// region; after ONE barrier all threads know the per-warp counts, thread j picks the j-th entry, and
// after a second barrier the owners read their row back.
struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// `tid` is the thread's index inside the group of NT threads that `sync` synchronises (a whole CTA by default,
// one 128-thread tile in the tensor-core rollout kernel).
template <int MODE, int NT, class Sync = BlockSync>
__device__ __forceinline__ void block_autoreset(const QsParams& P, const Tables& T, uint32_t gid_block_first,
                                                Env& e, float* obs, bool need, int parity, ResetScratch<NT>& S,
                                                int tid = threadIdx.x, Sync sync = Sync()) {
    static_assert(ModeTraits<MODE>::kGym, "Philox re-sampling exists in the gym modes only");
    constexpr int NW = NT / 32;
    const int lane = tid & 31, w = tid >> 5;
    const unsigned ballot = __ballot_sync(0xffffffffu, need);
    const int rank = __popc(ballot & ((1u << lane) - 1u));
    if (need) { S.owner[w * 32 + rank] = (unsigned short)tid; S.epi[w * 32 + rank] = e.episode; }
    if (lane == 0) S.wcnt[parity][w] = __popc(ballot);
    sync();
    int cnt[NW], total = 0, base = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) { cnt[k] = S.wcnt[parity][k]; base += (k < w) ? cnt[k] : 0; total += cnt[k]; }
    if (total == 0) return;                         // block-uniform
    for (int j = tid; j < total; j += NT) {
        int k = j, ww = 0;
#pragma unroll
        for (int q = 0; q < NW - 1; ++q) { if (k >= cnt[q] && ww == q) { k -= cnt[q]; ww = q + 1; } }
        const int src = ww * 32 + k;
        Env r;
        r.episode = S.epi[src];
        r.wp_idx = 0; r.wp_reached = 0; r.laps = 0;
        float rpy[3];
        reset_env<MODE>(P, T, gid_block_first + (uint32_t)S.owner[src], r, rpy);
        float o_[12];
        compute_obs<MODE>(P, r, rpy, o_);
        float* d = S.slot + j * kSlotF;
        d[0] = r.b.p[0]; d[1] = r.b.p[1]; d[2] = r.b.p[2];
        d[3] = r.b.q[0]; d[4] = r.b.q[1]; d[5] = r.b.q[2]; d[6] = r.b.q[3];
        d[7] = r.b.v[0]; d[8] = r.b.v[1]; d[9] = r.b.v[2];
        d[10] = r.b.w[0]; d[11] = r.b.w[1]; d[12] = r.b.w[2];
        d[13] = r.target[0]; d[14] = r.target[1]; d[15] = r.target[2];
#pragma unroll
        for (int k2 = 0; k2 < 12; ++k2) d[16 + k2] = o_[k2];
    }
    sync();
    if (need) {
        const float* d = S.slot + (base + rank) * kSlotF;
        e.b.p[0] = d[0]; e.b.p[1] = d[1]; e.b.p[2] = d[2];
        e.b.q[0] = d[3]; e.b.q[1] = d[4]; e.b.q[2] = d[5]; e.b.q[3] = d[6];
        e.b.v[0] = d[7]; e.b.v[1] = d[8]; e.b.v[2] = d[9];
        e.b.w[0] = d[10]; e.b.w[1] = d[11]; e.b.w[2] = d[12];
        e.target[0] = d[13]; e.target[1] = d[14]; e.target[2] = d[15];
#pragma unroll
        for (int k = 0; k < 4; ++k) { e.b.th[k] = 0.f; e.b.s[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 12; ++k) obs[k] = d[16 + k];
        e.step_count = 0; e.ep_steps = 0; e.done_prev = 0.f; e.voltage = P.v_nominal;
#pragma unroll
        for (int k = 0; k < 3; ++k) e.rate_int[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) e.prev_action[k] = 0.f;
    }
}

// Warp-cooperative alternative without any block barrier: the Philox blocks of up to 8 finished lanes are computed
// by 32 lanes at once (lane L -> reset L/4, block L%4) and handed to the owners with shuffles; only the cheap
// post-processing (range map, Euler -> quaternion, observation) runs at low lane efficiency.
template <int MODE>
__device__ __forceinline__ void warp_autoreset(const QsParams& P, const Tables& T, uint32_t gid_warp_first, Env& e,
                                               float* obs, bool need) {
    static_assert(ModeTraits<MODE>::kGym, "Philox re-sampling exists in the gym modes only");
    const int lane = threadIdx.x & 31;
    unsigned pending = __ballot_sync(0xffffffffu, need);
    while (pending) {
        const int npass = min(__popc(pending), 8);
        const int r = lane >> 2, blk = lane & 3;
        // lane that owns the r-th pending reset (r < npass)
        const int src = r < npass ? (int)__fns(pending, 0, r + 1) : 0;
        const uint32_t epi_src = __shfl_sync(0xffffffffu, e.episode, src);
        U4 rnd = U4{0u, 0u, 0u, 0u};
        if (r < npass) rnd = philox4x32_10(U4{gid_warp_first + (uint32_t)src, epi_src, (uint32_t)blk, STREAM_RESET}, P.philox_key);
        const int myrank = __popc(pending & ((1u << lane) - 1u));
        const bool mine = need && ((pending >> lane) & 1u) && myrank < 8;
        const int from = (mine ? myrank : 0) * 4;
        uint32_t w[16];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            w[4 * b + 0] = __shfl_sync(0xffffffffu, rnd.x, from + b);
            w[4 * b + 1] = __shfl_sync(0xffffffffu, rnd.y, from + b);
            w[4 * b + 2] = __shfl_sync(0xffffffffu, rnd.z, from + b);
            w[4 * b + 3] = __shfl_sync(0xffffffffu, rnd.w, from + b);
        }
        if (mine) {
            float s12[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) s12[k] = uniform_(w[k], P.init_lo[k], P.init_hi[k]);
            e.b.p[0] = s12[0]; e.b.p[1] = s12[1]; e.b.p[2] = s12[2];
            rpy_to_quat(&s12[3], e.b.q);
            e.b.v[0] = s12[6]; e.b.v[1] = s12[7]; e.b.v[2] = s12[8];
            e.b.w[0] = s12[9]; e.b.w[1] = s12[10]; e.b.w[2] = s12[11];
            if constexpr (MODE == QS_MODE_HOVER_GYM) {
                e.target[0] = uniform_(w[12], P.target_lo[0], P.target_hi[0]);
                e.target[1] = uniform_(w[13], P.target_lo[1], P.target_hi[1]);
                e.target[2] = uniform_(w[14], P.target_lo[2], P.target_hi[2]);
            } else {
                e.target[0] = e.b.p[0]; e.target[1] = e.b.p[1]; e.target[2] = e.b.p[2];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { e.b.th[k] = 0.f; e.b.s[k] = 0.f; e.prev_action[k] = 0.f; }
#pragma unroll
            for (int k = 0; k < 3; ++k) e.rate_int[k] = 0.f;
            e.step_count = 0; e.ep_steps = 0; e.done_prev = 0.f; e.voltage = P.v_nominal;
            compute_obs<MODE>(P, e, &s12[3], obs);
        }
        // drop the (up to 8) lowest pending bits
        unsigned done_bits = 0u;
        unsigned tmp = pending;
        for (int k = 0; k < npass; ++k) { const unsigned low = tmp & (0u - tmp); done_bits |= low; tmp ^= low; }
        pending &= ~done_bits;
    }
}

// Strategy 3: warp-cooperative reset through a small per-warp shared scratch, no block barrier and no __fns.
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

template <int MODE>
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
            const U4 rnd = philox4x32_10(U4{gid_warp_first + S.src[r], S.epi[r], (uint32_t)blk, STREAM_RESET}, P.philox_key);
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

// whether this launch uses the compacted path (block-uniform)
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

// ------------------------------------------------------------------------------ TMA staging helpers
// 1-D bulk copies (cp.async.bulk, the TMA engine without a tensor map) global -> shared, completion on an mbarrier.
__device__ __forceinline__ uint32_t smem_addr_(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr_(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr_(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr_(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr_(bar)) : "memory");
}
// bounded spin: a tile is a few microseconds away at worst; trap instead of hanging the GPU if the copy never lands
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_addr_(bar);
    uint32_t ok = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

// which state planes a gym-mode step reads (block-uniform)
template <class F>
__device__ __forceinline__ bool gym_plane_needed(const QsParams& P, int p) {
    if (p <= 24 || p == 26) return true;
    if (p == 25) return F::kBattery && P.battery != 0;
    if (p >= 28 && p <= 30) return F::kWaypoint && P.waypoint_mode != 0;
    if (p >= 32 && p <= 34) return F::kRate && P.rate_wrapper != 0;
    return false;
}
constexpr int kStagePlanes = 35;          // planes 0..34 can be read by a gym-mode step

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
#if QS_RESET_STRATEGY == 1 || !QS_OBS_DIRECT
    constexpr size_t kGymScratch = sizeof(ResetScratch<kBlock>);
#else
    constexpr size_t kGymScratch = sizeof(WarpResetScratch) * (kBlock / 32);
#endif
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
#if QS_PREFETCH_L1
    if constexpr (kGym) {
        // ptxas splits the 27 independent plane loads into three rounds to stay inside the register budget, and the warp
        // would pay the memory latency once per round; one register-free L1 prefetch per thread at the very top requests
        // all of the tile's lines at once, so rounds two and three hit in L1.
        if (block_first + kBlock <= lo + count) {
            constexpr int kLines = kBlock * 4 / 128;
            for (int l = threadIdx.x; l < 27 * kLines; l += kBlock) {
                const int p = l / kLines, c = l - p * kLines;
                if (p != 25 || (F::kBattery && P.battery))
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(state + (size_t)p * n + block_first + c * 32));
            }
        }
    }
#endif
    float o_[D];
    Env e;
    StepOut so;
    so.needs_reset = false;
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool staged = false;
#if QS_TMA_STAGE
    if constexpr (kGym) {
        // Full, 16-byte aligned tiles: all planes and the actions of the CTA are fetched by the TMA engine into shared
        // memory with one bulk copy per plane (issued by one lane each, all in flight at once, no registers held while
        // they fly), and every thread then reads its env with immediate-offset shared loads.  Otherwise ptxas splits
        // the 27 dependent-free global loads into several rounds to stay inside the register budget and the warp eats
        // the HBM latency once per round.
        __shared__ __align__(128) float stage[kStagePlanes][kBlock];
        __shared__ __align__(16) float4 stage_act[kBlock];
        __shared__ __align__(8) uint64_t stage_bar;
        staged = (block_first + kBlock <= lo + count) && ((n & 3) == 0) && ((block_first & 3) == 0) &&
                 ((reinterpret_cast<uintptr_t>(state) & 15u) == 0);
        if (staged) {                                       // block-uniform
            if (threadIdx.x == 0) mbar_init_(&stage_bar, 1);
            __syncthreads();
            const int t = threadIdx.x;
            if (t == 0) {
                uint32_t np = 0;
                for (int p = 0; p < kStagePlanes; ++p) np += gym_plane_needed<F>(P, p) ? 1u : 0u;
                mbar_expect_tx_(&stage_bar, np * (uint32_t)(kBlock * sizeof(float)) + (uint32_t)(kBlock * sizeof(float4)));
            }
            if (t < kStagePlanes) {
                if (gym_plane_needed<F>(P, t))
                    bulk_g2s_(&stage[t][0], state + (size_t)t * n + block_first, (uint32_t)(kBlock * sizeof(float)), &stage_bar);
            } else if (t == kStagePlanes) {
                bulk_g2s_(&stage_act[0], action + block_first, (uint32_t)(kBlock * sizeof(float4)), &stage_bar);
            }
            mbar_wait_(&stage_bar, 0);
            load_env<MODE, F>(P, &stage[0][0], kBlock, threadIdx.x, e);
            a4 = stage_act[threadIdx.x];
        }
    }
#endif
    if (valid) {
        if (!staged) {
            load_env<MODE, F>(P, state, n, i, e);
#if QS_STREAM_HINTS
            a4 = __ldcs(action + i);
#else
            a4 = action[i];
#endif
        }
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
#if QS_RESET_STRATEGY == 1
                block_autoreset<MODE, kBlock>(P, T, P.env_id_offset + (uint32_t)block_first, e, o_, so.needs_reset, 0,
                                              *reinterpret_cast<ResetScratch<kBlock>*>(smem_raw));
#elif QS_RESET_STRATEGY == 2
                warp_autoreset<MODE>(P, T, P.env_id_offset + (uint32_t)(i - (int)(threadIdx.x & 31)), e, o_, so.needs_reset);
#elif QS_RESET_STRATEGY == 3
                warp_autoreset_smem<MODE>(P, P.env_id_offset + (uint32_t)(i - (int)(threadIdx.x & 31)), e, o_, so.needs_reset,
                                          reinterpret_cast<WarpResetScratch*>(smem_raw)[threadIdx.x >> 5]);
#else
                if (so.needs_reset) {
                    float rpy[3];
                    reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, rpy);
                    compute_obs<MODE>(P, e, rpy, o_);
                }
#endif
            }
        }
        // one store sequence for all lanes (splitting it by needs_reset makes nearly every warp run it twice)
        if (valid) store_env<MODE, F>(P, state, n, i, e);
#if QS_OBS_DIRECT
        if (valid) store_obs12(obs, i, o_);
#else
        __syncthreads();                            // reset slots are dead: reuse the scratch as the obs tile
        store_rows<D>(obs, lo + count, block_first, o_, valid, reinterpret_cast<float*>(smem_raw));
#endif
    } else {
        if (valid) store_env<MODE, F>(P, state, n, i, e);
        store_rows<D>(obs, lo + count, block_first, o_, valid, reinterpret_cast<float*>(smem_raw));
    }
}

// ------------------------------------------------------------------------------ step, persistent TMA-pipelined variant
// Plain north-star configuration (FeatLean), full 16-byte aligned tiles.  One wave of persistent CTAs walks over the
// 128-env tiles.  A tile's 26 state planes and its actions are fetched by the TMA engine (cp.async.bulk, 512 B per
// plane) into ONE shared-memory stage; as soon as every warp has copied its envs from the stage into registers
// (immediate-offset LDS, then one mbarrier arrive per warp on `empty`), one lane -- the role rotates over the four
// warps -- re-arms `full` and issues the bulk copies of the CTA's NEXT tile, which then fly during the ~1200
// instructions of compute and the stores of the current tile.  No warp ever waits on HBM latency after its first
// tile, there is no block barrier in the loop, and the load phases of the resident CTAs no longer line up (the plain
// kernel runs in waves: every CTA of an SM loads, computes and stores at the same time).
#ifndef QS_USE_PIPELINED_STEP
#define QS_USE_PIPELINED_STEP 0   /* measured 74.5 us vs 56.8 us for the plain kernel on B200: see profiles/README.md */
#endif
constexpr int kPpPlanes = 27;                       // planes 0..26 (25 = voltage is skipped)

__device__ __forceinline__ void mbar_arrive_(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr_(bar)) : "memory");
}

struct PpSmem {
    float plane[kPpPlanes][kBlock];
    float4 action[kBlock];
    WarpResetScratch scratch[kBlock / 32];
    uint64_t full, empty;
};

__device__ __forceinline__ void pp_issue(PpSmem& S, const float* __restrict__ state, int n, const float4* __restrict__ action,
                                         int first) {
    mbar_expect_tx_(&S.full, (uint32_t)((kPpPlanes - 1) * kBlock * sizeof(float) + kBlock * sizeof(float4)));
#pragma unroll 1
    for (int p = 0; p < kPpPlanes; ++p) {
        if (p == 25) continue;
        bulk_g2s_(&S.plane[p][0], state + (size_t)p * n + first, (uint32_t)(kBlock * sizeof(float)), &S.full);
    }
    bulk_g2s_(&S.action[0], action + first, (uint32_t)(kBlock * sizeof(float4)), &S.full);
}

template <int MODE>
__global__ void __launch_bounds__(kBlock, QS_STEP_MIN_BLOCKS)
step_kernel_pp(const __grid_constant__ QsParams P, Tables T, int n, int lo, int ntiles, float* __restrict__ state,
               const float4* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
               float* __restrict__ done, float* __restrict__ trunc) {
    static_assert(ModeTraits<MODE>::kGym, "pipelined step kernel: gym modes");
    using F = FeatLean;
    constexpr int D = 12;
    __shared__ __align__(128) PpSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init_(&S.full, 1); mbar_init_(&S.empty, kBlock / 32); }
    __syncthreads();
#if QS_USE_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) pp_issue(S, state, n, action, lo + tile * kBlock);
    uint32_t ph = 0;
    int round = 0;
#pragma unroll 1
    for (; tile < ntiles; tile += gridDim.x, ph ^= 1u, ++round) {
        const int i = lo + tile * kBlock + tid;
        Env e;
        mbar_wait_(&S.full, ph);
        load_env<MODE, F>(P, &S.plane[0][0], kBlock, tid, e);
        const float4 a4 = S.action[tid];
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_(&S.empty);
            const int nxt = tile + (int)gridDim.x;
            if (nxt < ntiles && warp == (round & (kBlock / 32 - 1))) {
                mbar_wait_(&S.empty, ph);                        // all four warps hold their envs in registers
                pp_issue(S, state, n, action, lo + nxt * kBlock);
            }
        }
        __syncwarp();
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float o_[D];
        StepOut so;
        env_step<MODE, true, F>(P, T, P.env_id_offset + (uint32_t)i, e, a, o_, nullptr, nullptr, n, so);
#if QS_STREAM_HINTS
        __stcs(reward + i, so.reward);
        __stcs(done + i, so.done);
        if (trunc) __stcs(trunc + i, so.truncated);
#else
        reward[i] = so.reward;
        done[i] = so.done;
        if (trunc) trunc[i] = so.truncated;
#endif
        if (P.auto_reset == QS_RESET_RESAMPLE)
            warp_autoreset_smem<MODE>(P, P.env_id_offset + (uint32_t)(i - lane), e, o_, so.needs_reset, S.scratch[warp]);
        store_env<MODE, F>(P, state, n, i, e);
        store_obs12(obs, i, o_);
    }
}

// ------------------------------------------------------------------------------ step, prefetching variant
// Gym modes.  Persistent CTAs walk over 128-env tiles; while tile k is being computed the planes and actions of
// tile k+1 stream into a second shared-memory stage with cp.async (LDGSTS), so no warp ever waits on HBM latency
// (ncu on the plain kernel: 30 % of warp time in long_scoreboard).  The staged tile keeps the planar layout with
// stride kBlock, so the very same load_env reads it.  Stores, the compacted reset and every result are identical
// to step_kernel.
#ifndef QS_PF_MIN_BLOCKS
#define QS_PF_MIN_BLOCKS 4
#endif
#ifndef QS_PF_STAGES
#define QS_PF_STAGES 2                              // 2: double-buffered; 1: single stage refilled after the register load
#endif
constexpr int kPfPlanes = 31;                       // planes 0..30 (31 = brax-only)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct PfStage {
    float plane[kPfPlanes][kBlock];
    float4 action[kBlock];
};

// which planes a gym-mode step reads (block-uniform)
__device__ __forceinline__ bool pf_plane_needed(const QsParams& P, int p) {
    if (p <= 24 || p == 26) return true;
    if (p == 25) return P.battery != 0;
    if (p >= 28 && p <= 30) return P.waypoint_mode != 0;
    return false;
}

// issue the asynchronous copies of tile [first, first + rows) into `st`
__device__ __forceinline__ void pf_issue(const QsParams& P, PfStage& st, const float* __restrict__ state, int n,
                                         const float4* __restrict__ action, int first, int rows) {
    const int t = threadIdx.x;
    if (rows == kBlock && (n & 3) == 0 && (first & 3) == 0) {
        // 16-byte copies: kBlock/4 chunks per plane
        constexpr int CH = kBlock / 4;
        for (int idx = t; idx < kPfPlanes * CH; idx += kBlock) {
            const int p = idx / CH, c = idx - p * CH;
            if (pf_plane_needed(P, p)) cp_async16(&st.plane[p][4 * c], state + (size_t)p * n + first + 4 * c);
        }
    } else {
        for (int idx = t; idx < kPfPlanes * kBlock; idx += kBlock) {
            const int p = idx / kBlock, c = idx - p * kBlock;
            if (c < rows && pf_plane_needed(P, p)) cp_async4(&st.plane[p][c], state + (size_t)p * n + first + c);
        }
    }
    if (t < rows) cp_async16(&st.action[t], action + first + t);
    cp_async_commit();
}

template <int MODE>
__global__ void __launch_bounds__(kBlock, QS_PF_MIN_BLOCKS)
step_kernel_pf(const __grid_constant__ QsParams P, Tables T, int n, int lo, int count, float* __restrict__ state,
               const float4* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
               float* __restrict__ done, float* __restrict__ trunc, float* __restrict__ metrics,
               float* __restrict__ term_obs) {
    static_assert(ModeTraits<MODE>::kGym, "prefetching step kernel: gym modes");
    constexpr int D = 12;
    extern __shared__ __align__(16) unsigned char pf_smem[];
    PfStage* stage = reinterpret_cast<PfStage*>(pf_smem);                         // [QS_PF_STAGES]
    ResetScratch<kBlock>& scratch = *reinterpret_cast<ResetScratch<kBlock>*>(pf_smem + QS_PF_STAGES * sizeof(PfStage));
    const int ntiles = (count + kBlock - 1) / kBlock;
    int tile = blockIdx.x;
    if (tile >= ntiles) return;
    pf_issue(P, stage[0], state, n, action, lo + tile * kBlock, min(kBlock, count - tile * kBlock));
    int buf = 0, parity = 0;
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1, parity ^= 1) {
        const int block_first = lo + tile * kBlock;
        const int rows = min(kBlock, count - tile * kBlock);
        const int i = block_first + threadIdx.x;
        const bool valid = threadIdx.x < rows;
        cp_async_wait_all();
        __syncthreads();                                   // tile `tile` has landed in stage[buf]; stage[buf^1] is free
        const int nxt = tile + gridDim.x;
        constexpr int kCur = QS_PF_STAGES == 2 ? 1 : 0;
        PfStage& cur = stage[buf & kCur];
#if QS_PF_STAGES == 2
        if (nxt < ntiles) pf_issue(P, stage[buf ^ 1], state, n, action, lo + nxt * kBlock, min(kBlock, count - nxt * kBlock));
#endif
        float o_[D];
        Env e;
        StepOut so;
        so.needs_reset = false;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            load_env<MODE>(P, &cur.plane[0][0], kBlock, threadIdx.x, e);
            a4 = cur.action[threadIdx.x];
        }
#if QS_PF_STAGES == 1
        __syncthreads();                                   // everyone holds its env in registers: refill the stage
        if (nxt < ntiles) pf_issue(P, cur, state, n, action, lo + nxt * kBlock, min(kBlock, count - nxt * kBlock));
#endif
        if (valid) {
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float tobs[D];
            env_step<MODE, true>(P, T, P.env_id_offset + (uint32_t)i, e, a, o_, term_obs ? tobs : nullptr, nullptr, n, so);
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
        }
        if (P.auto_reset == QS_RESET_RESAMPLE) {
            if (P.waypoint_mode) {
                if (so.needs_reset) {
                    float rpy[3];
                    reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, rpy);
                    compute_obs<MODE>(P, e, rpy, o_);
                }
            } else {
                block_autoreset<MODE, kBlock>(P, T, P.env_id_offset + (uint32_t)block_first, e, o_, so.needs_reset, parity, scratch);
            }
        }
        if (valid) { store_env<MODE>(P, state, n, i, e); store_obs12(obs, i, o_); }
    }
}

inline size_t step_pf_smem_bytes() { return QS_PF_STAGES * sizeof(PfStage) + sizeof(ResetScratch<kBlock>); }

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
    __shared__ __align__(16) unsigned char smem_raw[kGym ? sizeof(ResetScratch<kBlock>) : 16];
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
#if QS_RESET_STRATEGY == 3
                warp_autoreset_smem<MODE>(P, gid - (uint32_t)(threadIdx.x & 31), e, o_, so.needs_reset,
                                          reinterpret_cast<WarpResetScratch*>(smem_raw)[threadIdx.x >> 5]);
#else
                block_autoreset<MODE, kBlock>(P, T, P.env_id_offset + (uint32_t)block_first, e, o_, so.needs_reset, t & 1,
                                              *reinterpret_cast<ResetScratch<kBlock>*>(smem_raw));
#endif
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
