// qs_step2.cuh -- the north-star env step with TWO adjacent envs per thread on Blackwell's packed FP32 pipe.
//
// Replaces, for the plain configuration the launcher calls "lean" (HoverEnv semantics, envs/hover_env.py:159-198, no
// battery sag / rate wrapper / waypoints / action pre-clip, no metrics / terminal_obs outputs), the one-env-per-thread
// step_kernel<QS_MODE_HOVER_GYM, FeatLean> of qs_kernels.cuh.  That kernel is ISSUE-bound (ncu: 41.3 M warp
// instructions per 2^20-env launch, issue slots 75 % busy, HBM 0.73-0.79 of roofline), and 43 % of its instructions are
// FFMA / FMUL / FADD.  Here thread i owns envs 2i and 2i+1:
//   * every state plane is read and written as one 64-bit word per thread (two adjacent envs: a warp moves 256
//     contiguous bytes per plane), actions as two float4, observations as six float4 -- half the memory instructions and
//     half the address arithmetic per env;
//   * the dynamics (qs_dynamics.cuh: physics_step<f2>, the SAME source as the scalar kernels), Euler angles,
//     observation normalisation and reward polynomial run on f2 operands: one FFMA2 / FMUL2 / FADD2 per two envs, with
//     constants broadcast from uniform registers for free (qs_pack2.cuh); only MUFU, min / max, compares and selects stay
//     per lane;
//   * the Philox auto-reset is the warp-cooperative scheme of qs_kernels.cuh over both halves at once: a lane queues
//     at most one of its two envs per pass, so the ~100-instruction low-occupancy tail (Euler -> quaternion, reset
//     observation) runs once per pass for the 64 envs of the warp instead of once per 32.
// Results: identical arithmetic per env up to the contraction choices ptxas makes between separate multiplies and adds
// (a few ulp); tests tie it to the oracle (tests/test_gpu_parity.py) and to the general scalar kernel.
#pragma once

#include "qs_kernels.cuh"

namespace qs {

#ifndef QS_STEP2_BLOCK
#define QS_STEP2_BLOCK 128          /* threads per CTA = 256 envs */
#endif
#ifndef QS_STEP2_MIN_BLOCKS
#define QS_STEP2_MIN_BLOCKS 4       /* <= 128 registers */
#endif
#ifndef QS_STEP2_PREFETCH_AHEAD
#define QS_STEP2_PREFETCH_AHEAD 148 /* CTAs ahead for the L2 prefetch (148 SMs x 4 CTAs per wave; 0 / 148 / 296 / 592 measured 0.831 / 0.851 / 0.842 / 0.808 of the HBM roofline) */
#endif
constexpr int kBlock2 = QS_STEP2_BLOCK;

struct WarpReset2Scratch {
    uint32_t gid[8];
    uint32_t epi[8];
    float4 val[32];
};

__device__ __forceinline__ f2 ld2_(const float* __restrict__ plane, int i) { const float2 v = reinterpret_cast<const float2*>(plane)[i]; return f2{v.x, v.y}; }
__device__ __forceinline__ void st2_(float* __restrict__ plane, int i, f2 v) { reinterpret_cast<float2*>(plane)[i] = make_float2(v.x, v.y); }

// One env step of a PAIR of envs up to (not including) the auto-reset: action -> motor forces -> dynamics -> Euler angles ->
// reward / termination / truncation -> observation.  Shared by the per-launch kernel and the state-resident one.
struct PairOut { f2 rew, dn; b2 inside; bool tr0, tr1; };
__device__ __forceinline__ void step_pair(const QsParams& P, BodyT<f2>& b, const f2 tgt[3], int& sc0, int& sc1, const f2 a[4],
                                          f2 o_[12], PairOut& out) {
    // ---- action -> motor forces (qs_env.cuh: action_to_ctrl<FeatLean>; hover_env.py:169-177) -------------------------
    f2 ctrl[4];
    {
        f2 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = fma_((a[k] + 1.0f) * 0.5f, P.act_hi[k] - P.act_lo[k], P.act_lo[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const f2 Fk = fma_(P.mix_inv[4 * k], u[0], fma_(P.mix_inv[4 * k + 1], u[1],
                          fma_(P.mix_inv[4 * k + 2], u[2], P.mix_inv[4 * k + 3] * u[3])));
            ctrl[k] = clamp_(Fk, 0.0f, P.max_motor_thrust);
        }
    }
    physics_step(P, b, ctrl);
    sc0 += 1; sc1 += 1;

    // ---- Euler angles (qs_env.cuh: quat_to_rpy; scipy 'xyz') ------------------------------------------------------------
    f2 rpy[3];
    {
        const f2 w = b.q[0], x = b.q[1], y = b.q[2], z = b.q[3];
        rpy[0] = atan2_(2.f * fma_(y, z, w * x), fma_(-2.f, fma_(x, x, y * y), 1.f));
        rpy[1] = asin_unit_(clamp_(2.f * fma_(w, y, -x * z), -1.f, 1.f));
        rpy[2] = atan2_(2.f * fma_(x, y, w * z), fma_(-2.f, fma_(y, y, z * z), 1.f));
    }
    // ---- reward / termination (qs_env.cuh: evaluate, gym branch; hover_env.py:138-157,188) ----------------------------
    const f2 dx = b.p[0] - tgt[0], dy = b.p[1] - tgt[1], dz = b.p[2] - tgt[2];
    const f2 e2 = fma_(dx, dx, fma_(dy, dy, dz * dz));
    out.rew = exp_(-e2);
    const f2 s12[12] = {b.p[0], b.p[1], b.p[2], rpy[0], rpy[1], rpy[2], b.v[0], b.v[1], b.v[2], b.w[0], b.w[1], b.w[2]};
    b2 inside = b2{true, true};
#pragma unroll
    for (int k = 0; k < 12; ++k) inside = inside && ge_(s12[k], P.term_lo[k]) && le_(s12[k], P.term_hi[k]);   // false for NaN / +-Inf
    out.inside = inside;
    out.dn = f2{inside.x ? 0.f : 1.f, inside.y ? 0.f : 1.f};
    out.tr0 = sc0 >= P.max_episode_steps; out.tr1 = sc1 >= P.max_episode_steps;
    // ---- observation (qs_env.cuh: compute_obs, gym branch; hover_env.py:126-136) -----------------------------------------
    {
        const f2 x[12] = {tgt[0] - b.p[0], tgt[1] - b.p[1], tgt[2] - b.p[2], rpy[0], rpy[1], rpy[2],
                          b.v[0], b.v[1], b.v[2], b.w[0], b.w[1], b.w[2]};
#pragma unroll
        for (int k = 0; k < 12; ++k) o_[k] = fma_(x[k] - P.obs_lo[k], P.obs_scale[k], -1.0f);
    }
}

// Warp-cooperative Philox re-sampling of the finished envs of a warp's 64 envs (both halves at once; see the header).
// gid_lane: global id of this lane's even env.  Must be called by all 32 lanes.
__device__ __forceinline__ void reset_pairs(const QsParams& P, uint32_t gid_lane, bool need0, bool need1, BodyT<f2>& b, f2 tgt[3],
                                            int& sc0, int& sc1, uint32_t& ep0, uint32_t& ep1, f2 o_[12], WarpReset2Scratch& S) {
    constexpr int MODE = QS_MODE_HOVER_GYM;
    // ---- VecEnv auto-reset: fresh Philox sample for finished envs (qs_env.cuh: env_step, gym branch) -------------------
    {
        if (need0) ep0 += 1u;
        if (need1) ep1 += 1u;
        const int lane = threadIdx.x & 31;
        unsigned pending = __ballot_sync(0xffffffffu, need0 || need1);
        while (pending) {                                   // warp-uniform; a lane serves one of its halves per pass
            const int rank = __popc(pending & ((1u << lane) - 1u));
            const bool mine = ((pending >> lane) & 1u) && rank < 8;
            const int h = need0 ? 0 : 1;
            if (mine) { S.gid[rank] = gid_lane + (uint32_t)h; S.epi[rank] = h ? ep1 : ep0; }
            __syncwarp();
            const int npass = min(__popc(pending), 8);
            const int r = lane >> 2, blk = lane & 3;
            if (r < npass) {
                const U4 rnd = philox4x32_10(U4{S.gid[r], S.epi[r], (uint32_t)blk, STREAM_RESET}, P.philox_key);
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
                const float4 va = S.val[4 * rank], vb = S.val[4 * rank + 1], vc = S.val[4 * rank + 2], vd = S.val[4 * rank + 3];
                Env e;                                       // scalar view of the half being reset (same code as warp_autoreset_smem)
                e.b.p[0] = va.x; e.b.p[1] = va.y; e.b.p[2] = va.z;
                float r3[3] = {va.w, vb.x, vb.y};
                rpy_to_quat(r3, e.b.q);
                e.b.v[0] = vb.z; e.b.v[1] = vb.w; e.b.v[2] = vc.x;
                e.b.w[0] = vc.y; e.b.w[1] = vc.z; e.b.w[2] = vc.w;
                e.target[0] = vd.x; e.target[1] = vd.y; e.target[2] = vd.z;
                float on[12];
                compute_obs<MODE>(P, e, r3, on);
                if (h == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { b.p[k].x = e.b.p[k]; b.v[k].x = e.b.v[k]; b.w[k].x = e.b.w[k]; tgt[k].x = e.target[k]; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) { b.q[k].x = e.b.q[k]; b.th[k].x = 0.f; b.s[k].x = 0.f; }
#pragma unroll
                    for (int k = 0; k < 12; ++k) o_[k].x = on[k];
                    sc0 = 0; need0 = false;
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { b.p[k].y = e.b.p[k]; b.v[k].y = e.b.v[k]; b.w[k].y = e.b.w[k]; tgt[k].y = e.target[k]; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) { b.q[k].y = e.b.q[k]; b.th[k].y = 0.f; b.s[k].y = 0.f; }
#pragma unroll
                    for (int k = 0; k < 12; ++k) o_[k].y = on[k];
                    sc1 = 0; need1 = false;
                }
            }
            pending = __ballot_sync(0xffffffffu, need0 || need1);   // also orders this pass's shared reads before the next pass's writes
        }
    }
}

// n: plane stride in envs (even); pair0: first env pair of this launch; npairs: env pairs to step
__global__ void __launch_bounds__(kBlock2, QS_STEP2_MIN_BLOCKS)
step2_kernel(const __grid_constant__ QsParams P, int n, int pair0, int npairs, float* __restrict__ state,
             const float4* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
             float* __restrict__ done, float* __restrict__ trunc) {
    constexpr int MODE = QS_MODE_HOVER_GYM;
    __shared__ __align__(16) WarpReset2Scratch scratch[kBlock2 / 32];
    const int block_first = pair0 + blockIdx.x * kBlock2;
    const int i = block_first + threadIdx.x;                 // env pair index: envs 2i, 2i + 1
    const bool valid = i < pair0 + npairs;
#if QS_USE_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
#if QS_STEP2_PREFETCH_AHEAD > 0
    {
        // L2 prefetch of the tile QS_STEP2_PREFETCH_AHEAD CTAs ahead: 26 planes x 8 lines + 32 action lines
        const int pf_first = block_first + QS_STEP2_PREFETCH_AHEAD * kBlock2;
        if (pf_first + kBlock2 <= pair0 + npairs) {
            constexpr int kLines = kBlock2 * 8 / 128;
            for (int l = threadIdx.x; l < 27 * kLines; l += kBlock2) {
                const int p = l / kLines, c = l - p * kLines;
                if (p != 25) asm volatile("prefetch.global.L2 [%0];" :: "l"(state + (size_t)p * n + 2 * pf_first + c * 32));
            }
            if (threadIdx.x < kBlock2 / 4) asm volatile("prefetch.global.L2 [%0];" :: "l"(action + 2 * pf_first + threadIdx.x * 8));
        }
    }
#endif
    BodyT<f2> b;
    f2 tgt[3];
    int sc0 = 0, sc1 = 0;
    uint32_t ep0 = 0u, ep1 = 0u;
    f2 a[4];
    if (valid) {
        const float* s = state;
        const size_t N = (size_t)n;
        b.p[0] = ld2_(s + 0 * N, i); b.p[1] = ld2_(s + 1 * N, i); b.p[2] = ld2_(s + 2 * N, i);
        b.q[0] = ld2_(s + 3 * N, i); b.q[1] = ld2_(s + 4 * N, i); b.q[2] = ld2_(s + 5 * N, i); b.q[3] = ld2_(s + 6 * N, i);
#pragma unroll
        for (int k = 0; k < 4; ++k) b.th[k] = ld2_(s + (7 + k) * N, i);
        b.v[0] = ld2_(s + 11 * N, i); b.v[1] = ld2_(s + 12 * N, i); b.v[2] = ld2_(s + 13 * N, i);
        b.w[0] = ld2_(s + 14 * N, i); b.w[1] = ld2_(s + 15 * N, i); b.w[2] = ld2_(s + 16 * N, i);
#pragma unroll
        for (int k = 0; k < 4; ++k) b.s[k] = ld2_(s + (17 + k) * N, i);
        tgt[0] = ld2_(s + 21 * N, i); tgt[1] = ld2_(s + 22 * N, i); tgt[2] = ld2_(s + 23 * N, i);
        const f2 sc = ld2_(s + 24 * N, i), ep = ld2_(s + 26 * N, i);
        sc0 = f2i_(sc.x); sc1 = f2i_(sc.y); ep0 = f2u_(ep.x); ep1 = f2u_(ep.y);
#if QS_STREAM_HINTS
        const float4 a0 = __ldcs(action + 2 * i), a1 = __ldcs(action + 2 * i + 1);
#else
        const float4 a0 = action[2 * i], a1 = action[2 * i + 1];
#endif
        a[0] = f2{a0.x, a1.x}; a[1] = f2{a0.y, a1.y}; a[2] = f2{a0.z, a1.z}; a[3] = f2{a0.w, a1.w};
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { b.p[k] = bc_(0.f); b.v[k] = bc_(0.f); b.w[k] = bc_(0.f); tgt[k] = bc_(0.f); }
#pragma unroll
        for (int k = 0; k < 4; ++k) { b.q[k] = bc_(k == 0 ? 1.f : 0.f); b.th[k] = bc_(0.f); b.s[k] = bc_(0.f); a[k] = bc_(0.f); }
    }

    f2 o_[12];
    PairOut po;
    step_pair(P, b, tgt, sc0, sc1, a, o_, po);
    const f2 rew = po.rew, dn = po.dn;
    const bool tr0 = po.tr0, tr1 = po.tr1;
    if (valid) {
#if QS_STREAM_HINTS
        __stcs(reinterpret_cast<float2*>(reward) + i, make_float2(rew.x, rew.y));
        __stcs(reinterpret_cast<float2*>(done) + i, make_float2(dn.x, dn.y));
        if (trunc) __stcs(reinterpret_cast<float2*>(trunc) + i, make_float2(tr0 ? 1.f : 0.f, tr1 ? 1.f : 0.f));
#else
        reinterpret_cast<float2*>(reward)[i] = make_float2(rew.x, rew.y);
        reinterpret_cast<float2*>(done)[i] = make_float2(dn.x, dn.y);
        if (trunc) reinterpret_cast<float2*>(trunc)[i] = make_float2(tr0 ? 1.f : 0.f, tr1 ? 1.f : 0.f);
#endif
    }
    // ---- VecEnv auto-reset: fresh Philox sample for finished envs (qs_env.cuh: env_step, gym branch) -------------------
    if (P.auto_reset == QS_RESET_RESAMPLE)
        reset_pairs(P, P.env_id_offset + 2u * (uint32_t)i, valid && (!po.inside.x || tr0), valid && (!po.inside.y || tr1), b, tgt,
                    sc0, sc1, ep0, ep1, o_, scratch[threadIdx.x >> 5]);
    if (valid) {
        float* s = state;
        const size_t N = (size_t)n;
        st2_(s + 0 * N, i, b.p[0]); st2_(s + 1 * N, i, b.p[1]); st2_(s + 2 * N, i, b.p[2]);
        st2_(s + 3 * N, i, b.q[0]); st2_(s + 4 * N, i, b.q[1]); st2_(s + 5 * N, i, b.q[2]); st2_(s + 6 * N, i, b.q[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) st2_(s + (7 + k) * N, i, b.th[k]);
        st2_(s + 11 * N, i, b.v[0]); st2_(s + 12 * N, i, b.v[1]); st2_(s + 13 * N, i, b.v[2]);
        st2_(s + 14 * N, i, b.w[0]); st2_(s + 15 * N, i, b.w[1]); st2_(s + 16 * N, i, b.w[2]);
#pragma unroll
        for (int k = 0; k < 4; ++k) st2_(s + (17 + k) * N, i, b.s[k]);
        st2_(s + 21 * N, i, tgt[0]); st2_(s + 22 * N, i, tgt[1]); st2_(s + 23 * N, i, tgt[2]);
        st2_(s + 24 * N, i, f2{i2f_(sc0), i2f_(sc1)});
        st2_(s + 26 * N, i, f2{u2f_(ep0), u2f_(ep1)});
        float4* d = reinterpret_cast<float4*>(obs + (size_t)i * 24);
#if QS_STREAM_HINTS
        __stcs(d + 0, make_float4(o_[0].x, o_[1].x, o_[2].x, o_[3].x));
        __stcs(d + 1, make_float4(o_[4].x, o_[5].x, o_[6].x, o_[7].x));
        __stcs(d + 2, make_float4(o_[8].x, o_[9].x, o_[10].x, o_[11].x));
        __stcs(d + 3, make_float4(o_[0].y, o_[1].y, o_[2].y, o_[3].y));
        __stcs(d + 4, make_float4(o_[4].y, o_[5].y, o_[6].y, o_[7].y));
        __stcs(d + 5, make_float4(o_[8].y, o_[9].y, o_[10].y, o_[11].y));
#else
        d[0] = make_float4(o_[0].x, o_[1].x, o_[2].x, o_[3].x);
        d[1] = make_float4(o_[4].x, o_[5].x, o_[6].x, o_[7].x);
        d[2] = make_float4(o_[8].x, o_[9].x, o_[10].x, o_[11].x);
        d[3] = make_float4(o_[0].y, o_[1].y, o_[2].y, o_[3].y);
        d[4] = make_float4(o_[4].y, o_[5].y, o_[6].y, o_[7].y);
        d[5] = make_float4(o_[8].y, o_[9].y, o_[10].y, o_[11].y);
#endif
    }
}

// ------------------------------------------------------------------------------ resident dyn-only rollout, packed
// The state-resident kernel of qs_kernels.cuh (rollout_random_kernel: T env steps per launch with the state in registers,
// actions U(-1,1)^4 from Philox stream 2 keyed by (global env id, global step index)) for the lean north-star
// configuration, two adjacent envs per thread.  That kernel is FP32-ISSUE-bound (ncu: issue slots 85 % busy), which is
// exactly what the packed pipe relieves.  stats[4][n] += (sum reward, episodes finished, obs checksum, steps).
__global__ void __launch_bounds__(kBlock2, QS_STEP2_MIN_BLOCKS)
rollout_random2_kernel(const __grid_constant__ QsParams P, int n, int npairs, float* __restrict__ state, int steps, uint32_t t0,
                       float* __restrict__ stats) {
    __shared__ __align__(16) WarpReset2Scratch scratch[kBlock2 / 32];
    const int i = blockIdx.x * kBlock2 + threadIdx.x;
    const bool valid = i < npairs;
    BodyT<f2> b;
    f2 tgt[3];
    int sc0 = 0, sc1 = 0;
    uint32_t ep0 = 0u, ep1 = 0u;
    const size_t N = (size_t)n;
    if (valid) {
        const float* s = state;
        b.p[0] = ld2_(s + 0 * N, i); b.p[1] = ld2_(s + 1 * N, i); b.p[2] = ld2_(s + 2 * N, i);
        b.q[0] = ld2_(s + 3 * N, i); b.q[1] = ld2_(s + 4 * N, i); b.q[2] = ld2_(s + 5 * N, i); b.q[3] = ld2_(s + 6 * N, i);
#pragma unroll
        for (int k = 0; k < 4; ++k) b.th[k] = ld2_(s + (7 + k) * N, i);
        b.v[0] = ld2_(s + 11 * N, i); b.v[1] = ld2_(s + 12 * N, i); b.v[2] = ld2_(s + 13 * N, i);
        b.w[0] = ld2_(s + 14 * N, i); b.w[1] = ld2_(s + 15 * N, i); b.w[2] = ld2_(s + 16 * N, i);
#pragma unroll
        for (int k = 0; k < 4; ++k) b.s[k] = ld2_(s + (17 + k) * N, i);
        tgt[0] = ld2_(s + 21 * N, i); tgt[1] = ld2_(s + 22 * N, i); tgt[2] = ld2_(s + 23 * N, i);
        const f2 sc = ld2_(s + 24 * N, i), ep = ld2_(s + 26 * N, i);
        sc0 = f2i_(sc.x); sc1 = f2i_(sc.y); ep0 = f2u_(ep.x); ep1 = f2u_(ep.y);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { b.p[k] = bc_(0.f); b.v[k] = bc_(0.f); b.w[k] = bc_(0.f); tgt[k] = bc_(0.f); }
#pragma unroll
        for (int k = 0; k < 4; ++k) { b.q[k] = bc_(k == 0 ? 1.f : 0.f); b.th[k] = bc_(0.f); b.s[k] = bc_(0.f); }
    }
    const uint32_t gid = P.env_id_offset + 2u * (uint32_t)i;
    f2 sum_r = bc_(0.f), fin = bc_(0.f), chk = bc_(0.f);
#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
        const U4 r0 = philox4x32_10(U4{gid, t0 + (uint32_t)t, 0u, STREAM_ACTION}, P.philox_key);
        const U4 r1 = philox4x32_10(U4{gid + 1u, t0 + (uint32_t)t, 0u, STREAM_ACTION}, P.philox_key);
        const f2 a[4] = {f2{uniform_(r0.x, -1.f, 1.f), uniform_(r1.x, -1.f, 1.f)}, f2{uniform_(r0.y, -1.f, 1.f), uniform_(r1.y, -1.f, 1.f)},
                         f2{uniform_(r0.z, -1.f, 1.f), uniform_(r1.z, -1.f, 1.f)}, f2{uniform_(r0.w, -1.f, 1.f), uniform_(r1.w, -1.f, 1.f)}};
        f2 o_[12];
        PairOut po;
        step_pair(P, b, tgt, sc0, sc1, a, o_, po);
        const bool f0 = !po.inside.x || po.tr0, f1 = !po.inside.y || po.tr1;
        if (valid) {
            sum_r += po.rew;
            fin += f2{f0 ? 1.f : 0.f, f1 ? 1.f : 0.f};
        }
        if (P.auto_reset == QS_RESET_RESAMPLE)
            reset_pairs(P, gid, valid && f0, valid && f1, b, tgt, sc0, sc1, ep0, ep1, o_, scratch[threadIdx.x >> 5]);
        if (valid) {
            f2 c = bc_(0.f);
#pragma unroll
            for (int k = 0; k < 12; ++k) c += o_[k];
            chk += c;
        }
    }
    if (valid) {
        float* s = state;
        st2_(s + 0 * N, i, b.p[0]); st2_(s + 1 * N, i, b.p[1]); st2_(s + 2 * N, i, b.p[2]);
        st2_(s + 3 * N, i, b.q[0]); st2_(s + 4 * N, i, b.q[1]); st2_(s + 5 * N, i, b.q[2]); st2_(s + 6 * N, i, b.q[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) st2_(s + (7 + k) * N, i, b.th[k]);
        st2_(s + 11 * N, i, b.v[0]); st2_(s + 12 * N, i, b.v[1]); st2_(s + 13 * N, i, b.v[2]);
        st2_(s + 14 * N, i, b.w[0]); st2_(s + 15 * N, i, b.w[1]); st2_(s + 16 * N, i, b.w[2]);
#pragma unroll
        for (int k = 0; k < 4; ++k) st2_(s + (17 + k) * N, i, b.s[k]);
        st2_(s + 21 * N, i, tgt[0]); st2_(s + 22 * N, i, tgt[1]); st2_(s + 23 * N, i, tgt[2]);
        st2_(s + 24 * N, i, f2{i2f_(sc0), i2f_(sc1)});
        st2_(s + 26 * N, i, f2{u2f_(ep0), u2f_(ep1)});
        if (stats) {
            float2* st = reinterpret_cast<float2*>(stats);
            const size_t h = N / 2;
            float2 v0 = st[i], v1 = st[h + i], v2 = st[2 * h + i], v3 = st[3 * h + i];
            v0.x += sum_r.x; v0.y += sum_r.y; v1.x += fin.x; v1.y += fin.y; v2.x += chk.x; v2.y += chk.y;
            v3.x += (float)steps; v3.y += (float)steps;
            st[i] = v0; st[h + i] = v1; st[2 * h + i] = v2; st[3 * h + i] = v3;
        }
    }
}

}  // namespace qs
