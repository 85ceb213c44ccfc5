// qs_ppo.cuh -- the PPO update on device: clipped-surrogate loss, analytic backward pass of the
// 2x128 actor and critic, gradient reduction, global-norm clip and Adam, sharing the rollout's trajectory buffers.
//
// Replaces the caller of the rollout: SB3 ``PPO.train`` as configured by train.py:50-68 (MlpPolicy pi=[128,128]
// vf=[128,128] ReLU, clip_range, ent_coef, vf_coef, max_grad_norm, normalize_advantage, Adam eps 1e-5), i.e. per
// minibatch   loss = -mean(min(r A, clip(r, 1-e, 1+e) A)) + vf_coef * mean((ret - V)^2) - ent_coef * H(pi).
//
// Kernels in this file:
//   ppo_grad_tc2_kernel     the production gradient kernel of the SB3 policy (12-D obs, Gaussian head): one network per
//                           CTA, two tiles in flight, issuer + gather warps
//   ppo_grad_tc_kernel<D, DIST>  (qs_ppo_generic.cuh) the single-tile schedule, templated on observation size and action
//                           distribution: the product path for the 21-D / tanh-normal (Brax) policies, and for <12, 0> the
//                           A/B reference of the kernel above (QS_PPO_V1=1); its description below introduces the GEMM
//                           formulation both share
//   obs_stats_*_kernel      (qs_ppo_generic.cuh) running observation normaliser (Welford / Chan merge)
//   ppo_adv_stats_kernel    advantage mean / std of a minibatch         ppo_reduce_kernel   fixed-order sum of the per-CTA rows
//   ppo_adam_kernel         global-norm clip + Adam (single GPU / NCCL)  ppo_peer_adam_kernel  the same fused with the gradient
//   ppo_permutation_kernel  per-epoch shuffle as a keyed bijection                             exchange over NVLink peer memory
//
// ppo_grad_tc_kernel: persistent CTAs (one per SM), each walking over 128-sample tiles of the minibatch.  Thread i owns
// sample i of the tile and TMEM lane i.  All seven GEMM families of the forward AND backward pass run on tcgen05:
//
//   forward   H1 = A0 . W1 (+b1)            M128 N128 K16        A0 = bf16 normalised obs + two constant-1 slots
//             H2 = A1 . W2 (+b2)            M128 N128 K128+16    A1 = bf16 relu(H1)
//             OUT = A2 . W3                 M128 N16  K128       A2 = bf16 relu(H2)
//   backward  dH2 = dOUT . W3^T             M128 N128 K16        B = the forward W3 operand read MN-major
//             dH1 = D2 . W2^T               M128 N128 K128       D2 = bf16(dH2 * [h2 > 0]); B = W2 operand read MN-major
//   weights   dW3 += A2^T . dOUT            M128 N16  K128(samples)   both operands MN-major views of the buffers above
//             dW2 += A1^T . D2              M128 N128 K128(samples)
//             db2 += D2^T . A0              M128 N16  K128(samples)   column 12 (the constant-1 slot) is sum_s D2
//             dW1^T,db1 += D1^T . A0        M128 N16  K128(samples)   D1 = bf16(dH1 * [h1 > 0]); column 12 is db1
//
// No operand is ever transposed in memory: a [sample][feature] activation tile written K-major for the forward GEMM is
// exactly the canonical MN-major no-swizzle layout of the [feature][sample] operand the weight-gradient GEMM needs
// (swap LBO and SBO, set the a_major / b_major bits of the instruction descriptor), and the forward weight operands
// read MN-major are the transposed weights of the backward GEMMs.  The weight gradients accumulate in fp32 TMEM
// columns across ALL tiles of the CTA (352 accumulator columns + 128 working columns) and leave the SM once, at the end,
// as one row of a [CTAs][P] partial-gradient array that ppo_reduce_kernel sums in a fixed order (bitwise reproducible,
// no atomics).  Actor and critic run back to back through the same working columns and operand buffers; the in-order
// tensor pipe plus one mbarrier commit per phase is all the synchronisation the buffer re-use needs.
// Numerics: bf16 operands, fp32 accumulation, fp32 master weights and Adam state (oracle/ppo_update_ref.py models the
// same rounding points).
#pragma once

#include "qs_rollout_tc.cuh"

namespace qs {
namespace ppo {

using namespace qs::tc;

constexpr int kD = 12;                 // observation size of the SB3 path (HoverEnv / TrajectoryFollowEnv)
constexpr int kPartialStats = 8;       // per-CTA statistics appended to its partial-gradient row
// floats between consecutive rows of the partial-gradient array (rows stay 128-byte aligned for the float4 stores)
__host__ __device__ constexpr int partial_stride(int P) { return (P + kPartialStats + 31) & ~31; }
// stats layout (sums over the minibatch): 0 policy loss, 1 value loss, 2 clipped samples, 3 approx KL, 4 samples

struct Batch {
    const float* obs;        // [N][12]
    const float* act;        // [N][4]   raw (unclipped) Gaussian samples, as stored by the rollout
    const float* old_logp;   // [N]
    const float* adv;        // [N]
    const float* ret;        // [N]
    const int32_t* idx;      // [n] minibatch sample indices into the N rows, or null: rows 0..n-1
    int n;
    const float* packed;     // optional [N][kRowF]: obs[D] | act[4] | old_logp | adv | ret | 0..: ONE 128-byte line per sample
                             // (ppo_pack_kernel); when set, the five arrays above are not read by the gradient kernels
};
constexpr int kRowF = 32;    // floats per packed sample row

// Five time-major rollout / GAE arrays -> one 128-byte row per sample.  The minibatch gather of the update reads random
// rows: from five separate arrays that costs 2-3 sectors for the 48-byte observation plus one 32-byte sector for each
// 4..16-byte field per network (ncu, round 1: 534 MB of DRAM reads for 80 MB of algorithmic bytes per 2^20 samples);
// from packed rows it is one full line per sample, shared by the actor and the critic CTA through L2.
// One CTA = 256 rows staged through shared memory, so that every global access is a fully coalesced 16 bytes per lane:
// the five source spans are contiguous (obs: 256 D floats), and a warp stores four complete rows (512 contiguous bytes)
// per instruction.  (One thread per float with five small reads per row was transaction-bound: 0.97 ms for 2^23 rows in
// the training loop, 3.4 x the HBM time of its 216 bytes per row.)
constexpr int kPackRows = 256;
template <int D>
__global__ void __launch_bounds__(kPackRows)
ppo_pack_kernel(const float* __restrict__ obs, const float* __restrict__ act, const float* __restrict__ old_logp,
                const float* __restrict__ adv, const float* __restrict__ ret, long long n, float* __restrict__ packed) {
    __shared__ __align__(16) float sObs[kPackRows * D];
    __shared__ __align__(16) float sAct[kPackRows * 4];
    __shared__ float sScal[3][kPackRows];
    const long long row0 = (long long)blockIdx.x * kPackRows;
    const int rows = (int)min((long long)kPackRows, n - row0);
    const int t = threadIdx.x;
    {
        // row0 D floats is a multiple of 16 bytes (row0 is a multiple of 256): float4 loads; the ragged tail by floats
        const float4* o4 = reinterpret_cast<const float4*>(obs + row0 * D);
        const int nv = rows * D / 4;
        for (int i = t; i < nv; i += kPackRows) reinterpret_cast<float4*>(sObs)[i] = __ldg(o4 + i);
        for (int i = nv * 4 + t; i < rows * D; i += kPackRows) sObs[i] = __ldg(obs + row0 * D + i);
        if (t < rows) {
            reinterpret_cast<float4*>(sAct)[t] = __ldg(reinterpret_cast<const float4*>(act) + row0 + t);
            sScal[0][t] = __ldg(old_logp + row0 + t); sScal[1][t] = __ldg(adv + row0 + t); sScal[2][t] = __ldg(ret + row0 + t);
        }
    }
    __syncthreads();
    const int c = t & 7;                       // 16-byte chunk of the row
#pragma unroll
    for (int j = 0; j < kPackRows / 32; ++j) {
        const int r = j * 32 + (t >> 3);
        if (r >= rows) continue;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int f = 4 * c + e;
            v[e] = f < D ? sObs[r * D + f] : f < D + 4 ? sAct[r * 4 + f - D] : f < D + 7 ? sScal[f - D - 4][r] : 0.f;
        }
        reinterpret_cast<float4*>(packed + (row0 + r) * kRowF)[c] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

struct Hyper {
    float clip_range, vf_coef, ent_coef;
    int normalize_adv;
};

constexpr uint32_t kTmemCols = 512;


// dh * [h > 0] for two elements (h: the packed bf16 pair of the forward activation, >= 0 after the relu), bf16-packed
__device__ __forceinline__ uint32_t pack_mask_bf16(float lo, float hi, uint32_t hpair) {
    lo = (hpair & 0xFFFFu) ? lo : 0.f;
    hi = (hpair >> 16) ? hi : 0.f;
    return pack_bf16(lo, hi);
}

// bf16x2(lo, hi) & [h > 0] per half: one F2FP pack, one HSET2 (0xffff per true half), one LOP
__device__ __forceinline__ uint32_t pack_mask2_bf16(float lo, float hi, uint32_t hpair) {
    uint32_t m, p;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(hpair), "r"(0u));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
    return p & m;
}

#ifdef QS_PPO_PROFILE
#define QS_PPOP(k) do { const long long c_ = clock64(); prof_[k] += c_ - pc_; pc_ = c_; } while (0)
#else
#define QS_PPOP(k) do { } while (0)
#endif

}  // namespace ppo
}  // namespace qs
#include "qs_ppo_generic.cuh"      // ppo_grad_tc_kernel<D, DIST>: the single-tile schedule for every policy variant
namespace qs {
namespace ppo {

// ---------------------------------------------------------------------------------------------------------------------
// ppo_grad_tc2_kernel -- the production schedule.  Same GEMMs, operands and rounding points as ppo_grad_tc_kernel, but
//   * one network per CTA (the first n_actor CTAs: actor, the rest: critic; every CTA walks over ALL tiles with its network's stride), which
//     halves the TMEM accumulator columns (176) and the weight operands (44 KB) a CTA has to hold, so that
//   * TWO 128-sample tiles are in flight per CTA (two slots, each with its own operand buffers and 128 working TMEM
//     columns, served by kNH worker warpgroups that split the columns of every epilogue) and share the gradient
//     accumulators, and
//   * a dedicated issuer warp owns the tensor pipe: workers hand a phase over with fence + mbarrier.arrive on
//     ready[slot] and wait on done[slot]; the issuer serves the two slots alternately, so one tile's MMAs run under
//     the other tile's epilogue (ping-pong), and no worker warp ever spends issue slots on tcgen05.mma.
//   * a gather warp fetches the next tile's minibatch rows (random rows of the rollout buffers) into a small
//     shared staging area with cp.async, one tile ahead, and hands them over through full / empty mbarriers: the workers
//     never have a global load outstanding (ptxas parked the epilogues behind such loads: shared scoreboards) nor a
//     cp.async of their own (their fence.proxy.async would wait for it);
//   * D2 / D1 are written IN PLACE over relu(H2) / relu(H1) (their only other readers, the dW3 / dW2 MMAs, are
//     committed before the workers are released), so a slot needs 80 KB of operands.
// What bounds it (ncu, round 2): the shared-memory data pipe.  UMMA operand fetch (SS form: A and B both from shared
// memory, 8 KB per M128 N128 K16 instruction = 128 B per clock at the tensor pipe's full rate) takes 45 % of the pipe's
// peak, the workers' epilogue loads / stores another 45 %.  Hence, per tile and network:
//   * db2 | dW2^T = D2^T . [1 | A1] is ONE N = 144 instruction per K step (a constant ONES block sits directly below A1,
//     so [ONES | A1] is one MN-major operand): 8 instead of 16 instructions and 28 KB less operand traffic than the
//     separate dW2 = A1^T . D2 and db2 = D2^T . A0 of the single-tile kernel; the accumulator is the TRANSPOSE of dW2
//     (lane = output feature), which the flush writes out with warp-coalesced stores;
//   * the head layer OUT = A2 . W3 runs in TS form: relu(H2) is ALSO stored into tensor memory (in place over consumed
//     working columns), so its eight N = 16 instructions read 0.5 KB instead of 4.5 KB each.
// 51 tcgen05.mma per tile and network (43 SS + 8 TS) instead of 59.  Measured and NOT kept (commit ee1aa8d has the code,
// profiles/README.md the numbers): relu masks as register bit masks instead of re-reading the activations (-64 KB of
// loads, but +7 % time: the epilogues' ALU work is on the tile's latency chain), the gather through registers or as
// 80-byte bulk copies (cp.async costs a wavefront per 16 bytes, but both alternatives were slower), and a staggered
// dynamic service order (the slots then contend for the pipe instead of taking turns).
struct SmemQ {
    static constexpr int W1 = 0, W2 = 4096, W3 = W2 + 32768, B2 = W3 + 4096, WEND = B2 + 4096;
    // per slot; ONES (128 rows x 16: a constant 1 in K slot 0) sits directly below A1: [ONES | A1] is ONE MN-major operand
    static constexpr int A0 = 0, ONES = 2 * 4096, A1 = ONES + 4096, A2 = A1 + 32768, DOUT = A2 + 32768, SLOT_BYTES = DOUT + 4096;
    static constexpr int SLOT0 = WEND;
    static constexpr int F32 = SLOT0 + 2 * SLOT_BYTES;
    static constexpr int kB3 = 0, kLogStd = 4, kInvSig = 8, kMean = 12, kInvStd = 24, kNumF = 36;
    static constexpr int RED = F32 + kNumF * 4;               // [8 warps][16]
    static constexpr int BAR = RED + 8 * 16 * 4;              // ready[2], done[2], full[2], empty[2], tmem slot
    // per slot: cp.async landing zone for the next tile's gathered rows: obs [128][12] | act [128][4] | 3 x [128] scalars
    static constexpr int STG = (BAR + 80 + 15) & ~15;
    static constexpr int STG_ACT = 128 * 48, STG_SCAL = STG_ACT + 128 * 16, STG_BYTES = STG_SCAL + 128 * 16;   // five-array gather: obs | act | 3 x [128] scalars; packed rows: [128][80 B]
    static constexpr int TOTAL = STG + 2 * STG_BYTES;
};
// working columns [128 slot, 128 slot + 128) | db2 (column 0 of 16) + dW2^T (128): lane = output feature | dW1^T | dW3
constexpr uint32_t kQColW = 0 /* + 128 * slot */, kQColW2 = 256, kQColW1 = 400, kQColW3 = 416, kQColX = 432 /* + 32 * slot */;
#ifndef QS_PPO_TS_HALF
#define QS_PPO_TS_HALF 1         /* the first four K steps (features 0..63) of layer 2 (A = relu(H1)) and of dH1 = D2 . W2^T (A = D2)
                                    read their A operand from the 2 x 32 spare TMEM columns (TS form) instead of shared memory:
                                    32 KB less operand fetch per tile and network.  (All eight would need 2 x 64 columns:
                                    2 x (128 + 64) + 176 > 512.) */
#endif
#ifndef QS_PPO_HALVES
#define QS_PPO_HALVES 2          /* worker warpgroups per slot: each takes 128 / QS_PPO_HALVES of the 128 columns of every epilogue
                                    (a second warpgroup reads the same TMEM lanes).  Measured 0.3485 vs 0.3511 ms per 2^20-sample
                                    minibatch for 2 vs 1: the epilogues do NOT get twice as fast, they wait for the shared-memory
                                    data pipe, which UMMA operand fetch and epilogue traffic share (profiles/README.md) */
#endif
constexpr int kNH = QS_PPO_HALVES;
constexpr int kIssuerWarp = 8 * kNH, kGatherWarp = 8 * kNH + 1;
constexpr int kThreads2 = 2 * kNH * 128 + 64;       // 2 slots x kNH x 128 workers + issuer warp + gather warp

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// partial: [max(n_actor, n_critic)][partial_stride(P)]; row r = actor CTA r (its network's entries, log_std, the
// untrained tail, statistics 0 2 3 4) and critic CTA r (its entries, statistic 1)
__global__ void __launch_bounds__(kThreads2, 1)
ppo_grad_tc2_kernel(Batch b, Hyper hp, const float* __restrict__ params, const float* __restrict__ adv_norm,
                    float* __restrict__ partial, int n_actor) {
    using S = SmemQ;
    extern __shared__ __align__(1024) unsigned char smem[];
    const PolicyLayout L = policy_layout(kD, 0);
    // the first n_actor CTAs own the actor, the rest the critic (the actor's tiles cost ~13 % more: loss math, more
    // gathered columns, so it gets ~53 % of the SMs); every CTA walks over all tiles with its network's stride
    const int net = (int)blockIdx.x >= n_actor ? 1 : 0;
    const int cta = net ? (int)blockIdx.x - n_actor : (int)blockIdx.x, ncta = net ? (int)gridDim.x - n_actor : n_actor;
    // warp-uniform role index (the shuffle makes the uniformity visible to ptxas: the issuer warp's descriptors then live
    // in uniform registers and its tcgen05.mma need no per-lane election loops)
    const int warp_id = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int gtid = threadIdx.x, wg = warp_id >> 2, tid = gtid & 127, warp = warp_id & 3, lane = gtid & 31;
    float* sF = reinterpret_cast<float*>(smem + S::F32);
    float* sRed = reinterpret_cast<float*>(smem + S::RED);
    uint64_t* bar_ready = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint64_t* bar_done = bar_ready + 2;
    uint64_t* bar_full = bar_ready + 4;
    uint64_t* bar_empty = bar_ready + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 64);

    // ---- one-time setup: this network's fp32 weights -> bf16 UMMA operands ------------------------------------------
    for (int i = gtid; i < S::WEND / 4; i += kThreads2) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    __syncthreads();
    auto put = [&](int base, int rows, int row, int k, float w) {
        *reinterpret_cast<__nv_bfloat16*>(smem + base + op_offset(rows, row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(w);
    };
    const int oW1 = net ? L.cW1 : L.aW1, ob1 = net ? L.cb1 : L.ab1, oW2 = net ? L.cW2 : L.aW2, ob2 = net ? L.cb2 : L.ab2;
    for (int i = gtid; i < kD * kH; i += kThreads2) put(S::W1, 128, i % kH, i / kH, params[oW1 + i]);
    for (int n = gtid; n < kH; n += kThreads2) {
        const float bv[2] = {params[ob1 + n], params[ob2 + n]};
        const int dst[2] = {S::W1, S::B2};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(bv[q]);
            put(dst[q], 128, n, 12, __bfloat162float(hi));
            put(dst[q], 128, n, 13, bv[q] - __bfloat162float(hi));
        }
    }
#pragma unroll 8
    for (int i = gtid; i < kH * kH; i += kThreads2) put(S::W2, 128, i % kH, i / kH, __ldg(params + oW2 + i));   // 8 loads in flight
    if (net == 0) { for (int i = gtid; i < kH * kA; i += kThreads2) put(S::W3, 16, i % kA, i / kA, params[L.aW3 + i]); }
    else          { for (int k = gtid; k < kH; k += kThreads2) put(S::W3, 16, 0, k, params[L.cW3 + k]); }
    for (int i = gtid; i < 2 * 2 * 128; i += kThreads2) {          // ONES: [slot][K chunk][row] x 16 B, a bf16 1.0 in K slot 0
        const int sl_ = i >> 8, c = (i >> 7) & 1, row = i & 127;
        *reinterpret_cast<uint4*>(smem + S::SLOT0 + sl_ * S::SLOT_BYTES + S::ONES + op_offset(128, row, c)) =
            make_uint4(c == 0 ? 0x3F80u : 0u, 0u, 0u, 0u);
    }
    if (gtid < kA) {
        sF[S::kB3 + gtid] = net ? (gtid == 0 ? params[L.cb3] : 0.f) : params[L.ab3 + gtid];
        const float ls = params[L.log_std + gtid];
        sF[S::kLogStd + gtid] = ls;
        sF[S::kInvSig + gtid] = expf(-ls);
    }
    if (gtid < kD) { sF[S::kMean + gtid] = params[L.mean + gtid]; sF[S::kInvStd + gtid] = params[L.inv_std + gtid]; }
    if (gtid == 0) {
        mbar_init(&bar_ready[0], kM * kNH); mbar_init(&bar_ready[1], kM * kNH); mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1);
        mbar_init(&bar_full[0], 32); mbar_init(&bar_full[1], 32); mbar_init(&bar_empty[0], kM * kNH); mbar_init(&bar_empty[1], kM * kNH);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_id == kIssuerWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    fence_async_smem();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sb = smem_u32(smem);
    const int ntiles = (b.n + kM - 1) / kM;
    const int iters = (ntiles + 2 * ncta - 1) / (2 * ncta);

    float g_b3[kA] = {0.f, 0.f, 0.f, 0.f}, g_ls[kA] = {0.f, 0.f, 0.f, 0.f};
    float st_loss = 0.f, st_clip = 0.f, st_kl = 0.f, st_n = 0.f;

    if (warp_id == kGatherWarp) {
        // =============================================== gather warp ===================================================
        // tile t of slot s: rows (2 (cta + t ncta) + s) 128 .. + 127 of the minibatch, 4 rows per lane
        uint32_t ph_e[2] = {0u, 0u};
#pragma unroll 1
        for (int t = 0; t < iters; ++t) {
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                if (t > 0) { mbar_wait(&bar_empty[s], ph_e[s]); ph_e[s] ^= 1u; }       // the workers have read tile t - 1
                unsigned char* stg = smem + S::STG + s * S::STG_BYTES;
                const int row0 = (2 * (cta + t * ncta) + s) * kM;
                if (b.packed) {
                    // packed rows (one 128-byte line per sample: obs | act | old_logp, adv, ret, 0): the staging area is
                    // [128 rows][80 B] and lane l of copy k moves 16-byte chunk 32 k + l of it, so the 32 destinations of one
                    // LDGSTS are 512 CONTIGUOUS bytes (5 adjacent lanes share a row).  With one row per lane (destinations 48 B
                    // apart) every lane's 16 bytes cost a shared-memory wavefront of their own -- ncu: 32 instead of 4 per
                    // instruction, a quarter of the kernel's LSU wavefronts; 0.346 -> 0.336 ms per 2^20-sample minibatch
                    int jj[20];
#pragma unroll
                    for (int k = 0; k < 20; ++k) {
                        const int r = row0 + (32 * k + lane) / 5;
                        jj[k] = r < b.n ? (b.idx ? __ldg(b.idx + r) : r) : -1;
                    }
#pragma unroll
                    for (int k = 0; k < 20; ++k) {
                        const int g = 32 * k + lane, c = g % 5;
                        if (jj[k] >= 0)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                         :: "r"(smem_u32(stg + g * 16)), "l"(b.packed + (size_t)jj[k] * kRowF + c * 4) : "memory");
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(&bar_full[s])) : "memory");
                    continue;
                }
                int j[4];                                  // five separate arrays: one row per lane and copy
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = row0 + q * 32 + lane;
                    j[q] = r < b.n ? (b.idx ? __ldg(b.idx + r) : r) : -1;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (j[q] < 0) continue;
                    const int row = q * 32 + lane;
                    const uint32_t so = smem_u32(stg + row * 48);
                    const float* o = b.obs + (size_t)j[q] * kD;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(so + c * 16), "l"(o + c * 4) : "memory");
                    const uint32_t ss = smem_u32(stg + S::STG_SCAL + row * 4);
                    if (net == 0) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(stg + S::STG_ACT + row * 16)), "l"(b.act + (size_t)j[q] * 4) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(ss), "l"(b.old_logp + j[q]) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(ss + 512), "l"(b.adv + j[q]) : "memory");
                    } else {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(ss + 1024), "l"(b.ret + j[q]) : "memory");
                    }
                }
                // full[s] completes when the copies of all 32 lanes have landed
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(&bar_full[s])) : "memory");
            }
        }
    } else if (warp_id == kIssuerWarp) {
        // =============================================== issuer ========================================================
        // the whole warp walks the schedule (uniform control flow); one elected lane issues
        {
            // descriptor = per-layout base (start address 0 of the CTA's shared window) + byte offset / 16: one 64-bit add
            // per operand on the single issuing thread instead of a dozen shifts and ors
            const uint64_t dK128 = make_desc(sb, 2048u, 128u), dK16 = make_desc(sb, 256u, 128u);
            const uint64_t dMN2048 = make_desc(sb, 128u, 2048u), dMN256 = make_desc(sb, 128u, 256u);
            auto dk = [&](int off, int rows) { return (rows == 128 ? dK128 : dK16) + (uint64_t)(off >> 4); };
            auto dmn = [&](int off, uint32_t grp_stride) { return (grp_stride == 2048u ? dMN2048 : dMN256) + (uint64_t)(off >> 4); };
            const uint32_t id_kk128 = idesc_mn(128, 128, 0, 0), id_kk16 = idesc_mn(128, 16, 0, 0);
            const uint32_t id_kmn128 = idesc_mn(128, 128, 0, 1);
            const uint32_t id_mm144 = idesc_mn(128, 144, 1, 1), id_mm16 = idesc_mn(128, 16, 1, 1);
            uint32_t ph[2] = {0u, 0u};
            // Every hand-off is strictly "workers signal ready -> issuer issues -> commit -> workers wait", one phase at
            // a time per slot: a slot can never run two phases ahead of the issuer, which would alias the mbarrier parity.
            // one phase of slot s (tile `it` of the slot); `first` = 0 only for the very first accumulation of the CTA
            auto issue_phase = [&](int s, int phase, int it) {
                const int base = S::SLOT0 + s * S::SLOT_BYTES;
                const int a0 = base + S::A0 + (it & 1) * 4096, a1 = base + S::A1, a2 = base + S::A2, dout = base + S::DOUT;
                const int ones = base + S::ONES;
                const uint32_t tw = tmem + kQColW + (uint32_t)(s * 128), tx = tmem + kQColX + (uint32_t)(s * 32);
                const uint32_t first = (it == 0 && s == 0) ? 0u : 1u;
                if (elect_one()) {
                    switch (phase) {
                    case 0:     // H1 of the slot's first tile
                        mma_bf16(tw, dk(base + S::A0, 128), dk(S::W1, 128), id_kk128, 0u);
                        break;
                    case 1:     // H2 = A1 . W2 + b2
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (QS_PPO_TS_HALF && j < 4) mma_bf16_ts(tw, tx + 8u * (uint32_t)j, dk(S::W2 + j * 4096, 128), id_kk128, j > 0);
                            else mma_bf16(tw, dk(a1 + j * 4096, 128), dk(S::W2 + j * 4096, 128), id_kk128, j > 0);
                        }
                        mma_bf16(tw, dk(a0, 128), dk(S::B2, 128), id_kk128, 1u);
                        break;
                    case 2:     // OUT = A2 . W3: TS form, relu(H2) read from working columns [0, 32) | [64, 96), OUT -> [32, 48)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            mma_bf16_ts(tw + 32u, tw + 64u * (uint32_t)(j >> 2) + 8u * (uint32_t)(j & 3), dk(S::W3 + j * 512, 16), id_kk16, j > 0);
                        }
                        break;
                    case 3:     // dH2 = dOUT . W3^T ; dW3 += A2^T . dOUT (must finish before D2 overwrites A2)
                        mma_bf16(tw, dk(dout, 128), dmn(S::W3, 256u), id_kmn128, 0u);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            mma_bf16(tmem + kQColW3, dmn(a2 + j * 256, 2048u), dmn(dout + j * 256, 2048u), id_mm16, j > 0 ? 1u : first);
                        break;
                    case 4:     // dH1 = D2 . W2^T ; db2 | dW2^T += D2^T . [1 | A1]   (D2 lives in the A2 buffer; the constant
                                // ONES block sits right below A1, so one N = 144 instruction per K step reads both)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (QS_PPO_TS_HALF && j < 4) mma_bf16_ts(tw, tx + 8u * (uint32_t)j, dmn(S::W2 + j * 256, 2048u), id_kmn128, j > 0);
                            else mma_bf16(tw, dk(a2 + j * 4096, 128), dmn(S::W2 + j * 256, 2048u), id_kmn128, j > 0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            mma_bf16(tmem + kQColW2, dmn(a2 + j * 256, 2048u), dmn(ones + j * 256, 2048u), id_mm144, j > 0 ? 1u : first);
                        break;
                    default:    // dW1^T | db1 += D1^T . A0 (D1 lives in the A1 buffer), then H1 of the slot's NEXT tile,
                                // whose A0 the workers wrote into the other A0 buffer together with D1
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            mma_bf16(tmem + kQColW1, dmn(a1 + j * 256, 2048u), dmn(a0 + j * 256, 2048u), id_mm16, j > 0 ? 1u : first);
                        if (it + 1 < iters)
                            mma_bf16(tw, dk(base + S::A0 + ((it + 1) & 1) * 4096, 128), dk(S::W1, 128), id_kk128, 0u);
                        break;
                    }
                    mma_commit(&bar_done[s]);
                }
                __syncwarp();
            };
#ifdef QS_PPO_PROFILE
            long long prof_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long pc_ = clock64();
#endif
            // fixed service order slot 0, slot 1, slot 0, ...: the accumulation order into the gradient accumulators is
            // then the same in every run, i.e. the gradient is bitwise reproducible
#pragma unroll
            for (int s = 0; s < 2; ++s) {               // prologue: H1 of each slot's first tile
                mbar_wait(&bar_ready[s], ph[s]); ph[s] ^= 1u;
                fence_after();
                issue_phase(s, 0, 0);
            }
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int phase = 1; phase < 6; ++phase) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        mbar_wait(&bar_ready[s], ph[s]); ph[s] ^= 1u;
                        fence_after();
                        QS_PPOP(0);
                        issue_phase(s, phase, it);
                        QS_PPOP(1);
                    }
                }
            }
#ifdef QS_PPO_PROFILE
            if (cta == 0 && lane == 0)
                printf("ppoprof net %d issuer (cycles per iteration of 2 tiles): issuing %lld | waiting for ready %lld\n",
                       net, prof_[1] / iters, prof_[0] / iters);
#endif
        }
    } else {
        // =============================================== workers =======================================================
        // kNH warpgroups per slot; all see the slot's 128 TMEM lanes (warp & 3 selects the lane quarter), each takes its
        // share of the column chunks of every epilogue.  The first one also owns the per-sample loss, the last one the
        // next tile's layer-1 operand.
        const int slot = wg / kNH, half = wg % kNH;
        const bool do_loss = half == 0, do_a0 = half == kNH - 1;
        constexpr int kCh = 4 / kNH;                     // 32-column chunks per warpgroup
        const int c_lo = half * kCh;
        unsigned char* sl = smem + S::SLOT0 + slot * S::SLOT_BYTES;
        const uint32_t tw = tmem + ((uint32_t)(warp * 32) << 16) + kQColW + (uint32_t)(slot * 128);
        uint64_t* ready = &bar_ready[slot];
        uint64_t* done = &bar_done[slot];
        uint32_t ph = 0u;
        const float adv_mean = hp.normalize_adv ? adv_norm[0] : 0.f;
        const float adv_istd = hp.normalize_adv ? adv_norm[1] : 1.f;
        auto signal = [&]() { fence_async_smem(); fence_before(); mbar_arrive(ready); };
        auto wait_done = [&]() { mbar_wait(done, ph); ph ^= 1u; fence_after(); };
        // Epilogues: the TMEM load of column chunk c + 1 is in flight while chunk c is converted and stored.
        // also_tmem: the packed activations additionally go back into tensor memory as the A operand of a TS-form MMA, in
        // place over consumed working columns: the 16 packed columns of chunk c (columns [32 c, 32 c + 32), loaded before)
        // go to [64 (c >> 1) + 16 (c & 1), + 16) -- inside the range this warpgroup itself has already loaded, whatever kNH
        const uint32_t tx = tmem + ((uint32_t)(warp * 32) << 16) + kQColX + (uint32_t)(slot * 32);    // TS-form copy of features 0..63
        auto epilogue_relu = [&](int dst, bool also_tmem) {
            uint32_t r[2][32];
            tmem_ld32_async(tw + (uint32_t)(c_lo * 32), r[0]);
            tmem_ld_wait(r[0]);
#pragma unroll
            for (int i = 0; i < kCh; ++i) {
                const int c = c_lo + i;
                if (i + 1 < kCh) tmem_ld32_async(tw + (uint32_t)((c + 1) * 32), r[(i + 1) & 1]);
                const uint32_t* v = r[i & 1];
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) pk[q] = pack_relu_bf16_u(v[2 * q], v[2 * q + 1]);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(sl + dst + op_offset(128, tid, c * 4 + q)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                if (also_tmem) tmem_st16(tw + (uint32_t)(64 * (c >> 1) + 16 * (c & 1)), pk);
                else if (QS_PPO_TS_HALF && c < 2) tmem_st16(tx + (uint32_t)(16 * c), pk);        // layer 1: A1's first four K steps
                if (i + 1 < kCh) tmem_ld_wait(r[(i + 1) & 1]);
            }
            if (also_tmem || (QS_PPO_TS_HALF && c_lo < 2)) tmem_st_wait();
        };
        // also_tx: the first four K steps of the result additionally go to the spare TMEM columns (dH1's TS-form A operand)
        auto epilogue_mask_inplace = [&](int buf, bool also_tx) {       // buf <- bf16(working columns * [buf > 0])
            uint32_t r[2][32];
            tmem_ld32_async(tw + (uint32_t)(c_lo * 32), r[0]);
            tmem_ld_wait(r[0]);
#pragma unroll
            for (int i = 0; i < kCh; ++i) {
                const int c = c_lo + i;
                if (i + 1 < kCh) tmem_ld32_async(tw + (uint32_t)((c + 1) * 32), r[(i + 1) & 1]);
                const uint32_t* v = r[i & 1];
#pragma unroll
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t* g = v + q * 8;
                    uint4* p4 = reinterpret_cast<uint4*>(sl + buf + op_offset(128, tid, c * 4 + q));
                    const uint4 h = *p4;
                    pk[4 * q] = pack_mask2_bf16(__uint_as_float(g[0]), __uint_as_float(g[1]), h.x);
                    pk[4 * q + 1] = pack_mask2_bf16(__uint_as_float(g[2]), __uint_as_float(g[3]), h.y);
                    pk[4 * q + 2] = pack_mask2_bf16(__uint_as_float(g[4]), __uint_as_float(g[5]), h.z);
                    pk[4 * q + 3] = pack_mask2_bf16(__uint_as_float(g[6]), __uint_as_float(g[7]), h.w);
                    *p4 = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                if (QS_PPO_TS_HALF && also_tx && c < 2) tmem_st16(tx + (uint32_t)(16 * c), pk);
                if (i + 1 < kCh) tmem_ld_wait(r[(i + 1) & 1]);
            }
            if (QS_PPO_TS_HALF && also_tx && c_lo < 2) tmem_st_wait();
        };
        // this tile's rows arrive through the gather warp's staging area (full / empty mbarriers); every warpgroup of the
        // slot takes what its role needs and releases the buffer: the loss warpgroup keeps action + scalars in registers
        // until the tile's loss, the layer-1 warpgroup turns the observation straight into the A0 operand
        unsigned char* stg = smem + S::STG + slot * S::STG_BYTES;
        uint32_t ph_f = 0u;
        struct LossIn { float4 a; float old_logp, adv, ret; bool valid; };
        auto take_rows = [&](int tile, int a0, LossIn& s) {
            mbar_wait(&bar_full[slot], ph_f); ph_f ^= 1u;
            const bool valid = tile * kM + tid < b.n;
            if (do_loss) {
                s.valid = valid;
                s.a = make_float4(0.f, 0.f, 0.f, 0.f);
                s.old_logp = 0.f; s.adv = 0.f; s.ret = 0.f;
                if (valid) {
                    if (b.packed) {
                        if (net == 0) s.a = *reinterpret_cast<const float4*>(stg + tid * 80 + 48);
                        const float4 sc4 = *reinterpret_cast<const float4*>(stg + tid * 80 + 64);
                        s.old_logp = sc4.x; s.adv = sc4.y; s.ret = sc4.z;
                    } else {
                        if (net == 0) s.a = *reinterpret_cast<const float4*>(stg + S::STG_ACT + tid * 16);
                        const float* sc = reinterpret_cast<const float*>(stg + S::STG_SCAL) + tid;
                        if (net == 0) { s.old_logp = sc[0]; s.adv = sc[128]; }
                        else s.ret = sc[256];
                    }
                }
            }
            if (do_a0) {
                // bf16 normalised observation, constant 1 in K slots 12 / 13 (0 for the padding rows of a ragged tile)
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, o2 = o0;
                if (valid) {
                    const float4* o = reinterpret_cast<const float4*>(stg + tid * (b.packed ? 80 : 48));
                    o0 = o[0]; o1 = o[1]; o2 = o[2];
                }
                const float o[kD] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x, o2.y, o2.z, o2.w};
                const float4* m4 = reinterpret_cast<const float4*>(sF + S::kMean);        // 6 x LDS.128: mean | inv_std
                const float4 m0 = m4[0], m1 = m4[1], m2 = m4[2], i0 = m4[3], i1 = m4[4], i2 = m4[5];
                const float mu[kD] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
                const float is[kD] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w};
                const float one = valid ? 1.0f : 0.f;
                float x[16];
#pragma unroll
                for (int k = 0; k < kD; ++k) x[k] = (o[k] - mu[k]) * is[k] * one;          // rows of padding samples are zero
                x[12] = one; x[13] = one; x[14] = 0.f; x[15] = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    *reinterpret_cast<uint4*>(sl + a0 + op_offset(128, tid, c)) =
                        make_uint4(pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                                   pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
            }
            mbar_arrive(&bar_empty[slot]);          // the staged values have been consumed (the arrive orders after the loads)
        };
        LossIn cur;
        take_rows(2 * cta + slot, S::A0, cur);
        signal();
#ifdef QS_PPO_PROFILE
        long long prof_[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long pc_ = clock64();
#endif
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            wait_done();                           // H1
            QS_PPOP(0);
            epilogue_relu(S::A1, false);
            signal();
            QS_PPOP(1);
            wait_done();                           // H2
            QS_PPOP(2);
            epilogue_relu(S::A2, true);
            signal();
            QS_PPOP(3);
            wait_done();                           // OUT
            QS_PPOP(4);
            if (do_loss) {
                float out[16];
                tmem_ld16(tw + 32u, out);
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                if (net == 0) {
                    const float a[4] = {cur.a.x, cur.a.y, cur.a.z, cur.a.w};
                    float z[4], logp = 0.f;
#pragma unroll
                    for (int k = 0; k < kA; ++k) {
                        z[k] = (a[k] - (out[k] + sF[S::kB3 + k])) * sF[S::kInvSig + k];
                        logp += -0.5f * z[k] * z[k] - sF[S::kLogStd + k] - 0.9189385332046727f;
                    }
                    const float lr = logp - cur.old_logp;
                    const float ratio = expf(lr);
                    const float A = (cur.adv - adv_mean) * adv_istd;
                    const float lo = 1.0f - hp.clip_range, hi = 1.0f + hp.clip_range;
                    const float unclipped = A * ratio, clipped = A * fminf(fmaxf(ratio, lo), hi);
                    const bool inside = ratio >= lo && ratio <= hi;
                    const bool active = inside || (unclipped < clipped);
                    const float g = (cur.valid && active) ? -A * ratio : 0.f;
#pragma unroll
                    for (int k = 0; k < kA; ++k) {
                        d[k] = g * z[k] * sF[S::kInvSig + k];
                        g_b3[k] += d[k];
                        g_ls[k] += g * (z[k] * z[k] - 1.0f);
                    }
                    if (cur.valid) {
                        st_loss += -fminf(unclipped, clipped);
                        st_clip += inside ? 0.f : 1.f;
                        st_kl += (ratio - 1.0f) - lr;
                        st_n += 1.f;
                    }
                } else {
                    const float err = out[0] + sF[S::kB3] - cur.ret;
                    d[0] = cur.valid ? 2.0f * hp.vf_coef * err : 0.f;
                    g_b3[0] += d[0];
                    if (cur.valid) st_loss += err * err;
                }
                *reinterpret_cast<uint4*>(sl + S::DOUT + op_offset(128, tid, 0)) =
                    make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), 0u, 0u);
                *reinterpret_cast<uint4*>(sl + S::DOUT + op_offset(128, tid, 1)) = make_uint4(0u, 0u, 0u, 0u);
            }
            signal();
            QS_PPOP(5);
            wait_done();                           // dH2 (and dW3: A2 may be overwritten)
            QS_PPOP(6);
            epilogue_mask_inplace(S::A2, true);    // D2 (+ its first four K steps in TMEM for dH1)
            signal();
            QS_PPOP(7);
            wait_done();                           // dH1 (and dW2, db2: A1 may be overwritten)
            QS_PPOP(8);
            epilogue_mask_inplace(S::A1, false);   // D1
            QS_PPOP(9);
            if (it + 1 < iters)                    // next tile's rows (copied a whole tile ago) -> loss inputs, A0 in the other buffer
                take_rows(2 * (cta + (it + 1) * ncta) + slot, S::A0 + ((it + 1) & 1) * 4096, cur);
            QS_PPOP(10);
            QS_PPOP(11);
            QS_PPOP(12);
            signal();                              // dW1 of this tile + H1 of the next one
            QS_PPOP(13);
        }
#ifdef QS_PPO_PROFILE
        if (cta == 0 && tid == 0)
            printf("ppoprof net %d slot %d half %d iters %d: waitH1 %lld | E1 %lld | waitH2 %lld | E2 %lld | waitOUT %lld | loss %lld | wait3 %lld | E4 %lld | wait4 %lld | E5 %lld | gather_wait %lld | write_a0 %lld | gather_async %lld | signal %lld (cycles/tile)\n",
                   net, slot, half, iters, prof_[0] / iters, prof_[1] / iters, prof_[2] / iters, prof_[3] / iters, prof_[4] / iters,
                   prof_[5] / iters, prof_[6] / iters, prof_[7] / iters, prof_[8] / iters, prof_[9] / iters, prof_[10] / iters,
                   prof_[11] / iters, prof_[12] / iters, prof_[13] / iters);
#endif
        wait_done();                               // the last commit: every accumulator is complete
    }
    fence_before();
    __syncthreads();
    fence_after();

    // ---- flush ---------------------------------------------------------------------------------------------------------
    const int P = L.total;
    float* out = partial + (size_t)cta * partial_stride(P);
    const float scale = 1.0f / (float)b.n;
    if (wg == 0) {
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        {
            float v0[16];
            tmem_ld16(tmem + lane_off + kQColW2, v0);         // lane = output feature n of layer 2; column 0: db2
            out[ob2 + tid] = v0[0] * scale;
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {           // dW2^T: column 16 + k = input feature k; a warp stores 32 consecutive n per k
            float v[32];
            tmem_ld32(tmem + lane_off + kQColW2 + 16u + (uint32_t)(c * 32), v);
#pragma unroll
            for (int k = 0; k < 32; ++k) out[oW2 + (c * 32 + k) * kH + tid] = v[k] * scale;
        }
        float v[16];
        tmem_ld16(tmem + lane_off + kQColW1, v);  // lane = hidden n, column = obs k | 12: bias
#pragma unroll
        for (int k = 0; k < kD; ++k) out[oW1 + k * kH + tid] = v[k] * scale;
        out[ob1 + tid] = v[kD] * scale;
        tmem_ld16(tmem + lane_off + kQColW3, v);  // lane = hidden k, column = head output
        if (net == 0) {
#pragma unroll
            for (int j = 0; j < kA; ++j) out[L.aW3 + tid * kA + j] = v[j] * scale;
        } else {
            out[L.cW3 + tid] = v[0] * scale;
        }
    }
    if (net == 0 && wg == 1) {
        for (int i = L.mean + tid; i < P; i += kM) out[i] = 0.f;             // the observation normaliser is not trained
        if (tid < 3) out[P + 5 + tid] = 0.f;
    }
    {
        float r[12] = {g_b3[0], g_b3[1], g_b3[2], g_b3[3], g_ls[0], g_ls[1], g_ls[2], g_ls[3], st_loss, st_clip, st_kl, st_n};
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
        }
        if (wg < 2 * kNH && (wg % kNH) == 0 && lane == 0) {      // the loss warpgroup of each slot: 2 x 4 warps
#pragma unroll
            for (int k = 0; k < 12; ++k) sRed[((wg / kNH) * 4 + warp) * 16 + k] = r[k];
        }
        __syncthreads();
        if (gtid < 12) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += sRed[w * 16 + gtid];
            if (net == 0) {
                if (gtid < 4) out[L.ab3 + gtid] = s * scale;
                // entropy bonus: H = sum_k (0.5 + 0.5 log 2pi + log_std_k) does not depend on the sample; CTA 0 carries it
                else if (gtid < 8) out[L.log_std + gtid - 4] = s * scale - (cta == 0 ? hp.ent_coef : 0.f);
                else if (gtid == 8) out[P + 0] = s;
                else out[P + gtid - 7] = s;                                     // 9 -> clipped (2), 10 -> KL (3), 11 -> n (4)
            } else {
                if (gtid == 0) out[L.cb3] = s * scale;
                else if (gtid == 8) out[P + 1] = s;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp_id == kIssuerWarp) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Minibatch shuffling (SB3 RolloutBuffer.get: one random permutation of the N rollout rows per epoch).  Instead of a
// sort-based randperm (~1 ms for 2^23 rows) every index is computed independently: a keyed 4-round balanced Feistel
// network is a bijection of [0, 4^h) (h = ceil(bits(N) / 2)), and cycle-walking it (re-encrypt until the value is < N)
// restricts it to a bijection of [0, N) -- a pseudo-random permutation evaluated in O(1) per element, memory-bound
// (4 B written per row).  Integer arithmetic only: bit-exact against oracle/ppo_update_ref.py: feistel_permutation.
__host__ __device__ inline uint32_t perm_mix(uint32_t x) {          // murmur3 finaliser
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
struct PermKeys { uint32_t k0; uint32_t rk[4]; };      // rk[round] = perm_mix(k1 + round), hoisted out of the kernel
__host__ __device__ inline uint32_t feistel_index(uint32_t i, uint32_t n, int half_bits, const PermKeys& K) {
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t x = i;
    do {
        uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
        for (uint32_t round = 0; round < 4; ++round) {
            const uint32_t f = perm_mix(r * 0x9E3779B1u + K.k0 + round * 0x7F4A7C15u) ^ K.rk[round];
            const uint32_t t = l ^ (f & mask);
            l = r; r = t;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}
// one application of the keyed Feistel bijection of [0, 4^half_bits) (feistel_index = this, repeated until < n)
__device__ __forceinline__ uint32_t feistel_once(uint32_t x, int half_bits, uint32_t mask, const PermKeys& K) {
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (uint32_t round = 0; round < 4; ++round) {
        const uint32_t f = perm_mix(r * 0x9E3779B1u + K.k0 + round * 0x7F4A7C15u) ^ K.rk[round];
        const uint32_t t = l ^ (f & mask);
        l = r; r = t;
    }
    return (l << half_bits) | r;
}
// out[i] = feistel_index(i) for a chunk of kPermChunk consecutive i per warp.  The cycle walk needs a geometric number of
// applications per element (2 on average when n is half the domain) -- one thread per element makes every warp wait for
// its unluckiest lane (~6 applications).  Here a lane whose element is finished immediately takes the next unstarted
// element of the warp's chunk (ballot + prefix count), so the lanes stay busy: 0.113 -> 0.062 ms for 2^23 elements.
constexpr int kPermChunk = 1024;
__global__ void __launch_bounds__(256)
ppo_permutation_kernel(uint32_t n, int half_bits, PermKeys K, int32_t* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * 256u + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const uint32_t lo = warp * (uint32_t)kPermChunk;
    if (lo >= n) return;
    const uint32_t hi = min(lo + (uint32_t)kPermChunk, n);
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t next = lo + 32u;                     // first element nobody has started (warp-uniform)
    uint32_t i = lo + lane, x = i;
    bool active = i < hi;
    while (__any_sync(0xffffffffu, active)) {
        bool finished = false;
        if (active) {
            x = feistel_once(x, half_bits, mask, K);
            if (x < n) { out[i] = (int32_t)x; finished = true; }
        }
        const uint32_t fin = __ballot_sync(0xffffffffu, finished);
        if (finished) {
            i = next + __popc(fin & ((1u << lane) - 1u));
            x = i;
            active = i < hi;
        }
        next += __popc(fin);
    }
}

// grad[e] = sum over CTAs of partial[c][e], fixed order (bitwise reproducible); e < len
__global__ void __launch_bounds__(256)
ppo_reduce_kernel(const float* __restrict__ partial, int rows_a, int rows_c, int crit_lo, int crit_hi, int crit_stat,
                  int stride, int len, float* __restrict__ grad) {
    // entries [crit_lo, crit_hi) and crit_stat were written by the rows_c critic CTAs, everything else by the rows_a actor CTAs
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= len) return;
    const int rows = ((e >= crit_lo && e < crit_hi) || e == crit_stat) ? rows_c : rows_a;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = 0;
    for (; r + 4 <= rows; r += 4) {
        s0 += partial[(size_t)r * stride + e]; s1 += partial[(size_t)(r + 1) * stride + e];
        s2 += partial[(size_t)(r + 2) * stride + e]; s3 += partial[(size_t)(r + 3) * stride + e];
    }
    for (; r < rows; ++r) s0 += partial[(size_t)r * stride + e];
    grad[e] = (s0 + s1) + (s2 + s3);
}

// mean and 1 / (std + 1e-8) of the minibatch's advantages (ddof 1: unbiased std, as torch.Tensor.std; ddof 0: population).  Double-precision block
// sums, one atomicAdd pair per block; the last block to finish writes the result and re-arms the scratch.
// scratch: 2 doubles (sums) + 1 unsigned (blocks done), zero before the first call.
// gridDim.y > 1: ALL minibatches of an epoch in one launch -- minibatch k = blockIdx.y covers idx[k mb_stride .. + n) (the last
// one up to n_total), with its own 32-byte scratch record scratch + 4 k and its own result adv_norm + 2 k.
__global__ void __launch_bounds__(256)
ppo_adv_stats_kernel(const float* __restrict__ adv, const int32_t* __restrict__ idx, int n, int ddof,
                     double* __restrict__ scratch, float* __restrict__ adv_norm, int mb_stride, int n_total) {
    __shared__ double sh[2][8];
    __shared__ bool last;
    if (gridDim.y > 1) {
        const int k = blockIdx.y;
        if (idx) idx += (size_t)k * mb_stride; else adv += (size_t)k * mb_stride;
        if (k + 1 == (int)gridDim.y) n = n_total - k * mb_stride;
        scratch += 4 * k; adv_norm += 2 * k;
    }
    double s = 0.0, q = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const double a = (double)adv[idx ? idx[i] : i];
        s += a; q += a * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, Q = 0.0;
        for (int w = 0; w < 8; ++w) { S += sh[0][w]; Q += sh[1][w]; }
        atomicAdd(&scratch[0], S); atomicAdd(&scratch[1], Q);
        __threadfence();
        unsigned* cnt = reinterpret_cast<unsigned*>(scratch + 2);
        last = atomicAdd(cnt, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            const double St = atomicAdd(&scratch[0], 0.0), Qt = atomicAdd(&scratch[1], 0.0);
            const double mean = St / n;
            // ddof 1: torch.Tensor.std (SB3); ddof 0: jnp.std (brax)
            const double var = n > ddof ? fmax(Qt - St * mean, 0.0) / (n - ddof) : 0.0;
            adv_norm[0] = (float)mean;
            adv_norm[1] = (float)(1.0 / (sqrt(var) + 1e-8));
            scratch[0] = 0.0; scratch[1] = 0.0; *cnt = 0u;
        }
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, max_grad_norm, grad_scale;
    float bias1, bias2;        // 1 - beta1^t, 1 - beta2^t
    int n_train;               // leading elements of the packed vector that are trained (everything before obs_mean)
};

// One CTA: global gradient norm (torch clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))) and the torch.optim.Adam
// step on the packed parameter vector (train.py:50-68 uses SB3's default optimiser: Adam, eps 1e-5, no weight decay).
// norm_out (optional): the pre-clip norm.
constexpr int kAdamPerThread = 40;        // 1024 threads x 40 >= the trained part of any supported policy (<= 40 960 floats)
// grid = ceil(n_train / 1024) CTAs.  Every CTA computes the global norm itself (the whole gradient is 150 KB: 40 loads in
// flight per thread, identical summation order in every CTA, no grid-wide synchronisation) and then updates its own 1024
// parameters.  A single CTA doing both passes with dependent trips to L2 took 43 us.
__global__ void __launch_bounds__(1024)
ppo_adam_kernel(AdamArgs a, float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ m,
                float* __restrict__ v, float* __restrict__ norm_out, const float* __restrict__ stats,
                float* __restrict__ stats_acc) {
    // (optional) running sums of the minibatch statistics: stats_acc[0..8) += stats[0..8)
    if (stats_acc && blockIdx.x == 0 && threadIdx.x < kPartialStats) stats_acc[threadIdx.x] += stats[threadIdx.x];
    __shared__ double sh[32];
    __shared__ float coef_s;
    float g[kAdamPerThread];
#pragma unroll
    for (int k = 0; k < kAdamPerThread; ++k) {
        const int i = threadIdx.x + k * 1024;
        g[k] = i < a.n_train ? __ldg(grad + i) * a.grad_scale : 0.f;
    }
    const int mine = blockIdx.x * 1024 + threadIdx.x;
    const bool ok = mine < a.n_train;
    const float mi0 = ok ? m[mine] : 0.f, vi0 = ok ? v[mine] : 0.f, p0 = ok ? params[mine] : 0.f;
    const float gmine = ok ? __ldg(grad + mine) * a.grad_scale : 0.f;
    double ss = 0.0;
#pragma unroll
    for (int k = 0; k < kAdamPerThread; ++k) ss += (double)g[k] * (double)g[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sh[w];
        const float norm = (float)sqrt(t);
        if (norm_out && blockIdx.x == 0) *norm_out = norm;
        float coef = 1.0f;
        if (a.max_grad_norm > 0.f) coef = fminf(1.0f, a.max_grad_norm / (norm + 1e-6f));
        coef_s = coef;
    }
    __syncthreads();
    if (ok) {
        const float gi = gmine * coef_s;
        const float mi = a.beta1 * mi0 + (1.0f - a.beta1) * gi;
        const float vi = a.beta2 * vi0 + (1.0f - a.beta2) * gi * gi;
        m[mine] = mi; v[mine] = vi;
        params[mine] = p0 - (a.lr / a.bias1) * mi / (sqrtf(vi) * rsqrtf(a.bias2) + a.eps);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-GPU: the update's one compute -> collective pair as ONE kernel over NVLink peer memory.  Every rank's
// ppo_reduce_kernel leaves its gradient (+ statistics) in a slot of its own CUDA-IPC-exported buffer; this kernel then
//   1. posts "my slot of epoch e is complete" into every peer's flag array (st.release.sys) and waits until all ranks
//      have posted (ld.acquire.sys, bounded spin -> trap instead of a hang),
//   2. sums the world's slots element-wise straight out of peer memory in rank order (identical order on every rank, so
//      the parameters stay bitwise in sync by construction -- no broadcast, no NCCL launch),
//   3. computes the global gradient norm (per-CTA partials -> device-wide counter barrier -> fixed-order sum) and
//   4. applies the Adam step to its 1024 parameters (every rank redundantly).
// Slots are double-buffered by epoch parity: a rank can only be one epoch ahead of a peer that still reads (it posts epoch
// e + 1 after finishing epoch e, and nobody passes the barrier of e + 1 before everybody has posted it).
struct PeerLayout {              // offsets (bytes) inside every rank's buffer; F = floats per slot (multiple of 32)
    int F;
    __host__ __device__ size_t slot(int parity) const { return (size_t)parity * F * sizeof(float); }
    __host__ __device__ size_t flags() const { return 2 * (size_t)F * sizeof(float); }                    // uint32[64]
    __host__ __device__ size_t partials(int parity) const { return flags() + 256 + (size_t)parity * 64 * sizeof(double); }
    __host__ __device__ size_t counters() const { return flags() + 256 + 2 * 64 * sizeof(double); }     // uint32[2] grid barriers | [4] error word
    __host__ __device__ size_t total() const { return counters() + 64; }
};
constexpr int kMaxPeers = 8;
struct PeerPtrs { unsigned char* base[kMaxPeers]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// grid = ceil(F / 1024) CTAs of 1024 threads; the in-kernel grid barrier needs them co-resident, which qs_ppo_comm_create
// verifies with an occupancy query.  stats_acc (optional, device): += the 8 summed statistics.
// A peer that does not arrive within `timeout_ns` (rank skew: lazy module load, a checkpoint on one rank, a debugger) does
// NOT kill the context: every CTA raises the error word of this rank's buffer and returns without touching the
// parameters; the host reads it with qs_ppo_comm_error.
__global__ void __launch_bounds__(1024)
ppo_peer_adam_kernel(AdamArgs a, PeerLayout L, PeerPtrs peers, int world, int rank, uint32_t epoch, int P,
                     float* __restrict__ params, float* __restrict__ m, float* __restrict__ v, float* __restrict__ norm_out,
                     float* __restrict__ stats_acc, unsigned long long timeout_ns) {
    __shared__ double sh[32];
    __shared__ float coef_s;
    __shared__ int abort_s;
    const int parity = (int)(epoch & 1u);
    unsigned char* self = peers.base[rank];
    uint32_t* counters = reinterpret_cast<uint32_t*>(self + L.counters());
    uint32_t* err_word = counters + 4;
    double* partials = reinterpret_cast<double*>(self + L.partials(parity));
    const unsigned long long t_start = globaltimer_ns();
    if (threadIdx.x == 0) abort_s = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[parity ^ 1] = 0u;        // re-arm the other epoch's grid barrier
    __syncthreads();
    // 1. post + wait
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + L.flags()) + rank, epoch);
    }
    if ((int)threadIdx.x < world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(self + L.flags()) + threadIdx.x;
        bool ok = false;
#pragma unroll 1
        for (;;) {
            if ((int32_t)(ld_acquire_sys(f) - epoch) >= 0) { ok = true; break; }
            if (*reinterpret_cast<volatile uint32_t*>(err_word) != 0u || globaltimer_ns() - t_start > timeout_ns) break;
        }
        if (!ok) { atomicExch(err_word, 1u + (uint32_t)threadIdx.x); abort_s = 1; }   // which peer never arrived
    }
    __syncthreads();
    if (abort_s) return;                                                             // whole CTA, before any state changes
    // 2. element-wise sum over the ranks' slots, rank order
    const int i = blockIdx.x * 1024 + threadIdx.x;
    float g = 0.f;
    if (i < L.F) {
        for (int p = 0; p < world; ++p) g += ld_relaxed_sys(reinterpret_cast<const float*>(peers.base[p] + L.slot(parity)) + i);
    }
    if (stats_acc && i >= P && i < P + kPartialStats) stats_acc[i - P] += g;
    g *= a.grad_scale;
    // 3. global norm: per-CTA partial -> counter barrier -> every CTA sums the partials in the same order
    double ss = (i < a.n_train) ? (double)g * (double)g : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sh[w];
        partials[blockIdx.x] = t;
        __threadfence();
        atomicAdd(&counters[parity], 1u);
        bool ok = false;
#pragma unroll 1
        for (;;) {
            if (*reinterpret_cast<volatile uint32_t*>(&counters[parity]) >= gridDim.x) { ok = true; break; }
            if (*reinterpret_cast<volatile uint32_t*>(err_word) != 0u || globaltimer_ns() - t_start > 2 * timeout_ns) break;
        }
        if (!ok) { atomicExch(err_word, 0x100u); abort_s = 1; }                      // a CTA of this grid bailed out / never ran
        __threadfence();
        double tot = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) tot += *reinterpret_cast<volatile double*>(&partials[b]);
        const float norm = (float)sqrt(tot);
        if (norm_out && blockIdx.x == 0) *norm_out = norm;
        coef_s = a.max_grad_norm > 0.f ? fminf(1.0f, a.max_grad_norm / (norm + 1e-6f)) : 1.0f;
    }
    __syncthreads();
    if (abort_s) return;
    // 4. Adam
    if (i < a.n_train) {
        const float gi = g * coef_s;
        const float mi = a.beta1 * m[i] + (1.0f - a.beta1) * gi;
        const float vi = a.beta2 * v[i] + (1.0f - a.beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        params[i] -= (a.lr / a.bias1) * mi / (sqrtf(vi) * rsqrtf(a.bias2) + a.eps);
    }
}

}  // namespace ppo
}  // namespace qs
