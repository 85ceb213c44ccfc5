// qs_rollout_tc.cuh -- tcgen05 / TMEM variant of the state-resident policy rollout (sm_100a only).
//
// Same contract as rollout_policy_kernel (qs_rollout.cuh; replaces acting.generate_unroll in
// ppo_train.train, train_brax_ppo.py:589-620, and SB3 collect_rollouts, train.py:133-137), but the three
// dense layers of the 2x128 actor and critic run on the 5th-generation tensor cores:
//
//   * a tile = 128 envs = one UMMA M-tile; within a tile thread i owns env i (state in registers) AND TMEM lane i, so after an
//     MMA every thread reads exactly its own env's activations with tcgen05.ld 32x32b.  Three CTA forms share this file and
//     record bitwise the same trajectories (tests/test_gpu_rollout.py): ONE tile + a partner warpgroup (256 threads; small
//     batches, where the per-step latency chain is everything: the partners take the critic, the noise and speculative reset
//     candidates off the owners' chain), TWO COMPACT tiles with a partner warpgroup each (512 threads, 256 TMEM columns per
//     tile; large 12-D batches: 16 warps per SM), and two plain tiles (256 threads; the A/B reference, QS_TC_FORM=2);
//   * operands are bf16 in shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte
//     core matrices, SBO = 128 B between row groups, LBO = rows/8 * 128 B between 16-byte K chunks), written
//     by the owners themselves: a warp's 16-byte stores for one K chunk are 512 contiguous bytes;
//   * accumulators are fp32 in TMEM (256 columns): L1 computes actor|critic together (N = 256, K = 16),
//     L2 is 8 + 8 instructions of M128 N128 K16, L3 is 8 + 8 of M128 N16 K16 (heads padded to 16 columns);
//   * the biases of layers 1 and 2 ride inside the MMAs: the padded K slots 12 and 13 of the layer-1 A operand carry a
//     constant 1, and the matching B rows carry the bias split into a bf16 high and a bf16 low part (hi + lo
//     reproduces the fp32 bias to 2^-17 relative); layer 2 gets one extra K = 16 step  A1 . B2bias  for the same
//     purpose.  Both epilogues are therefore just TMEM load -> ReLU + bf16 pack -> shared store;
//   * a single elected thread issues tcgen05.mma and tcgen05.commit -> mbarrier; everybody waits on the
//     barrier, then runs the epilogue (ReLU + bf16 pack -> next layer's A operand).
// fp32 weights arrive in the same packed vector as the FMA kernel and are converted to bf16 UMMA layout in
// shared memory once per CTA per launch.  Numerics: bf16 inputs, fp32 accumulation -- the oracle models
// exactly that rounding (oracle/ppo_ref.py forward(..., bf16=True)).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <type_traits>
#include "qs_kernels.cuh"
#include "qs_rollout.cuh"
#include "qs_umma_desc.cuh"

#ifndef QS_TC_PARTNER
#define QS_TC_PARTNER 1
#endif
#ifndef QS_TC_TS_HEADS
#define QS_TC_TS_HEADS 1     /* 1: relu(H2) stays in TENSOR MEMORY (tcgen05.st, in place over the consumed layer-2 accumulators) and the
                                head layer runs as TS-form MMAs (A from TMEM); 0: through shared memory (SS form) */
#endif
#ifndef QS_TC_SPLIT_CRITIC
#define QS_TC_SPLIT_CRITIC 1 /* 1 (partner-warpgroup CTAs): the critic's layer-2 / head MMAs and epilogues run BEHIND the actor's, off the
                                owners' per-step latency chain (own mbarrier, issued / consumed by the partner warpgroup) */
#endif
#ifndef QS_TC_PARTNER2
#define QS_TC_PARTNER2 1     /* 1: large 12-D batches run two COMPACT tiles per CTA, each with its own partner warpgroup (512 threads, 256
                                tensor-memory columns per tile: see the column plan in the kernel); 0: two plain tiles (256 threads) */
#endif
#ifndef QS_TC_SPEC_IN_FORWARD
#define QS_TC_SPEC_IN_FORWARD 1
#endif
#ifndef QS_TC_X_AFTER_HEAD
#define QS_TC_X_AFTER_HEAD 2 /* compact tiles: the critic's layer 2 is issued once the actor's head has been ISSUED (2: named barrier between the
                                two issuing warps), has COMPLETED (1: mbarrier barX), or right after the actor's layer 2 completed (0) */
#endif
#ifndef QS_TC_TS_L2
#define QS_TC_TS_L2 1        /* 1 (one-tile CTAs only: needs 384 of the 512 TMEM columns per tile): relu(H1) also stays in tensor
                                memory and layer 2 runs as TS-form MMAs */
#endif

namespace qs {
namespace tc {

// -DQS_TC_PROFILE: per-phase clock64 accounting of the step loop, printed by thread 0 of CTA 0 (tuning builds only)
#ifdef QS_TC_PROFILE
#define QS_TCP(k) do { const long long c_ = clock64(); prof_[k] += c_ - pc_; pc_ = c_; } while (0)
#else
#define QS_TCP(k) do { } while (0)
#endif

constexpr int kM = 128;                 // envs per tile = UMMA M
// TILES (template): independent 128-env tiles per CTA sharing the weights; 2 for large batches (2 warps per SMSP),
// 1 when the batch is too small to fill the SMs with 256-env CTAs
constexpr uint32_t kTileCols = 256;     // TMEM columns per tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// make_desc / make_idesc / idesc_mn: qs_umma_desc.cuh (host-testable bit packing)

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p; \n\t"
        "}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u));
}

// A operand from TENSOR MEMORY (TS form): lane = row, 32-bit column c holds K elements 2c (low half) and 2c + 1 of the
// row as bf16; a K = 16 step is 8 columns.  No shared-memory read for A: an N = 16 instruction then costs its 8-cycle
// tensor time instead of the ~64 cycles the 4 KB A fetch of the SS form takes (profiles/README.md, round 2).
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p; \n\t"
        "}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u));
}

// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers (SASS STTM)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tile-local barriers (128 threads each): the tiles of a CTA never wait for each other inside the step loop,
// so one tile's MMA phases overlap the other tile's epilogues / env steps
template <int NTHR = 128>
__device__ __forceinline__ void tile_sync(int tile) { asm volatile("bar.sync %0, %1;" :: "r"(tile + 1), "n"(NTHR) : "memory"); }
template <int NTHR = 128>
__device__ __forceinline__ bool tile_or(int tile, bool pred) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %2, 0;\n\tbar.red.or.pred p, %1, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(r) : "r"(tile + 1), "r"((uint32_t)pred), "n"(NTHR) : "memory");
    return r != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

// bounded wait: a tensor-core phase takes microseconds; if the barrier never flips something is wrong with
// the descriptors, and a trap is better than a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// relu + round-to-nearest bf16 + pack of two fp32 in ONE instruction (F2FP.RELU.BF16.F32.PACK_AB)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// tcgen05.ld without the wait (the registers are valid only after tmem_ld_wait on the same array)
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t r[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
// wait for every outstanding tcgen05.ld of this thread; the "+r" operands tie the register array to the wait so that
// no use of it can be scheduled above
__device__ __forceinline__ void tmem_ld_wait(uint32_t r[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t r[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t r[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint32_t pack_relu_bf16_u(uint32_t lo, uint32_t hi) {
    return pack_relu_bf16(__uint_as_float(lo), __uint_as_float(hi));
}

// one lane of a converged warp (warp-uniform control flow around it keeps tcgen05.mma free of per-lane election loops and
// lets ptxas hold the descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// two fp32 adds in one instruction (FADD2, Blackwell packed fp32)
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
    unsigned long long p, q;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}

// byte offset of (row, 8-element K chunk kc) in a K-major no-swizzle operand with `rows` rows
__device__ __forceinline__ uint32_t op_offset(int rows, int row, int kc) {
    return (uint32_t)((kc * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16);
}

// shared-memory map (bytes).  K1 = layer-1 K (observation + 2 bias slots, padded to a multiple of 16): 16 for the
// 12-D gym observation, 32 for the 21-D raw observation of the brax / mjx modes.
// COMPACT (two tiles per CTA, each with a partner warpgroup): the actor's activations never pass through shared memory
// (TS-form layers), so a tile keeps only the critic's layer-2 A operand (A2C); A2A has size 0.
template <int K1, bool PARTNER = false, bool COMPACT = false>
struct SmemT {
    static constexpr int W1 = 0;                          // B: [256 x K1] bf16
    static constexpr int W2A = W1 + 256 * K1 * 2;         // B: [128 x 128]
    static constexpr int W2C = W2A + 128 * 128 * 2;
    static constexpr int W3A = W2C + 128 * 128 * 2;       // B: [16 x 128]
    static constexpr int W3C = W3A + 16 * 128 * 2;
    static constexpr int B2A = W3C + 16 * 128 * 2;        // B: [128 x 16], two K rows = hi / lo halves of the layer-2 bias
    static constexpr int B2C = B2A + 128 * 16 * 2;
    static constexpr int WEND = B2C + 128 * 16 * 2;       // end of the (shared) B operands
    // per tile: A1 [128 x K1], A2A / A2C [128 x 128] (A2* are also the A operands of L3), and for the 21-D modes a
    // float staging tile for the coalesced trajectory store of the raw observations
    static constexpr int A1 = 0, A2A = 128 * K1 * 2, A2C = A2A + (COMPACT ? 0 : 128 * 128 * 2);
    static constexpr int OBS = A2C + 128 * 128 * 2;
    static constexpr int EPS = OBS + (K1 > 16 ? 128 * 21 * 4 : 0);      // PARTNER: 2 x [128] float4 sampling noise (double buffer)
    static constexpr int CAND = EPS + (PARTNER ? 2 * 128 * 16 : 0);     // PARTNER: speculative reset candidates [128][28] float + episode mailbox [128] u32
    static constexpr int kCandF = 28;                                   // p3 q4 v3 w3 target3 obs12
    static constexpr int EPI = CAND + (PARTNER ? 128 * kCandF * 4 : 0);
    static constexpr int VAL = EPI + (PARTNER ? 128 * 4 : 0);           // PARTNER: V(s) of the last forward, written by the partners
    static constexpr int SCR = VAL + (PARTNER ? 128 * 4 : 0);           // COMPACT: 4 x WarpResetScratch (576 B) of the owner warps
    static constexpr int TILE_BYTES = SCR + (COMPACT ? 4 * 576 : 0);
    static constexpr int TILE0 = WEND;
    __host__ __device__ static constexpr int f32_off(int tiles) { return TILE0 + tiles * TILE_BYTES; }   // fp32 constants, see below
    static constexpr int kB3 = 0 /*[32]*/, kLogStd = 32, kMean = 36, kInvStd = 60, kStd = 84 /*[4] exp(log_std)*/, kNumF = 88;
    // mbarriers: [tiles] the actor's / everybody's | tmem base (4 B, padded to 8) | PARTNER: [tiles] the critic's | COMPACT: [tiles]
    // "actor's layer 2 complete" for the partner warp that issues the critic's
    __host__ __device__ static constexpr int bar_off(int tiles) { return f32_off(tiles) + kNumF * 4; }
    __host__ __device__ static constexpr int total(int tiles) { return bar_off(tiles) + 24 * tiles + 8; }
};

// PARTNER (one tile per CTA, 12-D modes): a second warpgroup shares the tile.  Warp w + 4 reads the same 32 TMEM lanes
// as warp w, so the partners take the critic half of both epilogues (the owners keep the actor half), and while the
// owners sample, step and reset their envs the partners draw the NEXT step's Gaussian noise (Philox + Box-Muller)
// into a double-buffered shared array.  It shortens the per-step latency chain, which is all that matters when the
// batch is too small to give an SM more than one tile.
template <int MODE, int DIST, int TILES, bool PARTNER = false>
__global__ void __launch_bounds__(kM * TILES * (PARTNER ? 2 : 1), 1)
rollout_policy_tc_kernel(const __grid_constant__ QsParams P, Tables T, int n, float* __restrict__ state,
                         const float* __restrict__ params, int steps, uint32_t t0, int deterministic,
                         float bootstrap_gamma, RolloutBuffers rb, const float* __restrict__ first, int ept) {
    // ept = envs per tile (<= 128): the first `ept` rows of the 128-row UMMA tile carry envs, the rest are padding.
    // Spreading a small batch over more SMs this way (8192 envs: 147 CTAs x 56 envs instead of 64 CTAs x 128) was
    // MEASURED NOT TO HELP (profiles/README.md, round 2: 4.52 / 4.41 / 4.50 / 4.60 ms for ept 128 / 96 / 64 / 56 at
    // 8192 x 1024): a rollout is T sequential steps per CTA and the per-step latency chain of a CTA does not depend on
    // how many of its rows are live, so the launch takes T x chain whatever the grid.  The knob stays as a test /
    // tuning override (QS_TC_EPT); results are bitwise independent of it.
    constexpr int D = ModeTraits<MODE>::kObsDim;
    constexpr bool kGym = ModeTraits<MODE>::kGym;
    constexpr int K1 = (D + 2 <= 16) ? 16 : 32;            // observation + two bias slots
    constexpr int kS1 = K1 / 16;                           // layer-1 K = 16 steps
    constexpr int kBiasStep = D / 16, kBiasK = D % 16;     // K step / slot (and slot + 1) that carry the constant 1
    static_assert(D + 2 <= K1 && kBiasK + 1 < 16, "bias slots must fit one K = 16 step");
    // two tiles + two partner warpgroups (512 threads): the COMPACT tensor-memory plan below, 12-D observations only
    constexpr bool kCompact = PARTNER && TILES == 2;
    static_assert(!PARTNER || TILES == 1 || (TILES == 2 && K1 == 16 && QS_TC_TS_HEADS != 0), "partner warpgroups: one tile per CTA, or two compact tiles");
    using Smem = SmemT<K1, PARTNER, kCompact>;
    constexpr int Ao = DIST == 1 ? 2 * kA : kA;
    extern __shared__ __align__(1024) unsigned char smem[];
    // 256 accumulator columns per tile; one-tile CTAs with TS-form layer 2 also keep relu(H1) in columns [256, 384)
    constexpr uint32_t kTmemCols = (TILES == 1 && QS_TC_TS_L2 != 0 && QS_TC_TS_HEADS != 0) ? 512u : kTileCols * TILES;
    float* sF = reinterpret_cast<float*>(smem + Smem::f32_off(TILES));
    constexpr int kTT = PARTNER ? 2 * kM : kM;             // threads per tile
    constexpr int NT = kTT * TILES;
    const PolicyLayout L = policy_layout(D, DIST);
    const int gtid = threadIdx.x;                          // thread in the CTA (setup loops)
    const int tile = gtid / kTT, ltid = gtid % kTT;
    const int half = ltid / kM;                            // 0: owner warpgroup, 1: partner warpgroup
    const int tid = ltid % kM, warp = tid >> 5;            // row of the tile = TMEM lane; lane-quadrant warp index
    const int lwarp = __shfl_sync(0xffffffffu, ltid >> 5, 0);   // warp index within the tile, visibly warp-uniform: warp 0 issues the MMAs
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::bar_off(TILES)) + tile;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::bar_off(TILES) + 8 * TILES);
    uint64_t* barC = reinterpret_cast<uint64_t*>(smem + Smem::bar_off(TILES) + 8 * TILES + 8) + tile;          // PARTNER: the critic's commits
    uint64_t* barX = reinterpret_cast<uint64_t*>(smem + Smem::bar_off(TILES) + 16 * TILES + 8) + tile;         // COMPACT: the actor's layer 2 is complete
    const int tb = 4 * tile;                                // named barriers of this tile: tb + 1 whole tile, + 2 owners, + 3 partners, + 4 issue hand-off
    unsigned char* tsm = smem + Smem::TILE0 + tile * Smem::TILE_BYTES;   // this tile's A operands
    const int b0 = (blockIdx.x * TILES + tile) * ept;

    // ---- one-time setup: weights fp32 -> bf16 UMMA layout, constants, barrier, TMEM -------------------
    for (int idx = gtid; idx < (Smem::WEND) / 4; idx += NT) reinterpret_cast<uint32_t*>(smem)[idx] = 0u;   // zero all B operands
    __syncthreads();
    {
        // W1cat (n, k): n < 128 actor, n >= 128 critic; params store W[k][n]
        for (int idx = gtid; idx < D * 256; idx += NT) {
            const int k = idx / 256, nn = idx % 256;
            const float w = nn < 128 ? params[L.aW1 + k * kH + nn] : params[L.cW1 + k * kH + (nn - 128)];
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W1 + op_offset(256, nn, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(w);
        }
        // the layer-1 bias rides in the padded K slots D (bf16 high part) and D + 1 (bf16 low part); the A operand
        // carries a constant 1 in both.  The layer-2 bias uses the same two slots (of that K = 16 step) in B2A / B2C.
        for (int nn = gtid; nn < 256; nn += NT) {
            const float bv = nn < 128 ? params[L.ab1 + nn] : params[L.cb1 + (nn - 128)];
            const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
            const __nv_bfloat16 lo = __float2bfloat16_rn(bv - __bfloat162float(hi));
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W1 + op_offset(256, nn, D >> 3) + (D & 7) * 2) = hi;
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W1 + op_offset(256, nn, (D + 1) >> 3) + ((D + 1) & 7) * 2) = lo;
            const float b2 = nn < 128 ? params[L.ab2 + nn] : params[L.cb2 + (nn - 128)];
            const __nv_bfloat16 hi2 = __float2bfloat16_rn(b2);
            const __nv_bfloat16 lo2 = __float2bfloat16_rn(b2 - __bfloat162float(hi2));
            const int dst = nn < 128 ? Smem::B2A : Smem::B2C;
            *reinterpret_cast<__nv_bfloat16*>(smem + dst + op_offset(128, nn & 127, kBiasK >> 3) + (kBiasK & 7) * 2) = hi2;
            *reinterpret_cast<__nv_bfloat16*>(smem + dst + op_offset(128, nn & 127, (kBiasK + 1) >> 3) + ((kBiasK + 1) & 7) * 2) = lo2;
        }
        for (int idx = gtid; idx < kH * kH; idx += NT) {
            const int k = idx / kH, nn = idx % kH;
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W2A + op_offset(128, nn, k >> 3) + (k & 7) * 2) =
                __float2bfloat16_rn(params[L.aW2 + idx]);
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W2C + op_offset(128, nn, k >> 3) + (k & 7) * 2) =
                __float2bfloat16_rn(params[L.cW2 + idx]);
        }
        for (int idx = gtid; idx < kH * Ao; idx += NT) {
            const int k = idx / Ao, nn = idx % Ao;
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W3A + op_offset(16, nn, k >> 3) + (k & 7) * 2) =
                __float2bfloat16_rn(params[L.aW3 + idx]);
        }
        for (int k = gtid; k < kH; k += NT)
            *reinterpret_cast<__nv_bfloat16*>(smem + Smem::W3C + op_offset(16, 0, k >> 3) + (k & 7) * 2) =
                __float2bfloat16_rn(params[L.cW3 + k]);
        for (int idx = gtid; idx < Smem::kNumF; idx += NT) sF[idx] = 0.f;
    }
    __syncthreads();
    if (gtid < Ao) sF[Smem::kB3 + gtid] = params[L.ab3 + gtid];
    if (gtid == 0) sF[Smem::kB3 + 16] = params[L.cb3];
    if (DIST == 0 && gtid < kA) {
        sF[Smem::kLogStd + gtid] = params[L.log_std + gtid];
        sF[Smem::kStd + gtid] = expf(params[L.log_std + gtid]);          // once per launch instead of per env-step
    }
    if (gtid < D) { sF[Smem::kMean + gtid] = params[L.mean + gtid]; sF[Smem::kInvStd + gtid] = params[L.inv_std + gtid]; }
    if (ltid == 0) { mbar_init(bar, 1); if (PARTNER) mbar_init(barC, 1); if (kCompact) mbar_init(barX, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (gtid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    fence_async_smem();
    __syncthreads();
    fence_after();
    const uint32_t tmem_all = *tmem_slot;
    const uint32_t tmem = tmem_all + (uint32_t)tile * kTileCols;        // this tile's 256 accumulator columns
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 lanes

    const uint32_t sbase = smem_u32(smem);
    const uint32_t tbase = smem_u32(tsm);
    const uint32_t idesc_l1 = make_idesc(128, 256), idesc_l2 = make_idesc(128, 128), idesc_l3 = make_idesc(128, 16);
    // descriptors: LBO = (rows/8)*128 bytes between the two 16-byte K chunks of one instruction, SBO = 128
    const uint64_t dA1b = make_desc(tbase + Smem::A1 + kBiasStep * 4096, 16 * 128, 128);   // the K step with the constant-1 slots
    uint32_t phase = 0;

    const bool owner = half == 0 && tid < ept && (b0 + tid) < n;
    const uint32_t gid = P.env_id_offset + (uint32_t)(b0 + tid);
    float4* sEps = reinterpret_cast<float4*>(tsm + Smem::EPS);           // PARTNER only
    // standard-normal noise of (env gid, step index ts): Philox stream 1 + Box-Muller (fast intrinsics: the noise only
    // has to be N(0,1) to ~1e-6, it is not a state variable)
    auto draw_noise = [&](uint32_t ts) {
        const U4 r = philox4x32_10(U4{gid, ts, 0u, STREAM_POLICY}, P.philox_key);
        const float u0 = ((float)(r.x >> 8) + 1.0f) * 5.9604644775390625e-8f;
        const float u1 = (float)(r.y >> 8) * 5.9604644775390625e-8f;
        const float u2 = ((float)(r.z >> 8) + 1.0f) * 5.9604644775390625e-8f;
        const float u3 = (float)(r.w >> 8) * 5.9604644775390625e-8f;
        const float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
        float s0, c0, s1, c1;
        __sincosf(6.283185307179586f * u1, &s0, &c0);
        __sincosf(6.283185307179586f * u3, &s1, &c1);
        return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
    };
    if (PARTNER && half == 1) sEps[tid] = draw_noise(t0);               // noise of the first step
    // PARTNER, gym modes with Philox re-sampling: the partners also compute, every step and for every env, the state
    // the env WOULD be reset to (episode + 1) while the owners step their envs; a finished env then just picks its
    // candidate up from shared memory instead of running the ~310-instruction reset on the owners' critical path.
    float* sCand = reinterpret_cast<float*>(tsm + Smem::CAND);
    uint32_t* sEpi = reinterpret_cast<uint32_t*>(tsm + Smem::EPI);
    // compact tiles are the THROUGHPUT form (16 warps share the SM's issue slots): finished envs are re-sampled by their owner warps
    // (warp-cooperative, as in plain tiles).  Speculative candidates from the partners -- every step, or refreshed lazily only
    // after an env consumed its candidate -- were measured slower there (profiles/README.md: issue slots / 128-register cap).
    const bool spec_reset = PARTNER && !kCompact && kGym && P.auto_reset == QS_RESET_RESAMPLE && !P.waypoint_mode;   // CTA-uniform
    Env e;
    float obs_[D];
    if (owner) {
        load_env<MODE>(P, state, n, b0 + tid, e);
        float rpy[3] = {0.f, 0.f, 0.f};
        if constexpr (kGym) quat_to_rpy(e.b.q, rpy);
        compute_obs<MODE>(P, e, rpy, obs_);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) obs_[k] = 0.f;
    }
    if (spec_reset && half == 0) sEpi[tid] = owner ? e.episode : 0u;     // visible to the partners after the first tile barrier

#ifdef QS_TC_PROFILE
    long long prof_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long pc_ = clock64();
#endif
    // relu + bf16 pack epilogue of a hidden layer (the bias is already inside the MMA): TMEM -> A2A | A2C.  PARTNER: the
    // owners take the actor half, the partners the critic half.  The TMEM load of chunk i + 1 is in flight while chunk i
    // is converted and stored.
    // kDst: 0 = bf16 A operand in shared memory (A2A | A2C, SS-form consumer); 1 = tensor memory, IN PLACE over the first
    // half of each network's (consumed) accumulator columns: actor -> [0, 64), critic -> [128, 192) (TS-form head layer);
    // 2 = tensor memory columns [256, 320) | [320, 384) (TS-form layer 2, one-tile CTAs).  In-place is safe: a thread stores
    // the 16 packed columns of chunk c to [16 c, 16 c + 16) after it has loaded [32 c, 32 c + 32), the one load in flight
    // covers [32 (c + 1), 32 (c + 2)), and lanes are private to the warp that owns them.
    auto relu_epilogue = [&](auto dst_tag) {
        constexpr int kDst = decltype(dst_tag)::value;
        constexpr int kChunks = PARTNER ? 4 : 8;
        const int c_lo = PARTNER ? 4 * half : 0;
        uint32_t r[2][32];
        tmem_ld32_async(my_tmem + (uint32_t)(c_lo * 32), r[0]);
        tmem_ld_wait(r[0]);
#pragma unroll
        for (int i = 0; i < kChunks; ++i) {
            const int c = c_lo + i;
            if (i + 1 < kChunks) tmem_ld32_async(my_tmem + (uint32_t)((c + 1) * 32), r[(i + 1) & 1]);
            const uint32_t* v = r[i & 1];
            if constexpr (kDst == 0) {
                const int dst = (c < 4) ? Smem::A2A : Smem::A2C;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t* h = v + q * 8;
                    *reinterpret_cast<uint4*>(tsm + dst + op_offset(128, tid, (c & 3) * 4 + q)) =
                        make_uint4(pack_relu_bf16_u(h[0], h[1]), pack_relu_bf16_u(h[2], h[3]), pack_relu_bf16_u(h[4], h[5]),
                                   pack_relu_bf16_u(h[6], h[7]));
                }
            } else {
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) pk[q] = pack_relu_bf16_u(v[2 * q], v[2 * q + 1]);
                const uint32_t col = kDst == 1 ? (uint32_t)(c < 4 ? 16 * c : 128 + 16 * (c - 4))
                                               : (uint32_t)(256 + 16 * c);
                tmem_st16(my_tmem + col, pk);
            }
            if (i + 1 < kChunks) tmem_ld_wait(r[(i + 1) & 1]);
        }
        if constexpr (kDst != 0) tmem_st_wait();
    };
    // compact tiles (128 registers per thread): the same epilogue over 128 accumulator columns starting at `src`, in eight
    // 16-column chunks (32 registers in flight instead of 64).  to_tmem: packed bf16 -> tensor-memory columns [dst, dst + 64)
    // (in place when dst == src: chunk i is stored to [dst + 8 i, + 8) after [src + 16 i, + 16) was loaded, the load in flight
    // covers [src + 16 (i + 1), + 16)); else -> shared memory A2C.
    auto relu_epilogue_c = [&](bool to_tmem, uint32_t src, uint32_t dst) {
        uint32_t r[2][16];
        tmem_ld16_async(my_tmem + src, r[0]);
        tmem_ld_wait16(r[0]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i + 1 < 8) tmem_ld16_async(my_tmem + src + (uint32_t)((i + 1) * 16), r[(i + 1) & 1]);
            const uint32_t* v = r[i & 1];
            uint32_t pk[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) pk[q] = pack_relu_bf16_u(v[2 * q], v[2 * q + 1]);
            if (to_tmem) {
                tmem_st8(my_tmem + dst + (uint32_t)(8 * i), pk);
            } else {
                *reinterpret_cast<uint4*>(tsm + Smem::A2C + op_offset(128, tid, 2 * i)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(tsm + Smem::A2C + op_offset(128, tid, 2 * i + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            if (i + 1 < 8) tmem_ld_wait16(r[(i + 1) & 1]);
        }
        if (to_tmem) tmem_st_wait();
    };
    constexpr bool kTsHeads = QS_TC_TS_HEADS != 0;
    constexpr bool kTsL2 = QS_TC_TS_L2 != 0 && kTsHeads && TILES == 1;
    using DstSmem = std::integral_constant<int, 0>;
    using DstL3 = std::integral_constant<int, kTsHeads ? 1 : 0>;
    using DstL2 = std::integral_constant<int, kTsL2 ? 2 : 0>;
    // PARTNER + TS forms: the critic leaves the owners' per-step latency chain.  The env step needs the actor's head only;
    // V(s) goes to the trajectory (and to the rare timeout bootstrap).  So the layer-2 and head MMAs of the critic are
    // committed to their own mbarrier (barC): the owners wait for the actor's 9 layer-2 MMAs only, pack relu(H2) while the
    // critic's 9 run, issue the actor's head and go on to sample and step, while the partners -- who own the critic halves
    // of the epilogues anyway -- finish the critic behind them, issue its head themselves and leave V(s) in shared memory.
    // The owner / partner warpgroups synchronise among themselves (named barriers 2 / 3, 128 threads); both meet again at
    // the tile barrier that opens the next forward, before its layer-1 MMA overwrites the accumulator columns.
    constexpr bool kSplitCritic = PARTNER && TILES == 1 && QS_TC_SPLIT_CRITIC != 0 && QS_TC_TS_HEADS != 0 && QS_TC_TS_L2 != 0;
    // split-critic CTAs draw the reset candidates INSIDE the forward (behind the critic's layer-2 issue, while the owners wait for
    // the actor's layer 2 / head) rather than next to the owners' env step, where the two warps of a scheduler compete
    constexpr bool kSpecInForward = kSplitCritic && QS_TC_SPEC_IN_FORWARD != 0;
    constexpr bool kCriticBehind = kSplitCritic || kCompact;     // V(s) is produced (and stored) by the partner warpgroup
    uint32_t phaseC = 0, phaseX = 0;
    // speculative reset candidates (partner warpgroup, gym modes): the state env `tid` WOULD be reset to for its next episode
    int pending_noise_t = -1;               // kSpecInForward: step index whose noise the partners draw inside the coming forward
    auto spec_candidates = [&]() {
        if constexpr (PARTNER && kGym) {
            // speculative reset of env `tid` for its next episode (all 128 lanes busy, no divergence)
            Env r;
            r.episode = sEpi[tid] + 1u;
            r.wp_idx = 0; r.wp_reached = 0; r.laps = 0;
            float rpy[3], o[D];
            reset_env<MODE>(P, T, gid, r, rpy);
            compute_obs<MODE>(P, r, rpy, o);
            float4* d = reinterpret_cast<float4*>(sCand + tid * Smem::kCandF);
            d[0] = make_float4(r.b.p[0], r.b.p[1], r.b.p[2], r.b.q[0]);
            d[1] = make_float4(r.b.q[1], r.b.q[2], r.b.q[3], r.b.v[0]);
            d[2] = make_float4(r.b.v[1], r.b.v[2], r.b.w[0], r.b.w[1]);
            d[3] = make_float4(r.b.w[2], r.target[0], r.target[1], r.target[2]);
            d[4] = make_float4(o[0], o[1], o[2], o[3]);
            d[5] = make_float4(o[4], o[5], o[6], o[7]);
            d[6] = make_float4(o[8], o[9], o[10], o[11]);
        }
    };
    float* sVal = reinterpret_cast<float*>(tsm + Smem::VAL);
    auto group_sync = [&](int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(128) : "memory"); };
    // forward pass for the observation in `o`; returns head[Ao] and -- if need_value (split critic: otherwise the partners
    // keep it in sVal and store it themselves) -- value
    // noise_out (plain CTAs, may be null): this step's sampling noise is drawn UNDER the layer-2 MMAs -- Philox + Box-Muller
    // need only (env id, step index), and the threads would otherwise just wait there
    auto forward = [&](const float* o, float* head, float& value, bool need_value, float4* noise_out = nullptr, uint32_t noise_ts = 0u) {
      if constexpr (kCompact) {
        // COMPACT tiles (two tiles per CTA, each with a partner warpgroup; 256 tensor-memory columns per tile).  Column plan of a
        // forward (A = bf16 A operand of a TS-form MMA, two K elements per 32-bit column; D = fp32 accumulator):
        //   layer 1      D1 actor [0, 128) | D1 critic [128, 256)                    (one N = 256 instruction per K step)
        //   epilogue 1   owners:   relu(D1 actor)  -> A [0, 64)    in place
        //                partners: relu(D1 critic) -> shared memory A2C              (the critic's layer 2 is SS form)
        //   layer 2      actor  (warp 0): D2 [128, 256) <- A [0, 64) . W2A           on the owners' latency chain, TS form
        //                critic (warp 4): D2 [0, 128)   <- A2C . W2C                 issued only once the actor's HEAD is in the
        //                                  pipe (named barrier): its accumulator overwrites the actor's (dead) A columns
        //   epilogue 2   owners:   relu(D2 actor)  -> A [128, 192) in place;  head D [192, 208)
        //                partners: relu(D2 critic) -> A [0, 64)    in place;  head D [64, 80)
        // Everything the owners wait for is TS form, as in the one-tile split-critic CTAs; only the critic -- which runs behind
        // the owners' env step anyway -- pays a shared-memory hand-off, and the tile needs 256 columns instead of 384, so an SM
        // holds two tiles = 16 warps (plain two-tile CTAs: 8 warps, each tile's whole chain on one warpgroup).
        if (half == 0) {
            float x[K1];
#pragma unroll
            for (int k = 0; k < K1; ++k) x[k] = k < D ? (o[k] - sF[Smem::kMean + k]) * sF[Smem::kInvStd + k] : (k < D + 2 ? 1.0f : 0.f);
#pragma unroll
            for (int c = 0; c < K1 / 8; ++c)
                *reinterpret_cast<uint4*>(tsm + Smem::A1 + op_offset(128, tid, c)) =
                    make_uint4(pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                               pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
        }
        fence_async_smem();
        fence_before();
        tile_sync<kTT>(tb);                         // both warpgroups: the previous forward's TMEM reads are done too
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < kS1; ++j)
                    mma_bf16(tmem, make_desc(tbase + Smem::A1 + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W1 + j * 8192, 32 * 128, 128), idesc_l1, j > 0);
                mma_commit(bar);
                mma_commit(barC);
            }
            __syncwarp();
        }
        QS_TCP(0);
        if (half == 0) { mbar_wait(bar, phase); phase ^= 1; } else { mbar_wait(barC, phaseC); phaseC ^= 1; }
        fence_after();
        QS_TCP(1);
        if (half == 0) {
            relu_epilogue_c(true, 0u, 0u);          // actor: [0, 128) -> A [0, 64) in place
        } else {
            relu_epilogue_c(false, 128u, 0u);       // critic: [128, 256) -> A2C
            fence_async_smem();
        }
        QS_TCP(2);
        fence_before();
        tile_sync<kTT>(tb);
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    mma_bf16_ts(tmem + 128u, tmem + 8u * (uint32_t)j, make_desc(sbase + Smem::W2A + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
                mma_bf16(tmem + 128u, dA1b, make_desc(sbase + Smem::B2A, 16 * 128, 128), idesc_l2, 1u);
                mma_commit(bar);
                if (QS_TC_X_AFTER_HEAD == 0) mma_commit(barX);
            }
            __syncwarp();
        }
        if (lwarp == 4) {
            // issued only once the actor's HEAD is in the pipe: the critic's accumulator overwrites the A columns [0, 64) the
            // actor's layer 2 reads -- that layer has completed by then, the owners waited for it before their second epilogue
            // -- and the critic's nine SS-form MMAs must not sit in the in-order tensor pipe in front of the head MMAs the owners
            // are about to wait for (measured: "L3 sync + wait" 1240 vs 560 cycles per step)
            if (QS_TC_X_AFTER_HEAD == 2) asm volatile("bar.sync %0, 64;" :: "r"(tb + 4) : "memory");
            else { mbar_wait(barX, phaseX); phaseX ^= 1; }
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    mma_bf16(tmem, make_desc(tbase + Smem::A2C + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W2C + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
                mma_bf16(tmem, dA1b, make_desc(sbase + Smem::B2C, 16 * 128, 128), idesc_l2, 1u);
                mma_commit(barC);
            }
            __syncwarp();
        }
        if (half == 0) {
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
            QS_TCP(3);
            relu_epilogue_c(true, 128u, 128u);      // actor: relu(H2) in place -> A [128, 192)
            QS_TCP(4);
            fence_before();
            group_sync(tb + 2);
            if (lwarp == 0) {
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16_ts(tmem + 192u, tmem + 128u + 8u * (uint32_t)j, make_desc(sbase + Smem::W3A + j * 512, 2 * 128, 128), idesc_l3, j > 0);
                    mma_commit(bar);
                    if (QS_TC_X_AFTER_HEAD == 1) mma_commit(barX);               // the critic may start
                }
                __syncwarp();
                if (QS_TC_X_AFTER_HEAD == 2) asm volatile("bar.arrive %0, 64;" :: "r"(tb + 4) : "memory");
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
            QS_TCP(5);
            float v[16];
            tmem_ld16(my_tmem + 192u, v);
#pragma unroll
            for (int j = 0; j < Ao; ++j) head[j] = v[j] + sF[Smem::kB3 + j];
        } else {
            mbar_wait(barC, phaseC); phaseC ^= 1;
            fence_after();
            relu_epilogue_c(true, 0u, 0u);          // critic: relu(H2) in place -> A [0, 64)
            fence_before();
            group_sync(tb + 3);
            if (lwarp == 4) {
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16_ts(tmem + 64u, tmem + 8u * (uint32_t)j, make_desc(sbase + Smem::W3C + j * 512, 2 * 128, 128), idesc_l3, j > 0);
                    mma_commit(barC);
                }
                __syncwarp();
            }
            mbar_wait(barC, phaseC); phaseC ^= 1;
            fence_after();
            float w[16];
            tmem_ld16(my_tmem + 64u, w);
            value = w[0] + sF[Smem::kB3 + 16];
            sVal[tid] = value;
        }
        fence_before();       // the next forward's MMAs overwrite TMEM: order our loads before the coming barrier
        if (need_value) {     // (CTA-uniform per tile) the owners want V(s) now: timeout bootstrap, last value
            tile_sync<kTT>(tb);
            value = sVal[tid];
        }
        QS_TCP(6);
        return;
      }
      if constexpr (kSplitCritic) {
        if (half == 0) {
            float x[K1];
#pragma unroll
            for (int k = 0; k < K1; ++k) x[k] = k < D ? (o[k] - sF[Smem::kMean + k]) * sF[Smem::kInvStd + k] : (k < D + 2 ? 1.0f : 0.f);
#pragma unroll
            for (int c = 0; c < K1 / 8; ++c)
                *reinterpret_cast<uint4*>(tsm + Smem::A1 + op_offset(128, tid, c)) =
                    make_uint4(pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                               pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
        }
        fence_async_smem();
        fence_before();
        tile_sync<kTT>(tb);                       // both warpgroups: the previous forward's TMEM reads are done too
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < kS1; ++j)
                    mma_bf16(tmem, make_desc(tbase + Smem::A1 + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W1 + j * 8192, 32 * 128, 128), idesc_l1, j > 0);
                mma_commit(bar);
                mma_commit(barC);
            }
            __syncwarp();
        }
        QS_TCP(0);
        if (half == 0) { mbar_wait(bar, phase); phase ^= 1; } else { mbar_wait(barC, phaseC); phaseC ^= 1; }
        fence_after();
        QS_TCP(1);
        relu_epilogue(DstL2{});                     // owners: actor half -> [256, 320); partners: critic half -> [320, 384)
        QS_TCP(2);
        fence_before();
        tile_sync<kTT>(tb);
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    mma_bf16_ts(tmem, tmem + 256u + 8u * (uint32_t)j, make_desc(sbase + Smem::W2A + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
                mma_bf16(tmem, dA1b, make_desc(sbase + Smem::B2A, 16 * 128, 128), idesc_l2, 1u);
                mma_commit(bar);                    // the actor's layer 2 is all the owners wait for
            }
            __syncwarp();
            asm volatile("bar.arrive %0, 64;" :: "r"(tb + 4) : "memory");          // the actor's MMAs are in the pipe: the critic's may follow
        }
        // the critic's layer 2 is issued by a PARTNER warp (issuing blocks the thread at the pipe's pace: nine more MMAs
        // from warp 0 kept the owners' group barrier waiting for it), behind the actor's layer 2 (in-order pipe).  Holding it
        // back until the actor's HEAD is in the pipe too -- what the compact tiles do -- measured slower here (2.23e9 -> 2.10e9
        // at 8192 envs): with one tile per SM the pipe is mostly idle and the partners' chain is what gets longer.
        auto issue_critic_l2 = [&]() {              // warp 4
            asm volatile("bar.sync %0, 64;" :: "r"(tb + 4) : "memory");
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    mma_bf16_ts(tmem + 128u, tmem + 320u + 8u * (uint32_t)j, make_desc(sbase + Smem::W2C + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
                mma_bf16(tmem + 128u, dA1b, make_desc(sbase + Smem::B2C, 16 * 128, 128), idesc_l2, 1u);
                mma_commit(barC);
            }
            __syncwarp();
        };
        if (lwarp == 4) issue_critic_l2();
        if (kSpecInForward && half == 1) {
            if (spec_reset) spec_candidates();
            if (pending_noise_t >= 0) {             // the NEXT step's sampling noise (double-buffered: read one forward later)
                sEps[(pending_noise_t & 1) * kM + tid] = draw_noise(t0 + (uint32_t)pending_noise_t);
                pending_noise_t = -1;
            }
        }
        constexpr uint32_t kHeadA = 64u, kHeadC = 192u;
        if (half == 0) {
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
            QS_TCP(3);
            relu_epilogue(DstL3{});                 // actor: relu(H2) in place -> [0, 64)
            QS_TCP(4);
            fence_before();
            group_sync(tb + 2);
            if (lwarp == 0) {
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16_ts(tmem + kHeadA, tmem + 8u * (uint32_t)j, make_desc(sbase + Smem::W3A + j * 512, 2 * 128, 128), idesc_l3, j > 0);
                    mma_commit(bar);
                }
                __syncwarp();
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
            QS_TCP(5);
            float v[16];
            tmem_ld16(my_tmem + kHeadA, v);
#pragma unroll
            for (int j = 0; j < Ao; ++j) head[j] = v[j] + sF[Smem::kB3 + j];
        } else {
            mbar_wait(barC, phaseC); phaseC ^= 1;
            fence_after();
            relu_epilogue(DstL3{});                 // critic: relu(H2) in place -> [128, 192)
            fence_before();
            group_sync(tb + 3);
            if (lwarp == 4) {
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16_ts(tmem + kHeadC, tmem + 128u + 8u * (uint32_t)j, make_desc(sbase + Smem::W3C + j * 512, 2 * 128, 128), idesc_l3, j > 0);
                    mma_commit(barC);
                }
                __syncwarp();
            }
            mbar_wait(barC, phaseC); phaseC ^= 1;
            fence_after();
            float w[16];
            tmem_ld16(my_tmem + kHeadC, w);
            value = w[0] + sF[Smem::kB3 + 16];
            sVal[tid] = value;
        }
        fence_before();       // the next forward's MMAs overwrite TMEM: order our loads before the coming barrier
        if (need_value) {     // (CTA-uniform) the owners want V(s) now: timeout bootstrap, last value
            tile_sync<kTT>(tb);
            value = sVal[tid];
        }
        QS_TCP(6);
        return;
      }
        // A1: normalised obs, bf16, K padded D -> K1 with a constant 1 in slots D and D + 1
        if (half == 0) {
            float x[K1];
#pragma unroll
            for (int k = 0; k < K1; ++k) x[k] = k < D ? (o[k] - sF[Smem::kMean + k]) * sF[Smem::kInvStd + k] : (k < D + 2 ? 1.0f : 0.f);
#pragma unroll
            for (int c = 0; c < K1 / 8; ++c)
                *reinterpret_cast<uint4*>(tsm + Smem::A1 + op_offset(128, tid, c)) =
                    make_uint4(pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                               pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
        }
        fence_async_smem();
        fence_before();
        tile_sync<kTT>(tb);
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kS1; ++j)                                   // D1[128 x 256] = A1 . W1cat (+ b1)
                mma_bf16(tmem, make_desc(tbase + Smem::A1 + j * 4096, 16 * 128, 128),
                         make_desc(sbase + Smem::W1 + j * 8192, 32 * 128, 128), idesc_l1, j > 0);
            mma_commit(bar);
            }
            __syncwarp();
        }
        QS_TCP(0);
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();
        QS_TCP(1);
        // epilogue 1: h1 = relu(D1 + b1) -> bf16 A operand of layer 2: shared memory (A2A | A2C), or tensor memory columns
        // [256, 384) in one-tile CTAs  (PARTNER: owners take the actor half, partners the critic half)
        relu_epilogue(DstL2{});
        QS_TCP(2);
        if constexpr (!kTsL2) fence_async_smem();
        fence_before();
        tile_sync<kTT>(tb);
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {                                   // K = 128 in 8 steps of 16 (2 chunks of 2048 B each)
                if constexpr (kTsL2)
                    mma_bf16_ts(tmem, tmem + 256u + 8u * (uint32_t)j, make_desc(sbase + Smem::W2A + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
                else
                    mma_bf16(tmem, make_desc(tbase + Smem::A2A + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W2A + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if constexpr (kTsL2)
                    mma_bf16_ts(tmem + 128u, tmem + 320u + 8u * (uint32_t)j, make_desc(sbase + Smem::W2C + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
                else
                    mma_bf16(tmem + 128u, make_desc(tbase + Smem::A2C + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W2C + j * 4096, 16 * 128, 128), idesc_l2, j > 0);
            }
            // + b2: the layer-1 A operand (constant 1 in K slots 12 / 13) times the hi / lo bias rows
            mma_bf16(tmem, dA1b, make_desc(sbase + Smem::B2A, 16 * 128, 128), idesc_l2, 1u);
            mma_bf16(tmem + 128u, dA1b, make_desc(sbase + Smem::B2C, 16 * 128, 128), idesc_l2, 1u);
            mma_commit(bar);
            }
            __syncwarp();
        }
        if (noise_out) *noise_out = draw_noise(noise_ts);
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();
        QS_TCP(3);
        // epilogue 2: h2 = relu(D2) (b2 is inside D2) -> bf16 A operand of the head layer: tensor memory, in place over the
        // consumed accumulators (TS form), or shared memory A2A | A2C (the L2 MMAs have completed)
        relu_epilogue(DstL3{});
        QS_TCP(4);
        if constexpr (!kTsHeads) fence_async_smem();
        fence_before();
        tile_sync<kTT>(tb);
        // heads: N = 16 (B rows/8 = 2 -> LBO 256 B, K step 512 B); TS form: D at columns [64, 80) | [192, 208), the upper
        // (consumed) halves of the accumulator regions whose lower halves now hold the A operands
        constexpr uint32_t kHeadA = kTsHeads ? 64u : 0u, kHeadC = kTsHeads ? 192u : 16u;
        if (lwarp == 0) {
            fence_after();
            if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if constexpr (kTsHeads)
                    mma_bf16_ts(tmem + kHeadA, tmem + 8u * (uint32_t)j, make_desc(sbase + Smem::W3A + j * 512, 2 * 128, 128), idesc_l3, j > 0);
                else
                    mma_bf16(tmem + kHeadA, make_desc(tbase + Smem::A2A + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W3A + j * 512, 2 * 128, 128), idesc_l3, j > 0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if constexpr (kTsHeads)
                    mma_bf16_ts(tmem + kHeadC, tmem + 128u + 8u * (uint32_t)j, make_desc(sbase + Smem::W3C + j * 512, 2 * 128, 128), idesc_l3, j > 0);
                else
                    mma_bf16(tmem + kHeadC, make_desc(tbase + Smem::A2C + j * 4096, 16 * 128, 128),
                             make_desc(sbase + Smem::W3C + j * 512, 2 * 128, 128), idesc_l3, j > 0);
            }
            mma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();
        QS_TCP(5);
        if (half == 0) {
            if constexpr (kTsHeads) {
                float v[16], w[16];
                tmem_ld16(my_tmem + kHeadA, v);
                tmem_ld16(my_tmem + kHeadC, w);
#pragma unroll
                for (int j = 0; j < Ao; ++j) head[j] = v[j] + sF[Smem::kB3 + j];
                value = w[0] + sF[Smem::kB3 + 16];
            } else {
                float v[32];
                tmem_ld32(my_tmem, v);
#pragma unroll
                for (int j = 0; j < Ao; ++j) head[j] = v[j] + sF[Smem::kB3 + j];
                value = v[16] + sF[Smem::kB3 + 16];
            }
        }
        fence_before();       // the next forward's MMAs overwrite TMEM: order our loads before the coming barrier
        QS_TCP(6);
    };

    // trajectory store of this tile's observation rows.  12-D rows are three aligned float4 per thread; the 21-D rows
    // (84 B, unaligned) go through a shared staging tile and leave as one contiguous, fully coalesced span per tile.
    auto store_obs_rows = [&](float* dst_tile) {       // dst_tile = &out[(first env of the tile) * D]
        if constexpr (D == 12) {
            if (owner) {
                float4* dd = reinterpret_cast<float4*>(dst_tile + (size_t)tid * D);
                dd[0] = make_float4(obs_[0], obs_[1], obs_[2], obs_[3]);
                dd[1] = make_float4(obs_[4], obs_[5], obs_[6], obs_[7]);
                dd[2] = make_float4(obs_[8], obs_[9], obs_[10], obs_[11]);
            }
        } else {
            float* stage = reinterpret_cast<float*>(tsm + Smem::OBS);
            const int rows = min(ept, n - b0);
            if constexpr (PARTNER) {
                // owners only (their own named barrier): the partners are still behind, finishing the critic of the previous
                // step -- a tile-wide barrier here put the owners back behind them (ncu: `barrier` was the top stall)
                if (half == 0) {
#pragma unroll
                    for (int k = 0; k < D; ++k) stage[tid * D + k] = obs_[k];
                    group_sync(tb + 2);
                    for (int idx = tid; idx < rows * D; idx += kM) dst_tile[idx] = stage[idx];
                    group_sync(tb + 2);                          // the staging tile is rewritten next step
                }
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) stage[tid * D + k] = obs_[k];
                tile_sync<kTT>(tb);
                for (int idx = tid; idx < rows * D; idx += kM) dst_tile[idx] = stage[idx];
                tile_sync<kTT>(tb);                       // the staging tile is rewritten next step
            }
        }
    };

    for (int t = 0; t < steps; ++t) {
        const size_t o = (size_t)t * n + b0 + tid;
        if (rb.obs) store_obs_rows(rb.obs + ((size_t)t * n + b0) * D);
        QS_TCP(11);
        float head[Ao], value;
        float4 e4_pre = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kSpecInForward && t + 1 < steps) pending_noise_t = t + 1;
        forward(obs_, head, value, !kCriticBehind, PARTNER ? nullptr : &e4_pre, t0 + (uint32_t)t);
        if constexpr (kCriticBehind) {              // the partners computed V(s): they store it (thread tid <-> env tid, as the owners)
            if (half == 1 && tid < ept && (b0 + tid) < n && rb.value) rb.value[o] = value;
        }

        StepOut so;
        so.reward = 0.f; so.done = 0.f; so.truncated = 0.f; so.finished = false; so.needs_reset = false;
        float tobs[D];
        bool need_boot = false;
        if (owner) {
            // PARTNER: the partner warpgroup drew this step's noise during the previous step
            const float4 e4 = PARTNER ? sEps[(t & 1) * kM + tid] : e4_pre;      // plain CTAs: drawn under the layer-2 MMAs (forward)
            const float eps[4] = {e4.x, e4.y, e4.z, e4.w};
            float raw[4], act[4], logp = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = deterministic ? 0.f : eps[j];
                if constexpr (DIST == 0) {
                    const float ls = sF[Smem::kLogStd + j];
                    raw[j] = fmaf(sF[Smem::kStd + j], z, head[j]);
                    logp += -0.5f * z * z - ls - 0.9189385332046727f;
                    act[j] = clamp_(raw[j], -1.0f, 1.0f);
                } else {
                    const float scale = softplus_(head[kA + j]) + 0.001f;
                    raw[j] = fmaf(scale, z, head[j]);
                    const float ldj = 2.0f * (0.6931471805599453f - raw[j] - softplus_(-2.0f * raw[j]));
                    logp += -0.5f * z * z - logf(scale) - 0.9189385332046727f - ldj;
                    act[j] = tanhf(raw[j]);
                }
            }
            if (rb.act) reinterpret_cast<float4*>(rb.act)[o] = make_float4(raw[0], raw[1], raw[2], raw[3]);
            if (rb.logp) rb.logp[o] = logp;
            if (!kCriticBehind && rb.value) rb.value[o] = value;
            QS_TCP(7);
            env_step<MODE, true>(P, T, gid, e, act, obs_, tobs, first ? first + b0 + tid : nullptr, n, so);
            need_boot = bootstrap_gamma > 0.f && so.finished && so.truncated != 0.f && so.done == 0.f;
            if constexpr (ModeTraits<MODE>::kBrax) need_boot = bootstrap_gamma > 0.f && so.truncated != 0.f;
        }
        if (PARTNER && !kSpecInForward && half == 1 && t + 1 < steps)      // next step's noise, off the owners' critical path
            sEps[((t + 1) & 1) * kM + tid] = draw_noise(t0 + (uint32_t)(t + 1));
        bool any_boot = false;
        if constexpr (PARTNER && kGym) {
            if (spec_reset) {
                if (half == 1 && !kSpecInForward) spec_candidates();
                // candidates of this step are in place; the same tile-wide barrier carries the "somebody needs the timeout
                // bootstrap" vote of the step (one barrier per step fewer than a separate tile_or below)
                any_boot = tile_or<kTT>(tb, need_boot);
                if (so.needs_reset) {                                      // (owners only; e.episode was advanced by env_step)
                    const float4* d = reinterpret_cast<const float4*>(sCand + tid * Smem::kCandF);
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
                    sEpi[tid] = e.episode;                                 // read by the partner next step (3+ barriers later)
                }
            }
        }
        QS_TCP(8);
        // Philox re-sampling of finished envs, compacted per tile.  The scratch lives in this tile's A2A buffer,
        // which is idle between the head MMAs of this step and the first epilogue of the next one.
        if constexpr (kGym) {
          if (P.auto_reset == QS_RESET_RESAMPLE && half == 0 && !spec_reset) {
            if (P.waypoint_mode) {
                if (so.needs_reset) {
                    float rpy[3];
                    reset_env<MODE>(P, T, gid, e, rpy);
                    compute_obs<MODE>(P, e, rpy, obs_);
                }
            } else {
                // warp-cooperative Philox re-sampling (qs_kernels.cuh: warp_autoreset_smem), no tile barrier.  The
                // per-warp scratch lives at the start of this tile's A2A buffer, which is idle between the head MMAs
                // of this step and the first epilogue of the next forward (which every warp reaches only after the
                // tile barrier that follows the A1 store, i.e. after all warps have left this block).
                warp_autoreset_smem<MODE>(P, P.env_id_offset + (uint32_t)(b0 + warp * 32), e, obs_, so.needs_reset,
                                          reinterpret_cast<WarpResetScratch*>(tsm + (kCompact ? Smem::SCR : Smem::A2A))[warp]);
            }
          }
        }
        QS_TCP(9);
        // SB3 timeout bootstrap: reward += gamma * V(terminal_obs) for truncated-not-terminated episodes
        if (spec_reset ? any_boot : tile_or<kTT>(tb, need_boot)) {
            float h2[Ao], vt;
            forward(need_boot ? tobs : obs_, h2, vt, true);
            if (need_boot) so.reward = fmaf(bootstrap_gamma, vt, so.reward);
        }
        if (owner) {
            if (rb.reward) rb.reward[o] = so.reward;
            if (rb.done) rb.done[o] = so.done;
            if (rb.trunc) rb.trunc[o] = so.truncated;
        }
        QS_TCP(10);
    }
#ifdef QS_TC_PROFILE
    if (blockIdx.x == 0 && (gtid == 0 || gtid == kM - 1 || (PARTNER && gtid == kM)))
        printf("tcprof tid %d steps %d: A1+sync %lld | L1wait %lld | epi1 %lld | L2 sync+wait %lld | epi2 %lld | L3 sync+wait %lld | head %lld | sample %lld | env %lld | reset %lld | boot+store %lld | obs store %lld (cycles/step)\n",
               gtid, steps, prof_[0] / steps, prof_[1] / steps, prof_[2] / steps, prof_[3] / steps, prof_[4] / steps, prof_[5] / steps,
               prof_[6] / steps, prof_[7] / steps, prof_[8] / steps, prof_[9] / steps, prof_[10] / steps, prof_[11] / steps);
#endif

    {
        float head[Ao], value;
        forward(obs_, head, value, true);
        if (rb.last_obs) store_obs_rows(rb.last_obs + (size_t)b0 * D);
        if (owner) {
            if (rb.last_value) rb.last_value[b0 + tid] = value;
            store_env<MODE>(P, state, n, b0 + tid, e);
        }
    }
    fence_before();
    __syncthreads();
    if (gtid < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_all), "r"(kTmemCols) : "memory");
    }
}

template <int MODE, int DIST, int TILES, bool PARTNER = false>
inline int launch_rollout_tc_tt(const QsParams& P, const Tables& T, int n, float* state, const float* params, int steps,
                                uint32_t t0, const RolloutOpts& opt, const RolloutBuffers& rb, const float* first,
                                cudaStream_t s) {
    auto kern = rollout_policy_tc_kernel<MODE, DIST, TILES, PARTNER>;
    using Smem = SmemT<(ModeTraits<MODE>::kObsDim + 2 <= 16) ? 16 : 32, PARTNER, PARTNER && TILES == 2>;
    static_assert(Smem::total(TILES) <= 227 * 1024, "the CTA form does not fit the 227 KB of shared memory an sm_100 CTA can opt into");
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total(TILES));
    if (ce != cudaSuccess) return (int)ce;
    int ept = kM;                                    // envs per tile; QS_TC_EPT: test / tuning override (see the kernel)
    if (TILES == 1) {
        const char* ov = getenv("QS_TC_EPT");
        const int ept_env = ov ? atoi(ov) : 0;
        if (ept_env > 0 && ept_env <= kM) ept = ept_env;
    }
    kern<<<(n + ept * TILES - 1) / (ept * TILES), kM * TILES * (PARTNER ? 2 : 1), Smem::total(TILES), s>>>(
        P, T, n, state, params, steps, t0, opt.deterministic, opt.bootstrap_gamma, rb, first, ept);
    return 0;
}

template <int MODE, int DIST>
inline int launch_rollout_tc_t(const QsParams& P, const Tables& T, int n, float* state, const float* params, int steps,
                               uint32_t t0, const RolloutOpts& opt, const RolloutBuffers& rb, const float* first,
                               cudaStream_t s) {
    // two tiles per CTA only when that still gives every SM a CTA; the 21-D modes (K1 = 32 operands + the observation
    // staging tile) do not fit two tiles into 227 KB of shared memory
    // QS_TC_FORM (test / tuning override; every form records bitwise the same trajectories, tests/test_gpu_rollout.py):
    // 1 = one tile + partner warpgroup, 2 = two plain tiles, 3 = two compact tiles with a partner warpgroup each
    const char* fo = getenv("QS_TC_FORM");
    const int form = fo ? atoi(fo) : 0;
    if constexpr (ModeTraits<MODE>::kObsDim == 12) {
        const bool big = n >= 148 * 2 * kM;
#if QS_TC_PARTNER && QS_TC_PARTNER2 && QS_TC_TS_HEADS
        if (form == 3 || (form == 0 && big)) return launch_rollout_tc_tt<MODE, DIST, 2, true>(P, T, n, state, params, steps, t0, opt, rb, first, s);
#endif
        if (form == 2 || (form == 0 && big)) return launch_rollout_tc_tt<MODE, DIST, 2>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    }
#if QS_TC_PARTNER
    // one tile per CTA (small 12-D batches, every 21-D batch) plus a partner warpgroup that shortens the per-step latency chain
    return launch_rollout_tc_tt<MODE, DIST, 1, true>(P, T, n, state, params, steps, t0, opt, rb, first, s);
#else
    return launch_rollout_tc_tt<MODE, DIST, 1>(P, T, n, state, params, steps, t0, opt, rb, first, s);
#endif
}

inline int launch_rollout_policy_tc(const QsParams& P, const Tables& T, int n, float* state, const QsPolicyDesc& d,
                                    const float* params, int steps, uint32_t t0, const RolloutBuffers& rb,
                                    const float* first, cudaStream_t s) {
    RolloutOpts opt{d.deterministic, d.bootstrap_gamma};
    if (P.mode == QS_MODE_HOVER_GYM && d.dist == 0)
        return launch_rollout_tc_t<QS_MODE_HOVER_GYM, 0>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    if (P.mode == QS_MODE_HOVER_GYM && d.dist == 1)
        return launch_rollout_tc_t<QS_MODE_HOVER_GYM, 1>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    if (P.mode == QS_MODE_TRAJ_GYM && d.dist == 0)
        return launch_rollout_tc_t<QS_MODE_TRAJ_GYM, 0>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    // 21-D raw observation (train_brax_ppo.py:366-368): the Brax trainer's envs with its tanh-normal policy
    if (P.mode == QS_MODE_MJX_BRAX && d.dist == 1)
        return launch_rollout_tc_t<QS_MODE_MJX_BRAX, 1>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    if (P.mode == QS_MODE_HOVER_BRAX && d.dist == 1)
        return launch_rollout_tc_t<QS_MODE_HOVER_BRAX, 1>(P, T, n, state, params, steps, t0, opt, rb, first, s);
    return -100;
}

}  // namespace tc
}  // namespace qs
