// qs_ppo_generic.cuh -- the single-tile schedule of the PPO gradient, templated on observation size and action
// distribution.  (Included from qs_ppo.cuh; see there for the GEMM formulation.)
//
//   ppo_grad_tc_kernel<D, DIST>:  D = 12 (HoverEnv / TrajectoryFollowEnv observation) | 21 (raw [qpos, qvel] of the
//                                 MJX / Brax envs, train_brax_ppo.py:366-368);
//                                 DIST = 0 SB3 diagonal Gaussian with a state-independent log_std (train.py:61-64)
//                                      | 1 Brax NormalTanhDistribution, head = loc | raw_scale (train_brax_ppo.py:605-612)
//
// <12, 0> is the A/B reference of the production two-tile kernel (QS_PPO_V1=1); the other three instantiations are the
// product path for their policies ("Brax tanh-normal loss ... 21-D update").  Persistent CTAs, one
// 128-sample tile at a time, actor then critic through the same working columns; thread i = sample i = TMEM lane i.
//
// DIST = 1 loss (brax.training.agents.ppo.losses.compute_ppo_loss [third party, restated]; per minibatch):
//     scale = softplus(raw_scale) + 0.001          z = (a_raw - loc) / scale
//     logp  = sum_k [ -z^2 / 2 - log scale - log sqrt(2 pi) - ldj(a_raw) ],   ldj(x) = 2 (log 2 - x - softplus(-2 x))
//     L     = -mean(min(rho A, clip(rho, 1 - c, 1 + c) A)) + vf_coef mean((ret - V)^2) - ent_coef mean(H)
//     H     = sum_k [ 1/2 + log sqrt(2 pi) + log scale_k + ldj(loc_k + scale_k eps_k) ]     eps ~ N(0, 1) fresh per update
// (brax's value loss is 0.5 * 0.5 * mean(err^2): pass vf_coef = 0.25).  eps comes from Philox keyed by (sample_seed, row).
#pragma once

namespace qs {
namespace ppo {

template <int K1>
struct SmemPT {
    static constexpr int W1A = 0, W1C = 128 * K1 * 2;               // B: [128 x K1]
    static constexpr int W2A = 2 * 128 * K1 * 2, W2C = W2A + 32768;  // B: [128 x 128]
    static constexpr int W3A = W2C + 32768, W3C = W3A + 4096;        // B: [16 x 128]
    static constexpr int B2A = W3C + 4096, B2C = B2A + 4096;         // B: [128 x 16] bias rows (hi / lo in the constant-1 K slots)
    static constexpr int WEND = B2C + 4096;
    static constexpr int A0 = WEND;                                  // 2 x [128 x K1] (double buffered over tiles)
    static constexpr int A1 = A0 + 2 * 128 * K1 * 2;                 // [128 x 128] bf16 relu(H1)
    static constexpr int A2 = A1 + 32768;                            // [128 x 128] bf16 relu(H2); later D1
    static constexpr int D2 = A2 + 32768;                            // [128 x 128] bf16 masked dH2
    static constexpr int DOUT = D2 + 32768;                          // [128 x 16]
    static constexpr int F32 = DOUT + 4096;                          // fp32 constants
    static constexpr int kB3A = 0, kB3C = 8, kLogStd = 12, kInvSig = 16, kMean = 20, kInvStd = 44, kNumF = 68;
    static constexpr int RED = F32 + kNumF * 4;                      // [4 warps][24] block-reduction scratch
    static constexpr int BAR = RED + 4 * 24 * 4;
    static constexpr int TOTAL = BAR + 16;
};

struct HyperG {
    float clip_range, vf_coef, ent_coef;
    int normalize_adv;
    uint32_t sample_seed;        // DIST = 1: key of the entropy-term noise
};

enum : uint32_t { STREAM_PPO_ENTROPY = 3u };

template <int D, int DIST>
__global__ void __launch_bounds__(kM, 1)
ppo_grad_tc_kernel(Batch b, HyperG hp, const float* __restrict__ params, const float* __restrict__ adv_norm,
                   float* __restrict__ partial, int mn_swap) {
    constexpr int K1 = (D + 2 <= 16) ? 16 : 32;             // observation + two constant-1 slots, padded to K = 16 steps
    constexpr int kS1 = K1 / 16;
    constexpr int kBiasStep = D / 16, kBiasK = D % 16;      // K = 16 step / slot (and slot + 1) carrying the constant 1
    static_assert(kBiasK + 1 < 16, "the two bias slots must sit in one K = 16 step");
    constexpr int Ao = DIST == 1 ? 2 * kA : kA;
    using S = SmemPT<K1>;
    // TMEM columns: 128 working + per network {dW2 128, dW1^T|db1 K1, dW3 16, db2 16}
    constexpr uint32_t kNet = 128 + K1 + 32, cW = 0;
    constexpr uint32_t kCols = 128 + 2 * kNet <= 256 ? 256 : 512;
    static_assert(128 + 2 * kNet <= 512, "accumulators must fit the 512 TMEM columns");
    extern __shared__ __align__(1024) unsigned char smem[];
    const PolicyLayout L = policy_layout(D, DIST);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // warp index, visibly warp-uniform (the shuffle tells ptxas): the issuing warp's descriptors then live in uniform registers
    // and its tcgen05.mma need no per-lane election loops (a plain `if (tid == 0)` cost 130-240 cycles per MMA of issue)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    float* sF = reinterpret_cast<float*>(smem + S::F32);
    float* sRed = reinterpret_cast<float*>(smem + S::RED);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 8);

    // ---- one-time setup: fp32 weights -> bf16 UMMA operands (same layouts as the rollout kernel) -------------------
    for (int i = tid; i < S::WEND / 4; i += kM) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    __syncthreads();
    auto put = [&](int base, int rows, int row, int k, float w) {
        *reinterpret_cast<__nv_bfloat16*>(smem + base + op_offset(rows, row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(w);
    };
    for (int i = tid; i < D * kH; i += kM) {
        const int k = i / kH, n = i % kH;
        put(S::W1A, 128, n, k, params[L.aW1 + i]);
        put(S::W1C, 128, n, k, params[L.cW1 + i]);
    }
    for (int n = tid; n < kH; n += kM) {
        // hidden biases as bf16 hi + lo in the two constant-1 K slots: layer 1 inside W1 (slots D, D + 1), layer 2 in the
        // [128 x 16] bias operand that multiplies A0's K step `kBiasStep` (slots kBiasK, kBiasK + 1)
        const float bv[4] = {params[L.ab1 + n], params[L.cb1 + n], params[L.ab2 + n], params[L.cb2 + n]};
        const int dst[4] = {S::W1A, S::W1C, S::B2A, S::B2C};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(bv[q]);
            const float lo = bv[q] - __bfloat162float(hi);
            const int k0 = q < 2 ? D : kBiasK;
            put(dst[q], 128, n, k0, __bfloat162float(hi));
            put(dst[q], 128, n, k0 + 1, lo);
        }
    }
    for (int i = tid; i < kH * kH; i += kM) {
        const int k = i / kH, n = i % kH;
        put(S::W2A, 128, n, k, params[L.aW2 + i]);
        put(S::W2C, 128, n, k, params[L.cW2 + i]);
    }
    for (int i = tid; i < kH * Ao; i += kM) put(S::W3A, 16, i % Ao, i / Ao, params[L.aW3 + i]);
    for (int k = tid; k < kH; k += kM) put(S::W3C, 16, 0, k, params[L.cW3 + k]);
    if (tid < Ao) sF[S::kB3A + tid] = params[L.ab3 + tid];
    if (DIST == 0 && tid < kA) {
        const float ls = params[L.log_std + tid];
        sF[S::kLogStd + tid] = ls;
        sF[S::kInvSig + tid] = expf(-ls);
    }
    if (tid == 0) sF[S::kB3C] = params[L.cb3];
    if (tid < D) { sF[S::kMean + tid] = params[L.mean + tid]; sF[S::kInvStd + tid] = params[L.inv_std + tid]; }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    fence_async_smem();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t sb = smem_u32(smem);
    // descriptors.  K-major [rows x K]: LBO = rows/8 * 128 between the two 16-byte K chunks, SBO = 128 between 8-row
    // groups.  MN-major view of a K-major [128 x C] buffer (MN = the buffer's columns, K = its rows): 8-column groups are
    // 2048 B apart (SBO), 8-row groups 128 B (LBO); a K = 16 step advances the start address by 256 B.
    auto dk = [&](int off, int rows) { return make_desc(sb + off, (uint32_t)(rows / 8) * 128u, 128u); };
    auto dmn = [&](int off, uint32_t grp_stride) {
        return mn_swap ? make_desc(sb + off, grp_stride, 128u) : make_desc(sb + off, 128u, grp_stride);
    };
    const uint32_t id_kk128 = idesc_mn(128, 128, 0, 0), id_kk16 = idesc_mn(128, 16, 0, 0);
    const uint32_t id_kmn128 = idesc_mn(128, 128, 0, 1);
    const uint32_t id_mm128 = idesc_mn(128, 128, 1, 1), id_mm16 = idesc_mn(128, 16, 1, 1), id_mmK1 = idesc_mn(128, K1, 1, 1);
    uint32_t phase = 0;

    const float adv_mean = hp.normalize_adv ? adv_norm[0] : 0.f;
    const float adv_istd = hp.normalize_adv ? adv_norm[1] : 1.f;
    float g_b3a[Ao], g_ls[kA] = {0.f, 0.f, 0.f, 0.f}, g_b3c = 0.f;
#pragma unroll
    for (int k = 0; k < Ao; ++k) g_b3a[k] = 0.f;
    float st_pg = 0.f, st_v = 0.f, st_clip = 0.f, st_kl = 0.f, st_n = 0.f, st_ent = 0.f;

    auto handoff = [&]() { fence_async_smem(); fence_before(); __syncthreads(); };
    auto wait_phase = [&]() { mbar_wait(bar, phase); phase ^= 1; fence_after(); };
    auto epilogue_relu = [&](int dst) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(tmem + lane_off + cW + (uint32_t)(c * 32), v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* h = v + q * 8;
                *reinterpret_cast<uint4*>(smem + dst + op_offset(128, tid, c * 4 + q)) =
                    make_uint4(pack_relu_bf16(h[0], h[1]), pack_relu_bf16(h[2], h[3]), pack_relu_bf16(h[4], h[5]),
                               pack_relu_bf16(h[6], h[7]));
            }
        }
    };
    auto epilogue_mask = [&](int act, int dst) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(tmem + lane_off + cW + (uint32_t)(c * 32), v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* g = v + q * 8;
                const uint32_t off = op_offset(128, tid, c * 4 + q);
                const uint4 h = *reinterpret_cast<const uint4*>(smem + act + off);
                *reinterpret_cast<uint4*>(smem + dst + off) =
                    make_uint4(pack_mask_bf16(g[0], g[1], h.x), pack_mask_bf16(g[2], g[3], h.y),
                               pack_mask_bf16(g[4], g[5], h.z), pack_mask_bf16(g[6], g[7], h.w));
            }
        }
    };

    struct Row { float o[D]; float a[4]; float old_logp, adv, ret; int j; bool valid; };
    auto load_row = [&](int tile, Row& s) {
        const int r = tile * kM + tid;
        s.valid = r < b.n;
        s.j = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) s.o[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s.a[k] = 0.f;
        s.old_logp = 0.f; s.adv = 0.f; s.ret = 0.f;
        if (s.valid) {
            const size_t j = b.idx ? (size_t)b.idx[r] : (size_t)r;
            s.j = (int)j;
            if (b.packed) {
                // one 128-byte line per sample: obs[D] | act[4] | old_logp | adv | ret
                const float4* pr = reinterpret_cast<const float4*>(b.packed + j * kRowF);
                float f[(D + 7 + 3) / 4 * 4];
#pragma unroll
                for (int c = 0; c < (D + 7 + 3) / 4; ++c) { const float4 v = __ldg(pr + c); f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w; }
#pragma unroll
                for (int k = 0; k < D; ++k) s.o[k] = f[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) s.a[k] = f[D + k];
                s.old_logp = f[D + 4]; s.adv = f[D + 5]; s.ret = f[D + 6];
            } else if constexpr (D % 4 == 0) {
                const float4* o = reinterpret_cast<const float4*>(b.obs + j * D);
#pragma unroll
                for (int c = 0; c < D / 4; ++c) { const float4 v = __ldg(o + c); s.o[4 * c] = v.x; s.o[4 * c + 1] = v.y; s.o[4 * c + 2] = v.z; s.o[4 * c + 3] = v.w; }
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) s.o[k] = __ldg(b.obs + j * D + k);
            }
            if (!b.packed) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(b.act) + j);
                s.a[0] = a4.x; s.a[1] = a4.y; s.a[2] = a4.z; s.a[3] = a4.w;
                s.old_logp = __ldg(b.old_logp + j); s.adv = __ldg(b.adv + j); s.ret = __ldg(b.ret + j);
            }
        }
    };

    const int ntiles = (b.n + kM - 1) / kM;
    Row cur, nxt;
    load_row(blockIdx.x, cur);
    nxt = cur;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        if (tile + (int)gridDim.x < ntiles) load_row(tile + gridDim.x, nxt);    // in flight during this tile
        const uint32_t first = (it == 0) ? 0u : 1u;          // accumulate flag of the first K step into a gradient accumulator
        const int a0 = S::A0 + (it & 1) * (128 * K1 * 2);
        {
            // A0: bf16 normalised observation, constant 1 in K slots D / D + 1 (0 for the padding rows of a ragged tile)
            float x[K1];
#pragma unroll
            for (int k = 0; k < K1; ++k)
                x[k] = !cur.valid ? 0.f : (k < D ? (cur.o[k < D ? k : 0] - sF[S::kMean + (k < D ? k : 0)]) * sF[S::kInvStd + (k < D ? k : 0)]
                                                 : (k < D + 2 ? 1.0f : 0.f));
#pragma unroll
            for (int c = 0; c < K1 / 8; ++c)
                *reinterpret_cast<uint4*>(smem + a0 + op_offset(128, tid, c)) =
                    make_uint4(pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                               pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
        }
#pragma unroll 1
        for (int net = 0; net < 2; ++net) {
            const int W1 = net ? S::W1C : S::W1A, W2 = net ? S::W2C : S::W2A, W3 = net ? S::W3C : S::W3A;
            const int B2 = net ? S::B2C : S::B2A;
            const uint32_t base = 128u + (uint32_t)net * kNet;
            const uint32_t cW2 = base, cW1 = base + 128u, cW3 = base + 128u + (uint32_t)K1, cB2 = cW3 + 16u;
            // ---- forward -------------------------------------------------------------------------------------------
            handoff();
            if (warp_u == 0) {                  // warp-uniform branch + elect.sync: no per-lane election loop around the MMAs
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < kS1; ++j)
                        mma_bf16(tmem + cW, dk(a0 + j * 4096, 128), dk(W1 + j * 4096, 128), id_kk128, j > 0);
                    mma_commit(bar);
                }
                __syncwarp();
            }
            wait_phase();
            epilogue_relu(S::A1);
            handoff();
            if (warp_u == 0) {                  // warp-uniform branch + elect.sync: no per-lane election loop around the MMAs
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16(tmem + cW, dk(S::A1 + j * 4096, 128), dk(W2 + j * 4096, 128), id_kk128, j > 0);
                    mma_bf16(tmem + cW, dk(a0 + kBiasStep * 4096, 128), dk(B2, 128), id_kk128, 1u);
                    mma_commit(bar);
                }
                __syncwarp();
            }
            wait_phase();
            epilogue_relu(S::A2);
            handoff();
            if (warp_u == 0) {                  // warp-uniform branch + elect.sync: no per-lane election loop around the MMAs
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16(tmem + cW, dk(S::A2 + j * 4096, 128), dk(W3 + j * 512, 16), id_kk16, j > 0);
                    mma_commit(bar);
                }
                __syncwarp();
            }
            wait_phase();
            // ---- loss gradient w.r.t. the head outputs (fp32, this thread's sample) --------------------------------
            {
                float out[16];
                tmem_ld16(tmem + lane_off + cW, out);
                float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (net == 0) {
                    const float A = (cur.adv - adv_mean) * adv_istd;
                    const float lo = 1.0f - hp.clip_range, hi = 1.0f + hp.clip_range;
                    float z[4], logp = 0.f;
                    if constexpr (DIST == 0) {
#pragma unroll
                        for (int k = 0; k < kA; ++k) {
                            z[k] = (cur.a[k] - (out[k] + sF[S::kB3A + k])) * sF[S::kInvSig + k];
                            logp += -0.5f * z[k] * z[k] - sF[S::kLogStd + k] - 0.9189385332046727f;
                        }
                        const float lr = logp - cur.old_logp;
                        const float ratio = expf(lr);
                        const float unclipped = A * ratio, clipped = A * fminf(fmaxf(ratio, lo), hi);
                        const bool inside = ratio >= lo && ratio <= hi;
                        const bool active = inside || (unclipped < clipped);
                        const float g = (cur.valid && active) ? -A * ratio : 0.f;          // d loss_i / d logp_i (unscaled)
#pragma unroll
                        for (int k = 0; k < kA; ++k) {
                            d[k] = g * z[k] * sF[S::kInvSig + k];
                            g_b3a[k] += d[k];
                            g_ls[k] += g * (z[k] * z[k] - 1.0f);
                        }
                        if (cur.valid) {
                            st_pg += -fminf(unclipped, clipped);
                            st_clip += inside ? 0.f : 1.f;
                            st_kl += (ratio - 1.0f) - lr;
                            st_n += 1.f;
                        }
                    } else {
                        // entropy-term noise of this row: Philox(sample_seed; row, 0, 0, stream 3) + Box-Muller
                        const U4 rn = philox4x32_10(U4{(uint32_t)cur.j, 0u, 0u, STREAM_PPO_ENTROPY}, hp.sample_seed, 0x5eed0ea7u);
                        const float u0 = ((float)(rn.x >> 8) + 1.0f) * 5.9604644775390625e-8f, u1 = (float)(rn.y >> 8) * 5.9604644775390625e-8f;
                        const float u2 = ((float)(rn.z >> 8) + 1.0f) * 5.9604644775390625e-8f, u3 = (float)(rn.w >> 8) * 5.9604644775390625e-8f;
                        const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
                        float s0, c0, s1, c1;
                        sincosf(6.283185307179586f * u1, &s0, &c0);
                        sincosf(6.283185307179586f * u3, &s1, &c1);
                        const float eps[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
                        float scale[4], isc[4], sig[4], ent = 0.f;
#pragma unroll
                        for (int k = 0; k < kA; ++k) {
                            const float loc = out[k] + sF[S::kB3A + k], rs = out[kA + k] + sF[S::kB3A + kA + k];
                            scale[k] = softplus_(rs) + 0.001f;
                            isc[k] = 1.0f / scale[k];
                            sig[k] = 1.0f / (1.0f + expf(-rs));
                            z[k] = (cur.a[k] - loc) * isc[k];
                            const float ldj_a = 2.0f * (0.6931471805599453f - cur.a[k] - softplus_(-2.0f * cur.a[k]));
                            logp += -0.5f * z[k] * z[k] - logf(scale[k]) - 0.9189385332046727f - ldj_a;
                            // entropy sample x = loc + scale eps: H_k = 1/2 + log sqrt(2 pi) + log scale + ldj(x); ldj'(x) = -2 tanh x
                            const float xs = fmaf(scale[k], eps[k], loc);
                            const float th = tanhf(xs);
                            ent += 1.4189385332046727f + logf(scale[k]) + 2.0f * (0.6931471805599453f - xs - softplus_(-2.0f * xs));
                            const float e = cur.valid ? -hp.ent_coef : 0.f;                 // d L / d H (unscaled)
                            d[k] = e * (-2.0f * th);
                            d[kA + k] = e * (isc[k] - 2.0f * th * eps[k]) * sig[k];
                        }
                        const float lr = logp - cur.old_logp;
                        const float ratio = expf(lr);
                        const float unclipped = A * ratio, clipped = A * fminf(fmaxf(ratio, lo), hi);
                        const bool inside = ratio >= lo && ratio <= hi;
                        const bool active = inside || (unclipped < clipped);
                        const float g = (cur.valid && active) ? -A * ratio : 0.f;
#pragma unroll
                        for (int k = 0; k < kA; ++k) {
                            d[k] += g * z[k] * isc[k];
                            d[kA + k] += g * (z[k] * z[k] - 1.0f) * isc[k] * sig[k];
                        }
#pragma unroll
                        for (int k = 0; k < Ao; ++k) g_b3a[k] += d[k];
                        if (cur.valid) {
                            st_pg += -fminf(unclipped, clipped);
                            st_clip += inside ? 0.f : 1.f;
                            st_kl += (ratio - 1.0f) - lr;
                            st_n += 1.f;
                            st_ent += ent;
                        }
                    }
                } else {
                    const float v = out[0] + sF[S::kB3C];
                    const float err = v - cur.ret;
                    d[0] = cur.valid ? 2.0f * hp.vf_coef * err : 0.f;
                    g_b3c += d[0];
                    if (cur.valid) st_v += err * err;
                }
                *reinterpret_cast<uint4*>(smem + S::DOUT + op_offset(128, tid, 0)) =
                    make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
                *reinterpret_cast<uint4*>(smem + S::DOUT + op_offset(128, tid, 1)) = make_uint4(0u, 0u, 0u, 0u);
            }
            // ---- backward ------------------------------------------------------------------------------------------
            handoff();
            if (warp_u == 0) {                  // warp-uniform branch + elect.sync: no per-lane election loop around the MMAs
                fence_after();
                if (elect_one()) {
                    // dH2 = dOUT . W3^T   (B: forward W3 operand [16 x 128], MN-major: 8-hidden groups 256 B apart)
                    mma_bf16(tmem + cW, dk(S::DOUT, 128), dmn(W3, 256u), id_kmn128, 0u);
                    mma_commit(bar);
                    // dW3 += A2^T . dOUT  (not waited for here: the next commit covers it)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mma_bf16(tmem + cW3, dmn(S::A2 + j * 256, 2048u), dmn(S::DOUT + j * 256, 2048u), id_mm16, j > 0 ? 1u : first);
                }
                __syncwarp();
            }
            wait_phase();
            epilogue_mask(S::A2, S::D2);
            handoff();
            if (warp_u == 0) {                  // warp-uniform branch + elect.sync: no per-lane election loop around the MMAs
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)     // dH1 = D2 . W2^T
                        mma_bf16(tmem + cW, dk(S::D2 + j * 4096, 128), dmn(W2 + j * 256, 2048u), id_kmn128, j > 0);
                    mma_commit(bar);
#pragma unroll
                    for (int j = 0; j < 8; ++j)     // dW2 += A1^T . D2
                        mma_bf16(tmem + cW2, dmn(S::A1 + j * 256, 2048u), dmn(S::D2 + j * 256, 2048u), id_mm128, j > 0 ? 1u : first);
#pragma unroll
                    for (int j = 0; j < 8; ++j)     // db2 (column kBiasK) += D2^T . A0[K step kBiasStep]
                        mma_bf16(tmem + cB2, dmn(S::D2 + j * 256, 2048u), dmn(a0 + kBiasStep * 4096 + j * 256, 2048u), id_mm16, j > 0 ? 1u : first);
                }
                __syncwarp();
            }
            wait_phase();                       // also covers dW3: A2 may be overwritten now
            epilogue_mask(S::A1, S::A2);        // D1 -> the A2 buffer
            handoff();
            if (warp_u == 0) {                  // warp-uniform branch + elect.sync: no per-lane election loop around the MMAs
                fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)     // dW1^T (columns 0..D-1), db1 (column D) += D1^T . A0
                        mma_bf16(tmem + cW1, dmn(S::A2 + j * 256, 2048u), dmn(a0 + j * 256, 2048u), id_mmK1, j > 0 ? 1u : first);
                }
                __syncwarp();
            }
        }
        cur = nxt;
    }

    // ---- flush: TMEM accumulators -> this CTA's partial-gradient row --------------------------------------------------
    handoff();
    if (warp_u == 0) { fence_after(); if (elect_one()) mma_commit(bar); __syncwarp(); }
    wait_phase();
    const int P = L.total;
    float* out = partial + (size_t)blockIdx.x * partial_stride(P);
    const float scale = 1.0f / (float)b.n;
    for (int i = L.mean + tid; i < P; i += kM) out[i] = 0.f;                 // the observation normaliser is not trained
    const bool any = it > 0;                                                 // a CTA without tiles holds garbage in TMEM
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
        const uint32_t base = 128u + (uint32_t)net * kNet;
        const uint32_t cW2 = base, cW1 = base + 128u, cW3 = base + 128u + (uint32_t)K1, cB2 = cW3 + 16u;
        const int oW2 = net ? L.cW2 : L.aW2, oW1 = net ? L.cW1 : L.aW1, ob1 = net ? L.cb1 : L.ab1, ob2 = net ? L.cb2 : L.ab2;
        // dW2: lane = input feature k, column = output feature n  ->  W2[k][n], 512 contiguous bytes per thread
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(tmem + lane_off + cW2 + (uint32_t)(c * 32), v);
            float4* d4 = reinterpret_cast<float4*>(out + oW2 + tid * kH + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                d4[q] = any ? make_float4(v[4 * q] * scale, v[4 * q + 1] * scale, v[4 * q + 2] * scale, v[4 * q + 3] * scale)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {
            float v[K1];                                                      // lane = hidden n, column = obs k | D: bias
            tmem_ld16(tmem + lane_off + cW1, v);
            if constexpr (K1 == 32) tmem_ld16(tmem + lane_off + cW1 + 16u, v + 16);
#pragma unroll
            for (int k = 0; k < D; ++k) out[oW1 + k * kH + tid] = any ? v[k] * scale : 0.f;
            out[ob1 + tid] = any ? v[D] * scale : 0.f;
        }
        float v[16];
        tmem_ld16(tmem + lane_off + cB2, v);
        out[ob2 + tid] = any ? v[kBiasK] * scale : 0.f;
        tmem_ld16(tmem + lane_off + cW3, v);                                  // lane = hidden k, column = head output
        if (net == 0) {
#pragma unroll
            for (int j = 0; j < Ao; ++j) out[L.aW3 + tid * Ao + j] = any ? v[j] * scale : 0.f;
        } else {
            out[L.cW3 + tid] = any ? v[0] * scale : 0.f;
        }
    }
    // head biases, log_std and statistics: per-thread sums over the CTA's tiles -> block reduction
    {
        // 0..7 actor head bias | 8 critic head bias | 9..12 log_std | 13.. statistics: pg loss, v loss, clipped, KL, n, entropy
        float r[19];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = k < Ao ? g_b3a[k < Ao ? k : 0] : 0.f;
        r[8] = g_b3c;
#pragma unroll
        for (int k = 0; k < 4; ++k) r[9 + k] = g_ls[k];
        r[13] = st_pg; r[14] = st_v; r[15] = st_clip; r[16] = st_kl; r[17] = st_n; r[18] = st_ent;
#pragma unroll
        for (int k = 0; k < 19; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 19; ++k) sRed[warp * 24 + k] = r[k];
        }
        __syncthreads();
        if (tid < 19) {
            const float s = sRed[tid] + sRed[24 + tid] + sRed[48 + tid] + sRed[72 + tid];
            if (tid < 8) { if (tid < Ao) out[L.ab3 + tid] = s * scale; }
            else if (tid == 8) out[L.cb3] = s * scale;
            // DIST 0 entropy bonus: H = sum_k (0.5 + 0.5 log 2pi + log_std_k) does not depend on the sample; CTA 0 carries it
            else if (tid < 13) { if (DIST == 0) out[L.log_std + tid - 9] = s * scale - (blockIdx.x == 0 ? hp.ent_coef : 0.f); }
            else out[P + tid - 13] = s;                                       // statistics: plain sums (P + 0 .. P + 5)
        }
        if (tid >= 19 && tid < 13 + kPartialStats) out[P + tid - 13] = 0.f;
    }
    fence_before();
    __syncthreads();
    if (tid < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Running observation normaliser (brax.training.acme.running_statistics [third party, restated]; enabled by
// normalize_observations=True at train_brax_ppo.py:611): per training step the whole batch of observations is merged
// into (count, mean[D], summed_variance[D]) with the parallel Welford / Chan update, and the policy's
// (obs_mean, obs_inv_std) entries of the packed parameter vector are refreshed: std = clip(sqrt(max(M2 / count, 0)),
// std_min, std_max).  Two kernels: block partial sums in double, then one small CTA that merges and writes.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kObsStatBlocks = 296;

template <int D>
__global__ void __launch_bounds__(256)
obs_stats_partial_kernel(const float* __restrict__ obs, long long n, double* __restrict__ part /*[blocks][2 D]*/) {
    // thread t walks rows t, t + stride, ...: shifted sums (about the first row) keep the double accumulation exact enough
    __shared__ double sh[8][2 * D];
    double s[D], q[D];
    float ref[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { s[k] = 0.0; q[k] = 0.0; ref[k] = obs[k]; }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float* row = obs + i * D;
#pragma unroll
        for (int k = 0; k < D; ++k) { const double x = (double)(__ldg(row + k) - ref[k]); s[k] += x; q[k] += x * x; }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s[k] += __shfl_xor_sync(0xffffffffu, s[k], o); q[k] += __shfl_xor_sync(0xffffffffu, q[k], o); }
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { sh[w][k] = s[k]; sh[w][D + k] = q[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * D) {
        double t = 0.0;
        for (int ww = 0; ww < 8; ++ww) t += sh[ww][threadIdx.x];
        part[(size_t)blockIdx.x * 2 * D + threadIdx.x] = t;
    }
}

template <int D>
__global__ void __launch_bounds__(32)
obs_stats_merge_kernel(const float* __restrict__ obs, long long n, const double* __restrict__ part, int blocks,
                       double* __restrict__ running /*count | mean[D] | M2[D]*/, float* __restrict__ mean_out,
                       float* __restrict__ inv_std_out, float std_min, float std_max) {
    const int k = threadIdx.x;
    if (k >= D) return;
    double S = 0.0, Q = 0.0;
    for (int b = 0; b < blocks; ++b) { S += part[(size_t)b * 2 * D + k]; Q += part[(size_t)b * 2 * D + D + k]; }
    const double ref = (double)obs[k];
    const double nb = (double)n;
    const double mean_b = ref + S / nb;
    const double M2_b = fmax(Q - S * S / nb, 0.0);
    const double na = running[0], mean_a = running[1 + k], M2_a = running[1 + D + k];
    const double tot = na + nb;
    const double delta = mean_b - mean_a;
    const double mean = mean_a + delta * nb / tot;
    const double M2 = M2_a + M2_b + delta * delta * na * nb / tot;
    __syncwarp();
    running[1 + k] = mean; running[1 + D + k] = M2;
    if (k == 0) running[0] = tot;
    double sd = sqrt(fmax(M2 / tot, 0.0));
    sd = fmin(fmax(sd, (double)std_min), (double)std_max);
    mean_out[k] = (float)mean;
    inv_std_out[k] = (float)(1.0 / sd);
}

}  // namespace ppo
}  // namespace qs
