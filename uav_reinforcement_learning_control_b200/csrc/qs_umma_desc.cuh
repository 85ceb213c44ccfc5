// tcgen05 (UMMA) descriptor packing, in ONE place: the shared-memory matrix descriptor and the kind::f16 instruction
// descriptor every tensor-core kernel of this library builds (qs_rollout_tc.cuh, qs_ppo.cuh, qs_ppo_generic.cuh).
// Plain integer arithmetic, host + device, so tests/test_umma_desc.py can check the bit layouts on the CPU (through
// tests/host_harness) against an independent restatement of the published field tables
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QS_UMMA_HD __host__ __device__ __forceinline__
#else
#define QS_UMMA_HD inline
#endif

namespace qs {
namespace tc {

// SWIZZLE_NONE shared-memory matrix descriptor.  Core matrices are 8 rows x 16 bytes, stored as 128 contiguous bytes;
// lbo / sbo are the byte strides between core matrices along the two tile directions (which is "leading" depends on
// the operand's major-ness; the callers name them).  All three go in as bytes and are stored without their 4 LSBs.
QS_UMMA_HD uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                 // start address       [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;        // leading byte offset [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;        // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                                   // descriptor version 1 (Blackwell) [46,48)
    return d;                                                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// kind::f16 instruction descriptor: D = f32 (c_format 1), A = B = bf16 (format 1), dense, no negate, both K-major
QS_UMMA_HD uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// the same with either operand MN-major (a_major bit 15, b_major bit 16)
QS_UMMA_HD uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
    return make_idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

}  // namespace tc
}  // namespace qs
