// tcgen05 helpers shared by the tensor-core grad_value scatters (msda_backward_staged.cu, msda_backward_tc.cu): 3xTF32 split,
// K-major 128-byte-swizzled operand tiles, kind::tf32 MMA issue and commit (sm_100a).
#pragma once

#include <stdint.h>

namespace cape {

__device__ __forceinline__ float tf32_lo_part(float v) {
    const float rem = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(rem));
    return __uint_as_float(r);
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (same encoding as linear_tf32x3.cu).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    const uint32_t lo = (smem_addr >> 4) & 0x3fffu;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128; N = 64 ([G_hi | G_lo]) and N = 32 (G_hi).
constexpr uint32_t kIdescM64N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
constexpr uint32_t kIdescM64N32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24);
constexpr uint32_t kIdescM128N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescM128N32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Float index of Wt[row][query column] inside a weight tile (128-byte rows, 16-byte chunks XOR-swizzled by row & 7).
__device__ __forceinline__ uint32_t tile_index(int row, int col) {
    return static_cast<uint32_t>(row) * 32u + ((((static_cast<uint32_t>(col) >> 2) ^ (static_cast<uint32_t>(row) & 7u)) << 2) |
                                               (static_cast<uint32_t>(col) & 3u));
}

}  // namespace cape
