// y = act(x W^T + b) on the 5th-generation tensor cores with fp32-level accuracy ("3xTF32"), for the projections that
// surround the sampling op (MSDeformAttn.value_proj / output_proj, the FFN; /root/reference/models/deformable_transformer.py:
// 95,113,219-224).  The reference runs them as strict-fp32 nn.Linear; on B200 cuBLAS serves that with SIMT sgemm kernels
// (~60 TFLOP/s).  Here every fp32 operand is used as hi + lo with hi = the 19 bits tcgen05.mma.kind::tf32 actually reads
// (the hardware ignores the low 13 mantissa bits) and lo = the exact remainder, and the product is accumulated in fp32 in
// tensor memory as   x_lo W_hi + x_hi W_lo + x_hi W_hi   (the lo*lo term, 2^-22 relative, is dropped).
//
//   * operands: x (M, K) and W (N, K), both K-major, row-major fp32, K % 32 == 0, N % 128 == 0; W_lo is prepared once per
//     weight (cape_tf32_split_lo); x_lo is produced inside the kernel from the x tile the TMA already brought in and is
//     handed to the tensor core through TENSOR MEMORY (tcgen05.st; the MMA takes its A operand from there), so a
//     pipeline stage holds three 16 KB tiles (x, W_hi, W_lo) and four stages fit;
//   * persistent: one CTA per SM walks 128 x 128 output tiles; K step 32 (one 128-byte swizzle atom per row); two
//     accumulators in tensor memory so a tile's epilogue overlaps the next tile's main loop (512 columns in all:
//     2 x 128 accumulator columns + 4 x 32 columns of x_lo);
//   * 10 warps: warp 0 = TMA producer (one elected lane), warp 1 = tensor-memory allocator + MMA issuer (one elected
//     lane), warps 2..5 = x_lo converters (thread = row: 8 conflict-free LDS.128 of the swizzled row, 32 cvt.rna.tf32,
//     one tcgen05.st.32x32b.x32), warps 6..9 = epilogue (tcgen05.ld of their 32-lane quarter, bias, activation, swizzled
//     staging in shared memory, TMA tensor store — per-thread row stores would cost 32 LSU wavefronts each);
//   * mbarriers: full (TMA bytes landed), conv (x_lo in tensor memory), empty (tcgen05.commit: the MMAs that read the
//     stage are done), tmem_full / tmem_empty (accumulator complete / drained);
//   * split-K mode for weight gradients (tiny output, reduction over ~10^5 rows): work item = (tile, K slice of <= 1024),
//     partial tiles are ADDED into y by the TMA (cp.reduce.async.bulk.tensor ... add);
//   * MNMAJOR variant for those weight gradients, grad_w = g^T x: g (rows, N) and x (rows, K) are read where they are — the
//     reduction index is the slow one in memory, so both operands are MN-major tiles (128-byte swizzle with 32-byte atoms,
//     four TMA boxes of 32 rows x 32 columns per operand and stage) and both lo tiles are split in shared memory by the
//     converter warps; no transposed copies (they were 6.7 % of a training step): 96 us instead of 232 us for
//     108 800 x 256 x 256, 312 instead of 680 us for 108 800 x 1024 x 256 (cuBLAS fp32: 390 / 1130 us).  Work items run
//     output-tile-fastest so that concurrent CTAs share a row slice's operand tiles in L2.
// What bounds it (ncu): tensor pipe ~58 % active; per tile the SM takes in 384 KB of operand tiles (2/3 of it the W tiles
// every CTA re-reads from L2), ~31 B/clk/SM — the L2 -> shared-memory ingress of one SM, not the stage count (3 -> 4
// stages gained 3 %).  A 256-row tile per CTA (two MMAs per W tile) is the next step.
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32, kStages = 4;
constexpr int kTileBytes = kBM * kBK * 4;                 // 16 KB: one operand tile (128 rows x 128 B)
constexpr int kStageBytes = 3 * kTileBytes;               // x, W_hi, W_lo (x_lo lives in tensor memory)
constexpr int kTmemCols = 512;                            // 2 accumulators x 128 + kStages x 32 columns of x_lo
constexpr int kALoCol = 2 * kBN;
constexpr int kGemmThreads = 320;          // TMA warp, MMA warp, 4 converter warps, 4 epilogue warps
constexpr int kConvThreads = 128;
constexpr int kEpiThreads = 128;
constexpr int kStoreTile = 32 * 32 * 4;                   // epilogue staging: 32 rows x 32 columns per warp and step
constexpr size_t kGemmSmem = static_cast<size_t>(kStages) * kStageBytes + 1024 /* alignment slack */ + 1024 /* barriers */ +
                             4 * 2 * kStoreTile /* 4 epilogue warps x 2 buffers */;

// lo part of an fp32 number for the hi/lo split: the exact remainder x - trunc19(x) (<= 13 significant bits), rounded
// to nearest tf32 — the tensor core would otherwise truncate it, and a truncation bias adds up coherently over K.
__device__ __forceinline__ float tf32_lo(float v) {
    const float rem = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(rem));
    return __uint_as_float(r);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a broken pipeline traps (a CUDA error the caller sees) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 24)) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
        : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    const uint32_t lo = (smem_addr >> 4) & 0x3fffu;                       // start address; leading byte offset unused (0)
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);           // stride byte offset, version 1, SWIZZLE_128B
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// MN-major operand tile (the contraction index is the SLOW one in memory: weight gradients read grad_out / x as they are).
// For 32-bit MN-major operands the tensor core accepts one shared-memory layout only, the 128-byte swizzle with 32-byte
// atoms (cute::UMMA::Layout_MN_SW128_32B_Atom, layout type SWIZZLE_128B_BASE32B; the TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B = 32 consecutive M (or N) indices of one k, the 32-byte chunks of a
// row XOR-ed with k % 4; 4 k-rows per 512-byte atom, the next 4 k's 512 B further (stride byte offset), the next 32 M
// indices 4096 B further (leading byte offset) — ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3fffu) | ((4096u >> 4) << 16);
    const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = 128.
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kBN >> 3) << 17) | ((kBM >> 4) << 24);
constexpr uint32_t kInstrDescMN = kInstrDesc | (1u << 15) | (1u << 16);     // a_major = b_major = MN

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate,
                                          uint32_t idesc = kInstrDesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same MMA with the A operand read from tensor memory (128 lanes = rows, 8 columns = one K step of tf32 values)
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// MNMAJOR (weight gradients, y (M, N) = a^T b with a (K, M) and b (K, N) row-major, i.e. both operands MN-major): map_x = a,
// map_wh = b, map_wl unused; a stage is four tiles [a | a_lo | b | b_lo] (three stages), both lo tiles are produced in
// shared memory by the converter warps (an elementwise pass over the landed tiles: the layout is kept), no tensor-memory
// operand.  K is then the number of ROWS of a / b and may end inside a block (the TMA zero-fills).
template <int ACT, bool MNMAJOR = false>   // ACT: 0 none, 1 ReLU
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                     const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_y,
                     const float* __restrict__ bias, int M, int N, int K, int splits) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kStages = MNMAJOR ? 3 : cape::kStages;                  // shadows the file-level constants inside this kernel
    constexpr int kStageBytes = (MNMAJOR ? 4 : 3) * kTileBytes;           // (3 x 64 KB = 4 x 48 KB)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    // barriers: full[S], conv[S], empty[S], tmem_full[2], tmem_empty[2]; then the tensor-memory base address
    const uint32_t bars = base + kStages * kStageBytes;
    const auto full = [&](int s) { return bars + 8u * s; };
    const auto conv = [&](int s) { return bars + 8u * (kStages + s); };
    const auto empty = [&](int s) { return bars + 8u * (2 * kStages + s); };
    const auto tmem_full = [&](int b) { return bars + 8u * (3 * kStages + b); };
    const auto tmem_empty = [&](int b) { return bars + 8u * (3 * kStages + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + kStages * kStageBytes + 8 * (3 * kStages + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Work item t = (output tile, K split): with splits > 1 (weight gradients: tiny output, reduction over all rows)
    // every item covers a slice of the reduction dimension and the epilogue ADDS its partial tile to y (TMA reduce).
    const int k_blocks_all = MNMAJOR ? (K + kBK - 1) / kBK : K / kBK;
    const int kb_per = (k_blocks_all + splits - 1) / splits;
    const int n_tiles = N / kBN;
    const int tiles = n_tiles * ((M + kBM - 1) / kBM) * splits;            // persistent: item t = blockIdx.x, += gridDim.x
    const int out_tiles = n_tiles * ((M + kBM - 1) / kBM);
    const auto item = [&](int t, int& m0, int& n0, int& kb0, int& kb1) {
        // weight gradients: the output tile is the FAST index, so the CTAs running at the same time work on the same slice of
        // rows and share its operand tiles through L2 (every operand tile is needed by all tiles of the other dimension:
        // with the slice as the fast index DRAM read 1.63 GB for 108 800 x 1024 x 256, 2.9x the operands)
        const int sp = MNMAJOR ? t / out_tiles : t % splits, tile = MNMAJOR ? t % out_tiles : t / splits;
        m0 = (tile / n_tiles) * kBM;
        n0 = (tile % n_tiles) * kBN;
        kb0 = sp * kb_per;
        kb1 = min(k_blocks_all, kb0 + kb_per);
    };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(conv(s), kConvThreads);
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full(b), 1);
            mbar_init(tmem_empty(b), kEpiThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // 2 x 128 columns of tensor memory: two 128 x 128 fp32 accumulators (mainloop / epilogue overlap)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(tmem_slot))), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int m0, n0, kb0, kb1;
                item(t, m0, n0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);
                    const uint32_t st = base + s * kStageBytes;
                    if (MNMAJOR) {   // four boxes of 32 k-rows x 32 columns per operand: one per 32-wide M / N block
                        mbar_arrive_expect_tx(full(s), 2 * kTileBytes);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            tma_load_2d(st + j * 4096, &map_x, full(s), m0 + 32 * j, kb * kBK);
                            tma_load_2d(st + 2 * kTileBytes + j * 4096, &map_wh, full(s), n0 + 32 * j, kb * kBK);
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(full(s), 3 * kTileBytes);
                    tma_load_2d(st, &map_x, full(s), kb * kBK, m0);
                    tma_load_2d(st + kTileBytes, &map_wh, full(s), kb * kBK, n0);
                    tma_load_2d(st + 2 * kTileBytes, &map_wl, full(s), kb * kBK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer =====
            int it = 0, j = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++j) {
                const int ab = j & 1;
                mbar_wait(tmem_empty(ab), ((j >> 1) & 1) ^ 1);            // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + ab * kBN;
                int m0, n0, kb0, kb1;
                item(t, m0, n0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t parity = (it / kStages) & 1;
                    mbar_wait(full(s), parity);
                    mbar_wait(conv(s), parity);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * kStageBytes;
                    if (MNMAJOR) {
                        const uint64_t d_a = umma_desc_mn(st), d_alo = umma_desc_mn(st + kTileBytes);
                        const uint64_t d_b = umma_desc_mn(st + 2 * kTileBytes), d_blo = umma_desc_mn(st + 3 * kTileBytes);
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k) {   // 8 k-rows (two 512-byte atoms, 1024 B) per UMMA_K step
                            const uint64_t adv = static_cast<uint64_t>(k * (1024 >> 4));
                            umma_tf32(tmem_acc, d_alo + adv, d_b + adv, (kb != kb0) | (k != 0), kInstrDescMN);
                            umma_tf32(tmem_acc, d_a + adv, d_blo + adv, 1, kInstrDescMN);
                            umma_tf32(tmem_acc, d_a + adv, d_b + adv, 1, kInstrDescMN);
                        }
                        umma_commit(empty(s));
                        continue;
                    }
                    const uint64_t d_x = umma_desc(st);
                    const uint64_t d_wh = umma_desc(st + kTileBytes), d_wl = umma_desc(st + 2 * kTileBytes);
                    const uint32_t t_xlo = tmem_base + kALoCol + s * kBK;
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {   // UMMA_K = 8 for tf32: 32 bytes along K -> +2 in the address field
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        umma_tf32_ta(tmem_acc, t_xlo + k * 8, d_wh + adv, (kb != kb0) | (k != 0));
                        umma_tf32(tmem_acc, d_x + adv, d_wl + adv, 1);
                        umma_tf32(tmem_acc, d_x + adv, d_wh + adv, 1);
                    }
                    umma_commit(empty(s));                  // the stage may be refilled once these MMAs have read it
                }
                umma_commit(tmem_full(ab));
            }
        }
    } else if (warp < 2 + kConvThreads / 32) {
        // ===== x_lo converters: thread = row of the x tile; the lo parts go to tensor memory (lane = row, column = k) =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            int m0, n0, kb0, kb1;
            item(t, m0, n0, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % kStages;
                mbar_wait(full(s), (it / kStages) & 1);
                if (MNMAJOR) {   // lo tiles of both operands, written next to them in the same (swizzled) layout
                    uint8_t* st_ptr = base_ptr + s * kStageBytes;
                    const int t128 = (warp - 2) * 32 + lane;
#pragma unroll
                    for (int tile = 0; tile < 2; ++tile) {
                        const uint8_t* hi = st_ptr + tile * 2 * kTileBytes;
#pragma unroll
                        for (int i = 0; i < kTileBytes / (16 * kConvThreads); ++i) {
                            const int o = (i * kConvThreads + t128) * 16;
                            const float4 v = *reinterpret_cast<const float4*>(hi + o);
                            *reinterpret_cast<float4*>(const_cast<uint8_t*>(hi) + kTileBytes + o) =
                                make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(conv(s));
                    continue;
                }
                const uint8_t* src = base_ptr + s * kStageBytes + row * 128;
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {   // 16-byte chunk j of the row sits at chunk position j ^ (row % 8) (128-byte swizzle)
                    const float4 v = *reinterpret_cast<const float4*>(src + ((j ^ (row & 7)) << 4));
                    r[4 * j] = __float_as_uint(tf32_lo(v.x));
                    r[4 * j + 1] = __float_as_uint(tf32_lo(v.y));
                    r[4 * j + 2] = __float_as_uint(tf32_lo(v.z));
                    r[4 * j + 3] = __float_as_uint(tf32_lo(v.w));
                }
                const uint32_t taddr = tmem_base + kALoCol + s * kBK + (static_cast<uint32_t>(quarter * 32) << 16);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                    ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
                      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
                      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                    : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(conv(s));
            }
        }
    } else {
        // ===== epilogue: warp w owns tensor-memory lanes 32 (w % 4) .. +31 =====
        const int quarter = warp & 3;
        int j = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++j) {
            int m0, n0, kb0, kb1;
            item(t, m0, n0, kb0, kb1);
            const int ab = j & 1;
            mbar_wait(tmem_full(ab), (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // staging buffers of this warp (1024-aligned, 128-byte swizzle like the TMA store expects)
            uint8_t* stage_ptr = base_ptr + kStages * kStageBytes + 1024 + (warp - 6) * 2 * kStoreTile;
#pragma unroll 1
            for (int c = 0; c < kBN / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ab * kBN + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // the buffer used two steps ago must have been read by its TMA store
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                uint8_t* buf = stage_ptr + (c & 1) * kStoreTile;
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                    float4 o;
                    o.x = __uint_as_float(r[jj]), o.y = __uint_as_float(r[jj + 1]);
                    o.z = __uint_as_float(r[jj + 2]), o.w = __uint_as_float(r[jj + 3]);
                    if (bias) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + jj));
                        o.x += bv.x, o.y += bv.y, o.z += bv.z, o.w += bv.w;
                    }
                    if (ACT == 1) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
                    // row = lane (128 B), 16-byte chunk jj / 4, swizzled with the row's low 3 bits: conflict-free STS.128
                    *reinterpret_cast<float4*>(buf + lane * 128 + (((jj >> 2) ^ (lane & 7)) << 4)) = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {   // rows past M are clipped by the tensor map
                    if (splits > 1)
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&map_y)), "r"(smem_u32(buf)), "r"(n0 + c * 32),
                                       "r"(m0 + quarter * 32)
                                     : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&map_y)), "r"(smem_u32(buf)), "r"(n0 + c * 32),
                                       "r"(m0 + quarter * 32)
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tmem_empty(ab));                   // this accumulator may be overwritten by the next-but-one tile
        }
    }
    if (warp >= 6 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

// lo part of an fp32 tensor for the kernel above: x - (x with the 13 low mantissa bits cleared), exact in fp32.
__global__ void tf32_split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        lo[i] = tf32_lo(x[i]);
    }
}

// out[c][r] = in[r][c] for a row-major (R, C) matrix, optionally also the tf32 lo part of the transposed matrix: the
// operands of the weight-gradient GEMM g^T . x are the activations with the ROW index as the contiguous (reduction) one.
__global__ void __launch_bounds__(256)
transpose_lo_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out_lo, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        tile[ty + i][tx] = (r < R && c < C) ? in[static_cast<int64_t>(r) * C + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (c < C && r < R) {
            const float v = tile[tx][ty + i];
            out[static_cast<int64_t>(c) * R + r] = v;
            if (out_lo) out_lo[static_cast<int64_t>(c) * R + r] = tf32_lo(v);
        }
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// (rows, K) row-major fp32 -> boxes of box_rows rows x 32 columns, 128-byte swizzle; out-of-range rows read as zeros /
// are not written.
bool make_map(CUtensorMap* map, const float* ptr, int rows, int K, int box_rows = kBM,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * sizeof(float)};
    const cuuint32_t box[2] = {kBK, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t elem[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, elem,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

cudaError_t launch_tf32_split_lo(const float* x, float* lo, int64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int64_t grid = (n + 255) / 256;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    tf32_split_lo_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(x, lo, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_transpose_lo(const float* in, float* out, float* out_lo, int R, int C, cudaStream_t stream) {
    if (R == 0 || C == 0) return cudaSuccess;
    const dim3 grid((C + 31) / 32, (R + 31) / 32);
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
    transpose_lo_kernel<<<grid, 256, 0, stream>>>(in, out, out_lo, R, C);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_linear_tf32x3(const float* x, const float* w, const float* w_lo, const float* bias, float* y, int M, int N,
                                 int K, int act, int split_k, cudaStream_t stream) {
    if (M == 0) return cudaSuccess;
    CUtensorMap map_x, map_wh, map_wl, map_y;
    if (!make_map(&map_x, x, M, K) || !make_map(&map_wh, w, N, K) || !make_map(&map_wl, w_lo, N, K) ||
        !make_map(&map_y, y, M, N, 32))
        return cudaErrorNotSupported;
    static unsigned long long configured = 0;
    if (first_use_on_device(&configured)) {
        cudaFuncSetAttribute(linear_tf32x3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
        cudaFuncSetAttribute(linear_tf32x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int out_tiles = (N / kBN) * ((M + kBM - 1) / kBM);
    // split the reduction dimension when the output alone cannot fill the machine (split_k: 0 = never, 1 = automatic);
    // partial tiles are added into y by the TMA, so y is zero-filled first and there is no bias / activation
    // A slice is at most kMaxChain K-blocks long: the tensor core's fp32 accumulation is not round-to-nearest, and over a
    // reduction of 10^5 rows (weight gradients) one accumulator chain drifts to ~8e-5; 1024-element chains whose partial
    // tiles are added in L2 (round-to-nearest) stay at the forward's ~5e-6.
    constexpr int kMaxChain = 32;
    int splits = 1;
    if (split_k) {
        const int k_blocks = K / kBK;
        splits = out_tiles < sms ? sms / out_tiles : 1;
        if (splits * kMaxChain < k_blocks) splits = (k_blocks + kMaxChain - 1) / kMaxChain;
        if (splits > k_blocks) splits = k_blocks;
        if (splits < 1) splits = 1;
        const int per = (k_blocks + splits - 1) / splits;
        splits = (k_blocks + per - 1) / per;             // no empty slice
    }
    if (splits > 1) {
        if (bias || act) return cudaErrorInvalidValue;
        const cudaError_t e = cudaMemsetAsync(y, 0, static_cast<size_t>(M) * N * sizeof(float), stream);
        if (e != cudaSuccess) return e;
    }
    const int tiles = out_tiles * splits;
    const dim3 grid(tiles < sms ? tiles : sms);          // persistent: one CTA per SM walks the work items
    if (act == 1)
        linear_tf32x3_kernel<1><<<grid, kGemmThreads, kGemmSmem, stream>>>(map_x, map_wh, map_wl, map_y, bias, M, N, K, splits);
    else
        linear_tf32x3_kernel<0><<<grid, kGemmThreads, kGemmSmem, stream>>>(map_x, map_wh, map_wl, map_y, bias, M, N, K, splits);
    count_launch();
    return cudaGetLastError();
}

// grad_w (N, K) = grad_out^T . x for grad_out (rows, N) and x (rows, K), both row-major: the operands are read where they
// are (MN-major tiles), no transposed copies; rows are split into chains of <= 1024 whose partial tiles the TMA adds into
// grad_w (zero-filled first).  N % 32 == 0, K % 128 == 0.
cudaError_t launch_wgrad_tf32x3(const float* grad_out, const float* x, float* grad_w, int rows, int N, int K, cudaStream_t stream) {
    if (rows == 0) return cudaMemsetAsync(grad_w, 0, static_cast<size_t>(N) * K * sizeof(float), stream);
    CUtensorMap map_g, map_x, map_y;
    // boxes of 32 rows (reduction index) x 32 columns of the row-major operands
    if (!make_map(&map_g, grad_out, rows, N, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map(&map_x, x, rows, K, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) || !make_map(&map_y, grad_w, N, K, 32))
        return cudaErrorNotSupported;
    static unsigned long long configured = 0;
    if (first_use_on_device(&configured))
        cudaFuncSetAttribute(linear_tf32x3_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int out_tiles = (K / kBN) * ((N + kBM - 1) / kBM);
    constexpr int kMaxChain = 32;                      // see launch_linear_tf32x3
    const int k_blocks = (rows + kBK - 1) / kBK;
    int splits = out_tiles < sms ? sms / out_tiles : 1;
    if (splits * kMaxChain < k_blocks) splits = (k_blocks + kMaxChain - 1) / kMaxChain;
    if (splits > k_blocks) splits = k_blocks;
    if (splits < 1) splits = 1;
    const int per = (k_blocks + splits - 1) / splits;
    splits = (k_blocks + per - 1) / per;
    if (splits > 1) {
        const cudaError_t e = cudaMemsetAsync(grad_w, 0, static_cast<size_t>(N) * K * sizeof(float), stream);
        if (e != cudaSuccess) return e;
    }
    const int tiles = out_tiles * splits;
    const dim3 grid(tiles < sms ? tiles : sms);
    linear_tf32x3_kernel<0, true><<<grid, kGemmThreads, kGemmSmem, stream>>>(map_g, map_x, map_x, map_y, nullptr, N, K, rows, splits);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cape
