// MSDeformAttn backward with the two coarsest pyramid levels kept on the SM: their value rows are staged in shared
// memory by the TMA (gathers become LDS.128) and their grad_value scatter runs on the tensor cores instead of the L2
// reduction units.
//
// Derivative of ms_deform_attn_core_pytorch (/root/reference/models/deformable_transformer.py:129-141); same per-sample
// arithmetic as msda_bwd_fast_kernel (msda_backward.cu).  What bounds that kernel is the 44 M `red.global.add.v4.f32` rows
// per launch at the L2's ~51 G rows/s, with the LSU wavefront pipe right behind.  Here, per CTA
// (persistent, one per SM, a contiguous range of the (image, head, query) space):
//
//   warps 0..30  sample: one query per warp at a time, lane = (point, channel quad).  Levels staged in shared memory are
//                gathered with LDS.128; the other levels with LDG.128 as before.  grad_loc / grad_attn for all levels;
//                REDs into grad_value only for the levels the tensor cores do NOT cover.
//   warp 31      builds, for batches of 32 queries (thread = query), the bilinear weight matrix of the covered levels
//                Wt[pixel][query] = sum over the query's samples of A * w_corner  (<= 32 non-zeros per column; plain
//                LDS / FADD / STS read-modify-writes are race-free because a column has one owner), its tf32 "lo" part,
//                and the transposed grad_out tile G[channel][query] (hi + lo) from a TMA-loaded copy; then lane 0 issues
//                grad_value[pixel][channel] += Wt . G^T as tcgen05.mma.kind::tf32 (3xTF32: lo*hi + hi*lo + hi*hi, M = 128
//                pixels, N = 32 channels, K = 8 queries per instruction) into accumulators in TENSOR MEMORY that live for
//                the whole (image, head) segment.  After the MMAs of a batch complete (tcgen05.commit -> mbarrier) the
//                touched entries are zeroed again, so the tile is never cleared wholesale.
//   segment end  warps 0..3 read the accumulators (tcgen05.ld) and add them to grad_value with one RED per row and CTA.
//
// Covered levels = the last one or two levels whose pixels fit 384 rows (the CAPE pyramid: 16x16 + 8x8 = 320 pixels,
// 50 % of all samples, ~21 of the ~51 in-bounds corner rows of a (query, head)).
#include "async_copy.cuh"
#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kBsThreads = 1024;
constexpr int kSimtWarps = 31;
constexpr int kTcRows = 384;                      // pixels covered by the tensor-core scatter: 3 tiles of 128
constexpr int kATile = kTcRows * 128;             // bytes of Wt[pixel][32 queries] fp32 (rows of 128 B, 128-byte swizzle)
constexpr int kBTile = 32 * 128;                  // G[channel][32 queries]
constexpr int kOffALo = kATile, kOffBHi = 2 * kATile, kOffBLo = kOffBHi + kBTile, kOffG = kOffBLo + kBTile;
constexpr int kOffStage = kOffG + kBTile;         // staged value rows start here (1024-byte aligned)
constexpr int kBoxRowsB = 64;
constexpr int kTmemColsB = 128;                   // 3 accumulators x 32 columns, rounded up to a power of two
constexpr int kMaxDynSmemB = 227 * 1024 - 2048;

// Cycle counters of the builder warp's phases (CTA 0 only; read back by cape_debug_counters for profiling runs).
__device__ long long g_bs_cycles[16];
#define CAPE_TICK(slot)                                   \
    do {                                                  \
        if (kProfile) {                                   \
            const long long now_ = clock64();             \
            cyc[slot] += now_ - t_prev;                   \
            t_prev = now_;                                \
        }                                                 \
    } while (0)

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
// kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = 32.
constexpr uint32_t kIdescM128N32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kIdescM128N32), "r"(accumulate)
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

template <typename VT>
__device__ __forceinline__ float4 lds_row4_or_zero(uint32_t addr, bool pred);
template <>
__device__ __forceinline__ float4 lds_row4_or_zero<float>(uint32_t addr, bool pred) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "r"(addr), "r"(static_cast<int>(pred)));
    return v;
}
__device__ __forceinline__ uint2 lds_row2u_or_zero(uint32_t addr, bool pred) {
    uint2 r = make_uint2(0u, 0u);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.shared.v2.b32 {%0, %1}, [%2];\n\t}"
        : "+r"(r.x), "+r"(r.y)
        : "r"(addr), "r"(static_cast<int>(pred)));
    return r;
}
template <>
__device__ __forceinline__ float4 lds_row4_or_zero<__nv_bfloat16>(uint32_t addr, bool pred) {
    const uint2 r = lds_row2u_or_zero(addr, pred);
    float4 f;
    f.x = __uint_as_float(r.x << 16);
    f.y = __uint_as_float(r.x & 0xffff0000u);
    f.z = __uint_as_float(r.y << 16);
    f.w = __uint_as_float(r.y & 0xffff0000u);
    return f;
}
template <>
__device__ __forceinline__ float4 lds_row4_or_zero<__half>(uint32_t addr, bool pred) {
    const uint2 r = lds_row2u_or_zero(addr, pred);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// 12 per-lane partials summed over the 8 lanes of a point group with 12 shuffles (see msda_backward.cu).
__device__ __forceinline__ void transpose_reduce12s(const float (&v)[12], int k, float (&out)[3]) {
    const bool hi4 = k & 4, hi2 = k & 2;
    float h[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float keep = hi4 ? v[i + 6] : v[i];
        const float send = hi4 ? v[i] : v[i + 6];
        h[i] = keep + __shfl_xor_sync(kFullMask, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float keep = hi2 ? h[i + 3] : h[i];
        const float send = hi2 ? h[i] : h[i + 3];
        out[i] = keep + __shfl_xor_sync(kFullMask, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] += __shfl_xor_sync(kFullMask, out[i], 1);
}

// Float index of Wt[row][query column] inside a weight tile (128-byte rows, 16-byte chunks XOR-swizzled by row & 7).
__device__ __forceinline__ uint32_t tile_index(int row, int col) {
    return static_cast<uint32_t>(row) * 32u + ((((static_cast<uint32_t>(col) >> 2) ^ (static_cast<uint32_t>(row) & 7u)) << 2) |
                                               (static_cast<uint32_t>(col) & 3u));
}

// L = 4, P = 4, D = 32.  FUSED: locp / attnp are raw offsets / logits, refp the reference points (see msda_backward.cu).
template <typename VT, typename AT, bool FUSED>
__global__ void __launch_bounds__(kBsThreads, 1)
msda_bwd_staged_kernel(const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap gmap,
                       const VT* __restrict__ gout, const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ starts, const AT* __restrict__ locp, const AT* __restrict__ attnp,
                       const float* __restrict__ refp, float* __restrict__ gvalue, AT* __restrict__ gloc,
                       AT* __restrict__ gattn, int N, int S, int M, int Lq, int cap_rows, long long per_cta, int use_tc, int profile) {
    constexpr int L = 4, D = 32;
    constexpr int kRowB = D * static_cast<int>(sizeof(VT));
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_storage[3];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_addr_u32(smem_raw) + 1023u) & ~1023u;       // swizzle atoms need 1024-byte alignment
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
    const int rowStride = M * D;
    const uint32_t bar_stage = smem_addr_u32(&bar_storage[0]), bar_g = smem_addr_u32(&bar_storage[1]),
                   bar_mma = smem_addr_u32(&bar_storage[2]);

    int H[L], W[L], st[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        H[l] = static_cast<int>(__ldg(shapes + 2 * l));
        W[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const long long s0 = __ldg(starts + l);
        st[l] = static_cast<int>(s0);
        if (s0 < 0 || H[l] < 0 || W[l] < 0 || s0 + static_cast<long long>(H[l]) * W[l] > S) H[l] = W[l] = st[l] = 0;
    }
    // staged value rows [base_row, S): the longest suffix of levels that fits cap_rows
    int base_row = S;
    bool suffix = true;
#pragma unroll
    for (int l = L - 1; l >= 0; --l) {
        suffix = suffix && H[l] > 0 && st[l] < base_row && S - st[l] <= cap_rows;
        if (suffix) base_row = st[l];
    }
    const int rows = S - base_row;
    const int nboxes = (rows + kBoxRowsB - 1) / kBoxRowsB;
    // tensor-core scatter: the last level, or the last two, when their pixel rows fit kTcRows
    bool tc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) tc[l] = false;
    int tc_base = 0, tc_rows = 0;
    if (use_tc && H[L - 1] > 0 && H[L - 1] * W[L - 1] <= kTcRows) {
        tc[L - 1] = true;
        tc_base = st[L - 1];
        tc_rows = H[L - 1] * W[L - 1];
        if (H[L - 2] > 0 && st[L - 2] <= st[L - 1] && st[L - 1] + tc_rows - st[L - 2] <= kTcRows &&
            H[L - 2] * W[L - 2] <= st[L - 1] - st[L - 2]) {
            tc[L - 2] = true;
            tc_base = st[L - 2];
            tc_rows = st[L - 1] + tc_rows - st[L - 2];
        }
    }
    const bool any_tc = tc_rows > 0;
    const int n_mt = (tc_rows + 127) >> 7;
    bool in_smem[L];
    uint32_t lvl_off[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        in_smem[l] = rows > 0 && H[l] > 0 && st[l] >= base_row;
        lvl_off[l] = in_smem[l] ? static_cast<uint32_t>(st[l] - base_row) * kRowB : static_cast<uint32_t>(st[l]) * rowStride;
    }

    // ---- one-time set-up: barriers, tensor memory, zeroed weight tiles ------------------------------------------------
    if (tid == 0) {
        mbarrier_init(bar_stage, 1);
        mbarrier_init(bar_g, 1);
        mbarrier_init(bar_mma, 1);
        mbarrier_init_fence();
    }
    if (any_tc) {
        for (int i = tid; i < 2 * kATile / 16; i += kBsThreads)
            asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "f"(0.f) : "memory");
        fence_proxy_async_shared();
        if (warp == kSimtWarps) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(&tmem_slot)),
                         "n"(kTmemColsB)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = any_tc ? *reinterpret_cast<volatile uint32_t*>(&tmem_slot) : 0u;

    const long long total = static_cast<long long>(N) * M * Lq;
    long long pos = static_cast<long long>(blockIdx.x) * per_cta;
    const long long end = min(total, pos + per_cta);
    uint32_t stage_phase = 0, g_phase = 0, mma_phase = 0;
    uint32_t prev[16];                       // builder: packed tile indices (2 x 16 bit) of the entries written last batch
#pragma unroll
    for (int i = 0; i < 16; ++i) prev[i] = 0xffffffffu;

    while (pos < end) {
        const int nm = static_cast<int>(pos / Lq);
        const int q_begin = static_cast<int>(pos - static_cast<long long>(nm) * Lq);
        const int q_end = static_cast<int>(min(static_cast<long long>(Lq), q_begin + (end - pos)));
        const int n = nm / M, m = nm - n * M;
        __syncthreads();                     // previous segment: rows read, accumulators flushed
        if (rows > 0) {
            if (warp == 0) {
                if (lane == 0) mbarrier_arrive_expect_tx(bar_stage, static_cast<uint32_t>(nboxes) * kBoxRowsB * kRowB);
                __syncwarp();
                for (int b = lane; b < nboxes; b += 32)
                    tma_load_box_2d(base + kOffStage + b * kBoxRowsB * kRowB, &vmap, bar_stage, m * D,
                                    n * S + base_row + b * kBoxRowsB);
            }
            if (warp != kSimtWarps || !any_tc) mbarrier_wait(bar_stage, stage_phase);   // the builder never reads value rows
            stage_phase ^= 1;
        }

        if (warp < kSimtWarps || !any_tc) {
            // ===== sampling warps ==========================================================================================
            const int p = lane >> 3, k = lane & 7;
            const int nsimt = any_tc ? kSimtWarps : kBsThreads / 32;
            const int64_t headOff = (static_cast<int64_t>(n) * S * M + m) * D + k * 4;
            const VT* vbase = value + headOff;
            float* gbase = gvalue + headOff;
            const uint32_t sbase = base + kOffStage + k * (kRowB / 8);
            float dimf = 1.f;
#pragma unroll
            for (int l = 0; l < L; ++l)
                if ((lane >> 3) == l) dimf = static_cast<float>((lane & 1) ? H[l] : W[l]);
            const long long t_simt = (profile && blockIdx.x == 0 && tid == 0) ? clock64() : 0;
            for (int q = q_begin + warp; q < q_end; q += nsimt) {
                const int64_t qm = (static_cast<int64_t>(n) * Lq + q) * M + m;
                float locv = to_f32(locp[qm * (L * 8) + lane]);
                float attnv = FUSED ? -INFINITY : 0.f;
                if (lane < L * 4) attnv = to_f32(attnp[qm * (L * 4) + lane]);
                const float4 g = ld4(gout + qm * D + k * 4);
                if (FUSED) {
                    float mx = attnv;
#pragma unroll
                    for (int s = 8; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, s));
                    const float e = (lane < L * 4) ? expf(attnv - mx) : 0.f;
                    float sum = e;
#pragma unroll
                    for (int s = 8; s >= 1; s >>= 1) sum += __shfl_xor_sync(kFullMask, sum, s);
                    attnv = e / sum;
                    locv = __ldg(refp + (static_cast<int64_t>(n) * Lq + q) * (L * 2) + (lane >> 3) * 2 + (lane & 1)) + locv / dimf;
                }
                locv = pixel_coord(locv, dimf);
                float part[12], a_lvl[4];
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const float px = __shfl_sync(kFullMask, locv, l * 8 + p * 2);
                    const float py = __shfl_sync(kFullMask, locv, l * 8 + p * 2 + 1);
                    const float a = __shfl_sync(kFullMask, attnv, l * 4 + p);
                    const float xf = floorf(px), yf = floorf(py);
                    const float lx = px - xf, ly = py - yf;
                    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
                    const bool x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W[l]);
                    const bool x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W[l]);
                    const bool y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H[l]);
                    const bool y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H[l]);
                    const int r00 = y0 * W[l] + x0;
                    float4 v00, v01, v10, v11;
                    if (in_smem[l]) {
                        const uint32_t a00 = sbase + lvl_off[l] + static_cast<uint32_t>(r00 * kRowB);
                        const uint32_t a10 = a00 + static_cast<uint32_t>(W[l] * kRowB);
                        v00 = lds_row4_or_zero<VT>(a00, y0ok & x0ok);
                        v01 = lds_row4_or_zero<VT>(a00 + kRowB, y0ok & x1ok);
                        v10 = lds_row4_or_zero<VT>(a10, y1ok & x0ok);
                        v11 = lds_row4_or_zero<VT>(a10 + kRowB, y1ok & x1ok);
                    } else {
                        const int o00 = static_cast<int>(lvl_off[l]) + r00 * rowStride;
                        const int o10 = o00 + W[l] * rowStride;
                        v00 = ld4_or_zero(vbase + o00, y0ok & x0ok);
                        v01 = ld4_or_zero(vbase + o00 + rowStride, y0ok & x1ok);
                        v10 = ld4_or_zero(vbase + o10, y1ok & x0ok);
                        v11 = ld4_or_zero(vbase + o10 + rowStride, y1ok & x1ok);
                    }
                    const float hx = 1.f - lx, hy = 1.f - ly;
                    if (!tc[l]) {          // levels the tensor cores do not cover: vector REDs as in msda_bwd_fast_kernel
                        const int o00 = st[l] * rowStride + r00 * rowStride;
                        const int o10 = o00 + W[l] * rowStride;
                        const float ahy = a * hy, aly = a * ly;
                        float c = ahy * hx;
                        { const float4 cg = mul4(c, g); red_add4_if(gbase + o00, y0ok & x0ok, cg.x, cg.y, cg.z, cg.w); }
                        c = ahy * lx;
                        { const float4 cg = mul4(c, g); red_add4_if(gbase + o00 + rowStride, y0ok & x1ok, cg.x, cg.y, cg.z, cg.w); }
                        c = aly * hx;
                        { const float4 cg = mul4(c, g); red_add4_if(gbase + o10, y1ok & x0ok, cg.x, cg.y, cg.z, cg.w); }
                        c = aly * lx;
                        { const float4 cg = mul4(c, g); red_add4_if(gbase + o10 + rowStride, y1ok & x1ok, cg.x, cg.y, cg.z, cg.w); }
                    }
                    const float d00 = dot4(g, v00), d01 = dot4(g, v01), d10 = dot4(g, v10), d11 = dot4(g, v11);
                    part[l * 3] = hy * (hx * d00 + lx * d01) + ly * (hx * d10 + lx * d11);
                    const float gx = hy * (d01 - d00) + ly * (d11 - d10);
                    const float gy = hx * (d10 - d00) + lx * (d11 - d01);
                    part[l * 3 + 1] = FUSED ? a * gx : a * static_cast<float>(W[l]) * gx;
                    part[l * 3 + 2] = FUSED ? a * gy : a * static_cast<float>(H[l]) * gy;
                    a_lvl[l] = a;
                }
                float sum[3];
                transpose_reduce12s(part, k, sum);
                const int lvl = k >> 1;
                const bool owner = !(k & 1);
                if (FUSED) {
                    float r_a = 0.f;
#pragma unroll
                    for (int l = 0; l < L; ++l)
                        if (lvl == l) r_a = a_lvl[l];
                    float dot = owner ? r_a * sum[0] : 0.f;
#pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) dot += __shfl_xor_sync(kFullMask, dot, s);
                    sum[0] = r_a * (sum[0] - dot);
                }
                if (owner) {
                    const int si = lvl * 4 + p;
                    gattn[qm * (L * 4) + si] = from_f32<AT>(sum[0]);
                    gloc[(qm * (L * 4) + si) * 2] = from_f32<AT>(sum[1]);
                    gloc[(qm * (L * 4) + si) * 2 + 1] = from_f32<AT>(sum[2]);
                }
            }
            if (profile && blockIdx.x == 0 && tid == 0)
                atomicAdd(reinterpret_cast<unsigned long long*>(&g_bs_cycles[10]), static_cast<unsigned long long>(clock64() - t_simt));
        } else {
            // ===== builder warp: weight tiles of the covered levels + tensor-core scatter ===================================
            const uint32_t a_hi = base, a_lo = base + kOffALo, b_hi = base + kOffBHi, b_lo = base + kOffBLo, g_st = base + kOffG;
            bool first = true;              // first batch of the segment overwrites the accumulators
            bool mma_pending = false;
            const bool kProfile = profile && blockIdx.x == 0 && lane == 0;
            long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long t_prev = kProfile ? clock64() : 0;
            for (int qb = q_begin; qb < q_end; qb += 32) {
                const int q = qb + lane;
                const bool active = q < q_end;
                const int64_t qm = (static_cast<int64_t>(n) * Lq + (active ? q : q_begin)) * M + m;
                constexpr bool kGTma = sizeof(VT) == 4;   // 16-bit grad_out rows are read with plain loads below
                if (kGTma && lane == 0) {   // this batch's grad_out rows (32 queries x 32 channels of head m), 128-byte swizzle
                    mbarrier_arrive_expect_tx(bar_g, kBTile);
                    tma_load_box_2d(g_st, &gmap, bar_g, m * D, n * Lq + qb);
                }
                // the query's samples on the covered levels
                float sx[8], sy[8], sa[8];  // pixel coordinates and weights of samples (level slot, point)
#pragma unroll
                for (int li = 0; li < 2; ++li) {
                    const int l = L - 2 + li;
#pragma unroll
                    for (int pt = 0; pt < 4; ++pt) {
                        sx[li * 4 + pt] = -4.f;
                        sy[li * 4 + pt] = -4.f;
                        sa[li * 4 + pt] = 0.f;
                    }
                    if (!tc[l] || !active) continue;
                    const float wl = static_cast<float>(W[l]), hl = static_cast<float>(H[l]);
#pragma unroll
                    for (int pt = 0; pt < 4; ++pt) {
                        float lx_ = to_f32(locp[qm * (L * 8) + l * 8 + pt * 2]);
                        float ly_ = to_f32(locp[qm * (L * 8) + l * 8 + pt * 2 + 1]);
                        if (FUSED) {
                            const float* r = refp + (static_cast<int64_t>(n) * Lq + q) * (L * 2) + l * 2;
                            lx_ = __ldg(r) + lx_ / wl;
                            ly_ = __ldg(r + 1) + ly_ / hl;
                        }
                        sx[li * 4 + pt] = pixel_coord(lx_, wl);
                        sy[li * 4 + pt] = pixel_coord(ly_, hl);
                        if (!FUSED) sa[li * 4 + pt] = to_f32(attnp[qm * (L * 4) + l * 4 + pt]);
                    }
                }
                CAPE_TICK(0);               // sample loads + coordinates
                if (FUSED && active) {      // softmax over the (q, m)'s 16 logits (deformable_transformer.py:100-101)
                    float lg[16], mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        lg[i] = to_f32(attnp[qm * 16 + i]);
                        mx = fmaxf(mx, lg[i]);
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        lg[i] = expf(lg[i] - mx);
                        sum += lg[i];
                    }
#pragma unroll
                    for (int li = 0; li < 2; ++li)
#pragma unroll
                        for (int pt = 0; pt < 4; ++pt) sa[li * 4 + pt] = tc[L - 2 + li] ? lg[(L - 2 + li) * 4 + pt] / sum : 0.f;
                }
                // the tiles are free once the previous batch's MMAs have completed
                if (mma_pending) {
                    mbarrier_wait(bar_mma, mma_phase);
                    mma_phase ^= 1;
                    mma_pending = false;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                CAPE_TICK(1);               // wait for the previous batch's MMAs
#pragma unroll
                for (int i = 0; i < 16; ++i) {          // un-write last batch's entries (hi and lo tiles)
                    const uint32_t e0 = prev[i] & 0xffffu, e1 = prev[i] >> 16;
                    if (e0 != 0xffffu) {
                        sts_f32(a_hi + e0 * 4, 0.f);
                        sts_f32(a_lo + e0 * 4, 0.f);
                    }
                    if (e1 != 0xffffu) {
                        sts_f32(a_hi + e1 * 4, 0.f);
                        sts_f32(a_lo + e1 * 4, 0.f);
                    }
                }
                CAPE_TICK(2);               // zeroing
                // accumulate this query's corner weights into its column (the column has one owner: no race)
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int l = L - 2 + (s >> 2);
                    const float px = sx[s], py = sy[s], a = sa[s];
                    const float xf = floorf(px), yf = floorf(py);
                    const float lx = px - xf, ly = py - yf;
                    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
                    const bool x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W[l]);
                    const bool x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W[l]);
                    const bool y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H[l]);
                    const bool y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H[l]);
                    const int r00 = st[l] - tc_base + y0 * W[l] + x0;
                    const float hx = 1.f - lx, hy = 1.f - ly;
                    const float ahy = a * hy, aly = a * ly;
                    const bool ok[4] = {y0ok & x0ok, y0ok & x1ok, y1ok & x0ok, y1ok & x1ok};
                    const int row[4] = {r00, r00 + 1, r00 + W[l], r00 + W[l] + 1};
                    const float w[4] = {ahy * hx, ahy * lx, aly * hx, aly * lx};
                    uint32_t idx[4];
                    float cur[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {       // the 4 corners of one sample are distinct rows: independent RMWs
                        idx[c] = ok[c] ? tile_index(row[c], lane) : 0xffffu;
                        cur[c] = ok[c] ? lds_f32(a_hi + idx[c] * 4) : 0.f;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (ok[c]) sts_f32(a_hi + idx[c] * 4, cur[c] + w[c]);
                    prev[s * 2] = idx[0] | (idx[1] << 16);
                    prev[s * 2 + 1] = idx[2] | (idx[3] << 16);
                }
                CAPE_TICK(3);               // read-modify-writes
                // lo parts of the finished column (idempotent for entries hit more than once)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t e0 = prev[i] & 0xffffu, e1 = prev[i] >> 16;
                    if (e0 != 0xffffu) sts_f32(a_lo + e0 * 4, tf32_lo_part(lds_f32(a_hi + e0 * 4)));
                    if (e1 != 0xffffu) sts_f32(a_lo + e1 * 4, tf32_lo_part(lds_f32(a_hi + e1 * 4)));
                }
                CAPE_TICK(4);               // lo parts
                // G^T tile: channel rows, this query's column (zeros for a query slot past the segment)
                if (kGTma) {
                    mbarrier_wait(bar_g, g_phase);
                    g_phase ^= 1;
                }
                CAPE_TICK(5);               // wait for the grad_out tile
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (active) {
                        if (kGTma) {
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                         : "r"(g_st + lane * 128 + ((j ^ (lane & 7)) << 4))
                                         : "memory");
                        } else {
                            v = ld4(gout + qm * D + j * 4);   // 16-bit grad_out: plain loads (the fp32 tile map does not apply)
                        }
                    }
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ch = j * 4 + c;
                        const uint32_t o = static_cast<uint32_t>(ch) * 128u + ((((static_cast<uint32_t>(lane) >> 2) ^ (ch & 7)) << 4) |
                                                                               ((static_cast<uint32_t>(lane) & 3u) << 2));
                        sts_f32(b_hi + o, vv[c]);
                        sts_f32(b_lo + o, tf32_lo_part(vv[c]));
                    }
                }
                fence_proxy_async_shared();
                __syncwarp();
                CAPE_TICK(6);               // G^T tiles
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int ksteps = (min(32, q_end - qb) + 7) >> 3;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t adv = static_cast<uint64_t>(ks * 2);     // 8 tf32 = 32 bytes along K
                        const uint64_t dbh = umma_desc_sw128(b_hi) + adv, dbl = umma_desc_sw128(b_lo) + adv;
                        for (int mt = 0; mt < n_mt; ++mt) {
                            const uint32_t acc = tmem_base + mt * 32;
                            const uint64_t dah = umma_desc_sw128(a_hi + mt * 128 * 128) + adv;
                            const uint64_t dal = umma_desc_sw128(a_lo + mt * 128 * 128) + adv;
                            umma_tf32_ss(acc, dal, dbh, (first && ks == 0) ? 0u : 1u);
                            umma_tf32_ss(acc, dah, dbl, 1u);
                            umma_tf32_ss(acc, dah, dbh, 1u);
                        }
                    }
                    umma_commit_to(bar_mma);
                }
                __syncwarp();
                CAPE_TICK(7);               // MMA issue
                first = false;
                mma_pending = true;
                if (kProfile) cyc[9] += 1;
            }
            if (mma_pending) {              // accumulators complete before the segment's flush
                mbarrier_wait(bar_mma, mma_phase);
                mma_phase ^= 1;
            }
            CAPE_TICK(8);
            if (kProfile)
                for (int i = 0; i < 10; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&g_bs_cycles[i]),
                                                       static_cast<unsigned long long>(cyc[i]));
        }
        // ---- segment end: add the accumulators of the covered levels to grad_value ------------------------------------
        if (any_tc) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp < 4) {
                for (int mt = 0; mt < n_mt; ++mt) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + mt * 32 + (static_cast<uint32_t>(warp * 32) << 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int row = mt * 128 + warp * 32 + lane;
                    if (row < tc_rows) {
                        float* dst = gvalue + ((static_cast<int64_t>(n) * S + tc_base + row) * M + m) * D;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add4_if(dst + j, true, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                        __uint_as_float(r[j + 3]));
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        pos += q_end - q_begin;
    }
    __syncthreads();
    if (any_tc && warp == kSimtWarps)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemColsB) : "memory");
}

template <typename VT, typename AT, bool FUSED>
cudaError_t launch_bs_typed(const BwdArgs& a, const CUtensorMap& vmap, const CUtensorMap& gmap, int grid, int cap_rows,
                            size_t smem_bytes, long long per_cta, int use_tc, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (first_use_on_device(&configured)) {
        const cudaError_t e = cudaFuncSetAttribute(msda_bwd_staged_kernel<VT, AT, FUSED>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmemB);
        if (e != cudaSuccess) return e;
    }
    const cape_msda_dims& d = a.d;
    msda_bwd_staged_kernel<VT, AT, FUSED><<<grid, kBsThreads, smem_bytes, stream>>>(
        vmap, gmap, static_cast<const VT*>(a.grad_out), static_cast<const VT*>(a.value), a.shapes, a.starts,
        static_cast<const AT*>(a.loc), static_cast<const AT*>(a.attn), a.ref_points, a.grad_value, static_cast<AT*>(a.grad_loc),
        static_cast<AT*>(a.grad_attn), d.N, d.S, d.M, d.Lq, cap_rows, per_cta, use_tc, tuning(kTuneProfile, 0));
    return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_bs_value(const BwdArgs& a, const CUtensorMap& vmap, const CUtensorMap& gmap, int grid, int cap_rows,
                            size_t smem_bytes, long long per_cta, int use_tc, cudaStream_t stream) {
    if (a.fused) return launch_bs_typed<VT, float, true>(a, vmap, gmap, grid, cap_rows, smem_bytes, per_cta, use_tc, stream);
    if (a.aux_dtype == CAPE_DTYPE_F32)
        return launch_bs_typed<VT, float, false>(a, vmap, gmap, grid, cap_rows, smem_bytes, per_cta, use_tc, stream);
    return launch_bs_typed<VT, VT, false>(a, vmap, gmap, grid, cap_rows, smem_bytes, per_cta, use_tc, stream);
}

}  // namespace

cudaError_t read_backward_staged_cycles(long long* out16, bool reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_bs_cycles, sizeof(long long) * 16);
    if (e == cudaSuccess && reset) {
        const long long zeros[16] = {0};
        e = cudaMemcpyToSymbol(g_bs_cycles, zeros, sizeof(zeros));
    }
    return e;
}

// mode 2: staged value rows + tensor-core scatter of the coarse levels; mode 3: staged value rows only (all REDs).
// Returns cudaErrorNotSupported for configurations outside this kernel (the caller then uses msda_bwd_fast_kernel).
cudaError_t launch_backward_staged(const BwdArgs& a, int mode, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    if (d.D != 32 || d.P != 4 || d.L != 4) return cudaErrorNotSupported;
    const long long total = static_cast<long long>(d.N) * d.M * d.Lq;
    if (total < tuning(kTuneFwdStagedMinQm, 148 * 2048)) return cudaErrorNotSupported;
    const int esize = a.value_dtype == CAPE_DTYPE_F32 ? 4 : 2;
    const int row_bytes = 32 * esize;
    const int use_tc = mode == 2 ? 1 : 0;
    const int fixed = 1024 + kOffStage;            // alignment slack + the weight / G tiles (laid out in both modes)
    const int budget_kb = min(tuning(kTuneBwdStagedKb, 48), (kMaxDynSmemB - fixed) / 1024);
    int cap_rows = (budget_kb * 1024 / (kBoxRowsB * row_bytes)) * kBoxRowsB;
    if (cap_rows < 0) cap_rows = 0;
    if (d.S < cap_rows) cap_rows = (d.S + kBoxRowsB - 1) / kBoxRowsB * kBoxRowsB;
    // the kernel lays the staged rows out after the tile region whether or not the tensor-core path is used
    const size_t smem_bytes = 1024 + static_cast<size_t>(kOffStage) + static_cast<size_t>(cap_rows) * row_bytes;
    if (smem_bytes > static_cast<size_t>(kMaxDynSmemB)) return cudaErrorNotSupported;
    CUtensorMap vmap, gmap;
    if (!make_tensor_map_2d(&vmap, a.value, a.value_dtype, static_cast<uint64_t>(d.N) * d.S,
                            static_cast<uint64_t>(d.M) * d.D, kBoxRowsB, 32, CU_TENSOR_MAP_SWIZZLE_NONE))
        return cudaErrorNotSupported;
    // grad_out tile for the G^T operand (fp32 only; 16-bit grad_out is read with plain loads by the builder)
    gmap = vmap;
    if (esize == 4 && !make_tensor_map_2d(&gmap, a.grad_out, a.value_dtype, static_cast<uint64_t>(d.N) * d.Lq,
                                          static_cast<uint64_t>(d.M) * d.D, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
        return cudaErrorNotSupported;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long per_cta = (total + sms - 1) / sms;
    per_cta = (per_cta + 31) / 32 * 32;
    const int grid = static_cast<int>((total + per_cta - 1) / per_cta);
    switch (a.value_dtype) {
        case CAPE_DTYPE_F32: return launch_bs_value<float>(a, vmap, gmap, grid, cap_rows, smem_bytes, per_cta, use_tc, stream);
        case CAPE_DTYPE_BF16:
            return launch_bs_value<__nv_bfloat16>(a, vmap, gmap, grid, cap_rows, smem_bytes, per_cta, use_tc, stream);
        case CAPE_DTYPE_F16: return launch_bs_value<__half>(a, vmap, gmap, grid, cap_rows, smem_bytes, per_cta, use_tc, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cape
