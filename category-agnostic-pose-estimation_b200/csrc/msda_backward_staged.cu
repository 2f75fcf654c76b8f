// MSDeformAttn backward with the two coarsest pyramid levels kept on the SM: their value rows are staged in shared
// memory by the TMA (gathers become LDS.128) and their grad_value scatter runs on the tensor cores instead of the L2
// reduction units.
//
// Derivative of ms_deform_attn_core_pytorch (/root/reference/models/deformable_transformer.py:129-141); same per-sample
// arithmetic as msda_bwd_fast_kernel (msda_backward.cu).  What bounds that kernel is the 44 M `red.global.add.v4.f32` rows
// per launch at the L2's ~51 G rows/s, with the LSU wavefront pipe right behind.  Here, per CTA
// (persistent, one per SM, a contiguous range of the (image, head, query) space):
//
//   warps 0..30  sample: one query per warp at a time (batch b of a segment = 31 consecutive queries, warp w takes the
//                w-th), lane = (point, channel quad).  Levels staged in shared memory are gathered with LDS.128, the others
//                with LDG.128 as before; grad_loc / grad_attn for all levels; REDs into grad_value only for the levels the
//                tensor cores do NOT cover.  For the covered levels the warp then writes ITS COLUMN of the batch's
//                bilinear weight matrix Wt[pixel][query] (rows of 128 B = 32 query columns, 128-byte swizzle): lane =
//                (point, level slot, corner) adds A * w_corner with `red.shared.add.f32` (corners of different points
//                often coincide), re-reads the sum and stores its tf32 "lo" part in a second tile.  Entries written for
//                the previous batch are zeroed first, so the tiles are never cleared wholesale.
//                It also writes its query's column of the transposed grad_out operand G[channel | channel_lo][query]
//                (lane -> one channel, hi row + lo row).
//   warp 31      per batch: waits until all 31 columns are in (mbarrier, one arrival per sampling warp), and lane 0 issues
//                    acc[pixel][0:64]  += Wt_hi . [G_hi | G_lo]        (tcgen05.mma.kind::tf32, M = 128, N = 64, K = 8)
//                    acc[pixel][0:32]  += Wt_lo . G_hi                 (N = 32)
//                i.e. 3xTF32 (hi*hi + hi*lo + lo*hi) in two instructions per (8 queries, 128 pixels), into accumulators in
//                TENSOR MEMORY that live for the whole (image, head) segment; tcgen05.commit -> mbarrier tells the sampling
//                warps when the tiles may be modified again (they are busy sampling the next batch meanwhile).
//   segment end  warps 0..3 read the accumulators (tcgen05.ld), add the two column halves and add the result to grad_value
//                with one RED per row and CTA.
//
// Covered levels = the last one or two levels whose pixels fit 384 rows (the CAPE pyramid: 16x16 + 8x8 = 320 pixels,
// 50 % of all samples, ~21 of the ~51 in-bounds corner rows of a (query, head)).
#include "async_copy.cuh"
#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma_tf32.cuh"

namespace cape {

namespace {

#ifndef CAPE_BS_THREADS
#define CAPE_BS_THREADS 1024
#endif
constexpr int kBsThreads = CAPE_BS_THREADS;       // one CTA per SM; -DCAPE_BS_THREADS=896|768 builds trade warps for registers
constexpr int kSimtWarps = kBsThreads / 32 - 1;
constexpr int kTcRows = 384;                      // pixels covered by the tensor-core scatter: 3 tiles of 128
constexpr int kATile = kTcRows * 128;             // bytes of Wt[pixel][32 query columns] fp32 (rows of 128 B, 128-byte swizzle)
constexpr int kBTile = 64 * 128;                  // G[channel (hi) | channel (lo)][32 query columns]
constexpr int kOffALo = kATile, kOffB = 2 * kATile;
constexpr int kOffStage = kOffB + kBTile;         // staged value rows start here (1024-byte aligned)
constexpr int kBoxRowsB = 64;
constexpr int kTmemColsB = 256;                   // 3 accumulators x 64 columns, rounded up to a power of two
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
    __shared__ __align__(8) uint64_t bar_storage[4];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_addr_u32(smem_raw) + 1023u) & ~1023u;       // swizzle atoms need 1024-byte alignment
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
    const int rowStride = M * D;
    const uint32_t bar_stage = smem_addr_u32(&bar_storage[0]), bar_g = smem_addr_u32(&bar_storage[1]),
                   bar_mma = smem_addr_u32(&bar_storage[2]), bar_full = smem_addr_u32(&bar_storage[3]);

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
    const bool drop_only = use_tc == 2;      // PROFILING ONLY (BWD_MODE 4): covered levels' REDs dropped, nothing replaces them
    const bool any_tc = tc_rows > 0 && !drop_only;
    const int n_mt = (tc_rows + 127) >> 7;
    bool in_smem[L];
    uint32_t lvl_off[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        in_smem[l] = rows > 0 && H[l] > 0 && st[l] >= base_row;
        lvl_off[l] = in_smem[l] ? static_cast<uint32_t>(st[l] - base_row) * kRowB : static_cast<uint32_t>(st[l]) * rowStride;
    }

    // ---- one-time set-up: barriers, tensor memory, zeroed weight / operand tiles ---------------------------------------
    if (tid == 0) {
        mbarrier_init(bar_stage, 1);
        mbarrier_init(bar_g, 1);
        mbarrier_init(bar_mma, 1);
        mbarrier_init(bar_full, kSimtWarps);
        mbarrier_init_fence();
    }
    if (any_tc) {
        for (int i = tid; i < kOffStage / 16; i += kBsThreads)
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
    const uint32_t a_hi = base, a_lo = base + kOffALo, b_tile = base + kOffB;

    const long long total = static_cast<long long>(N) * M * Lq;
    long long pos = static_cast<long long>(blockIdx.x) * per_cta;
    const long long end = min(total, pos + per_cta);
    const int nsimt = any_tc ? kSimtWarps : kBsThreads / 32;
    uint32_t stage_phase = 0;
    uint32_t gb = 0;                         // batches completed by this CTA (same value in every warp): barrier phases
    uint32_t prev_idx = 0xffffu;             // sampling lanes: tile entry written for the previous batch
    const bool kProfile = profile && blockIdx.x == 0 && lane == 0;
    long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = kProfile ? clock64() : 0;

    while (pos < end) {
        const int nm = static_cast<int>(pos / Lq);
        const int q_begin = static_cast<int>(pos - static_cast<long long>(nm) * Lq);
        const int q_end = static_cast<int>(min(static_cast<long long>(Lq), q_begin + (end - pos)));
        const int n = nm / M, m = nm - n * M;
        const int nb = (q_end - q_begin + nsimt - 1) / nsimt;
        __syncthreads();                     // previous segment: rows read, accumulators flushed
        if (rows > 0) {
            if (warp == 0) {
                if (lane == 0) mbarrier_arrive_expect_tx(bar_stage, static_cast<uint32_t>(nboxes) * kBoxRowsB * kRowB);
                __syncwarp();
                for (int b = lane; b < nboxes; b += 32)
                    tma_load_box_2d(base + kOffStage + b * kBoxRowsB * kRowB, &vmap, bar_stage, m * D,
                                    n * S + base_row + b * kBoxRowsB);
            }
            if (warp != kSimtWarps || !any_tc) mbarrier_wait(bar_stage, stage_phase);   // the MMA warp never reads value rows
            stage_phase ^= 1;
        }

        if (warp < kSimtWarps || !any_tc) {
            // ===== sampling warps ==========================================================================================
            const int p = lane >> 3, k = lane & 7;
            const int64_t headOff = (static_cast<int64_t>(n) * S * M + m) * D + k * 4;
            const VT* vbase = value + headOff;
            float* gbase = gvalue + headOff;
            const uint32_t sbase = base + kOffStage + k * (kRowB / 8);
            float dimf = 1.f;
#pragma unroll
            for (int l = 0; l < L; ++l)
                if ((lane >> 3) == l) dimf = static_cast<float>((lane & 1) ? H[l] : W[l]);
            // this lane's entry of the weight tile: point p, level slot k >> 2 (of the two covered levels), corner k & 3
            const int li = k >> 2;
            const bool e_tc = li ? tc[L - 1] : tc[L - 2];
            const int e_W = li ? W[L - 1] : W[L - 2], e_H = li ? H[L - 1] : H[L - 2];
            const int e_row0 = (li ? st[L - 1] : st[L - 2]) - tc_base;
            for (int i = 0; i < nb; ++i) {
                const int q = q_begin + i * nsimt + warp;
                const bool active = q < q_end;
                float e_px = -4.f, e_py = -4.f, e_a = 0.f;      // this lane's sample of the covered levels (point p, level slot li)
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) {
                    const int64_t qm = (static_cast<int64_t>(n) * Lq + q) * M + m;
                    float locv = to_f32(locp[qm * (L * 8) + lane]);
                    float attnv = FUSED ? -INFINITY : 0.f;
                    if (lane < L * 4) attnv = to_f32(attnp[qm * (L * 4) + lane]);
                    g = ld4(gout + qm * D + k * 4);
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
                        if (l >= L - 2 && li == l - (L - 2)) {
                            e_px = px;
                            e_py = py;
                            e_a = a;
                        }
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
                if (any_tc) {
                    // ---- this warp's column (index = warp) of the batch's weight tiles ----
                    const float xf = floorf(e_px), yf = floorf(e_py);
                    const float lx = e_px - xf, ly = e_py - yf;
                    const int xc = static_cast<int>(xf) + (k & 1), yc = static_cast<int>(yf) + ((k >> 1) & 1);
                    const bool valid = e_tc && static_cast<unsigned>(xc) < static_cast<unsigned>(e_W) &&
                                       static_cast<unsigned>(yc) < static_cast<unsigned>(e_H);
                    const float wgt = e_a * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lx : 1.f - lx);
                    const uint32_t idx = valid ? tile_index(e_row0 + yc * e_W + xc, warp) : 0xffffu;
                    if (gb > 0) mbarrier_wait(bar_mma, (gb - 1) & 1);      // the previous batch's MMAs have read the tiles
                    if (prev_idx != 0xffffu) {
                        sts_f32(a_hi + prev_idx * 4, 0.f);
                        sts_f32(a_lo + prev_idx * 4, 0.f);
                    }
                    __syncwarp();
                    // corners of different points often coincide: shared-memory float add (a CAS loop on sm_100, ATOMS.CAST.SPIN).
                    // Summing coinciding lanes in registers first (match.any + shuffles, plain stores) measured the same: 955 vs 943 us.
                    if (valid) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a_hi + idx * 4), "f"(wgt) : "memory");
                    __syncwarp();
                    if (valid) sts_f32(a_lo + idx * 4, tf32_lo_part(lds_f32(a_hi + idx * 4)));
                    prev_idx = idx;
                    {   // this query's column of the transposed grad_out operand: lane -> channel 4k + p (hi row, lo row)
                        const int ch = k * 4 + p;
                        const float gv = p == 0 ? g.x : (p == 1 ? g.y : (p == 2 ? g.z : g.w));
                        const uint32_t o = static_cast<uint32_t>(ch) * 128u + ((((static_cast<uint32_t>(warp) >> 2) ^ (ch & 7)) << 4) |
                                                                               ((static_cast<uint32_t>(warp) & 3u) << 2));
                        sts_f32(b_tile + o, gv);
                        sts_f32(b_tile + 32 * 128 + o, tf32_lo_part(gv));
                    }
                    fence_proxy_async_shared();
                    __syncwarp();
                    if (lane == 0) mbarrier_arrive(bar_full);
                    ++gb;
                }
            }
        } else {
            // ===== MMA warp: transposed grad_out operand + tensor-core scatter of the batch ================================
            bool first = true;              // first batch of the segment overwrites the accumulators
            for (int i = 0; i < nb; ++i) {
                const int qb = q_begin + i * kSimtWarps;
                const int cnt = min(kSimtWarps, q_end - qb);
                CAPE_TICK(0);
                mbarrier_wait(bar_full, gb & 1);            // all 31 columns of the weight / grad_out tiles are in
                CAPE_TICK(4);
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int ksteps = (cnt + 7) >> 3;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t adv = static_cast<uint64_t>(ks * 2);     // 8 tf32 = 32 bytes along K
                        const uint64_t db = umma_desc_sw128(b_tile) + adv;
                        for (int mt = 0; mt < n_mt; ++mt) {
                            const uint32_t acc = tmem_base + mt * 64;
                            const uint64_t dah = umma_desc_sw128(a_hi + mt * 128 * 128) + adv;
                            const uint64_t dal = umma_desc_sw128(a_lo + mt * 128 * 128) + adv;
                            umma_tf32_ss(acc, dah, db, kIdescM128N64, (first && ks == 0) ? 0u : 1u);   // hi*hi | hi*lo
                            umma_tf32_ss(acc, dal, db, kIdescM128N32, 1u);                             // + lo*hi
                        }
                    }
                    umma_commit_to(bar_mma);
                }
                __syncwarp();
                CAPE_TICK(5);
                first = false;
                ++gb;
                if (kProfile) cyc[9] += 1;
            }
            if (nb > 0) {                   // accumulators complete before the segment's flush
                mbarrier_wait(bar_mma, (gb - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            CAPE_TICK(6);
        }
        // ---- segment end: add the accumulators of the covered levels to grad_value ------------------------------------
        if (any_tc) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp < 4) {
                for (int mt = 0; mt < n_mt; ++mt) {
                    const int row = mt * 128 + warp * 32 + lane;
                    float* dst = gvalue + ((static_cast<int64_t>(n) * S + tc_base + row) * M + m) * D;
#pragma unroll 1
                    for (int c0 = 0; c0 < 32; c0 += 16) {       // columns c0..c0+15 (hi*hi + lo*hi) and 32+c0.. (hi*lo)
                        uint32_t r[16], t[16];
                        const uint32_t taddr = tmem_base + mt * 64 + c0 + (static_cast<uint32_t>(warp * 32) << 16);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(taddr));
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]),
                              "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
                            : "r"(taddr + 32));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (row < tc_rows) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                red_add4_if(dst + c0 + j, true, __uint_as_float(r[j]) + __uint_as_float(t[j]),
                                            __uint_as_float(r[j + 1]) + __uint_as_float(t[j + 1]),
                                            __uint_as_float(r[j + 2]) + __uint_as_float(t[j + 2]),
                                            __uint_as_float(r[j + 3]) + __uint_as_float(t[j + 3]));
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        pos += q_end - q_begin;
    }
    if (kProfile)
        for (int i = 0; i < 10; ++i)
            atomicAdd(reinterpret_cast<unsigned long long*>(&g_bs_cycles[i]), static_cast<unsigned long long>(cyc[i]));
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
    const int use_tc = mode == 2 ? 1 : (mode == 4 ? 2 : 0);
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
    gmap = vmap;   // (second map parameter kept for ABI stability of the kernel signature; unused)
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
