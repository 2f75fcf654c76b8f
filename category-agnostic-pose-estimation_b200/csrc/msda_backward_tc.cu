// MSDeformAttn backward, small-CTA form, with the grad_value scatter of the COARSEST pyramid level on the tensor cores.
//
// Derivative of ms_deform_attn_core_pytorch (/root/reference/models/deformable_transformer.py:129-141); same per-sample
// arithmetic and CTA shape as msda_bwd_fast_kernel (msda_backward.cu): CTA = one (image, head) and a run of consecutive
// queries, 8 warps, a warp takes one query at a time, lane = (point, channel quad).  That kernel retires its 44 M
// `red.global.add.v4.f32` rows at 92 % of what the L2's reduction units can do; the only way down is fewer
// rows leaving the SM.  Here the last level (CAPE: 8 x 8 = 64 pixels, 25 % of all samples) never issues a RED per sample:
//
//   per query      the warp gathers / reduces levels 0 .. L-2 exactly as before, and for level L-1 writes ITS COLUMN of
//                  the batch's bilinear weight matrix Wt[pixel][query] (64 rows of 128 B = 32 query columns, 128-byte
//                  swizzle; fp32 tile + tf32 "lo" tile) — 16 lanes, one per (point, corner), `red.shared.add.f32` because
//                  corners of different points coincide — and its column of [G_hi | G_lo]^T (lane -> one channel).
//   per 32 queries warp 0 waits until the 8 warps have written their 4 columns each (mbarrier) and issues
//                      acc[pixel][0:64] += Wt_hi . [G_hi | G_lo]     (tcgen05.mma.kind::tf32, M = 64, N = 64, K = 8, x4)
//                      acc[pixel][0:32] += Wt_lo . G_hi              (N = 32)
//                  i.e. 3xTF32 into 64 tensor-memory columns; tcgen05.commit -> mbarrier releases the tiles.  Warp 0 does
//                  this right before ITS first tile write of the next batch, so nobody idles at a CTA-wide barrier.
//   CTA end        warps 0..3 read the accumulators (tcgen05.ld; M = 64 occupies lanes 0-15 of every 32-lane quadrant) and
//                  add them to grad_value with one RED per pixel row and CTA (64 rows instead of ~12 per query).
//
// Shared memory: 24 KB per CTA (4 CTAs / SM keep 130 KB of L1 for the gathers); tensor memory: 64 columns per CTA.
// A last level that does not fit 64 pixels keeps its REDs (the tiles are then unused).
#include "async_copy.cuh"
#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma_tf32.cuh"

namespace cape {

namespace {

constexpr int kTcThreads = 256;
constexpr int kTcWarps = kTcThreads / 32;
constexpr int kTcBatch = 32;                     // queries per MMA batch = the 32 fp32 columns of one 128-byte tile row
constexpr int kTcPerWarp = kTcBatch / kTcWarps;  // columns a warp writes per batch
constexpr int kTcRowsMax = 64;                   // pixels of the covered level (M = 64)
constexpr int kTcATile = kTcRowsMax * 128;       // bytes of one weight tile
constexpr int kTcBTile = 64 * 128;               // [G_hi rows 0..31 | G_lo rows 32..63][32 query columns]
constexpr int kTcTmemCols = 64;

__device__ __forceinline__ void transpose_reduce12t(const float (&v)[12], int k, float (&out)[3]) {
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

// L = 4 levels max (template), P = 4, D = 32.  FUSED as in msda_bwd_fast_kernel.
template <typename VT, typename AT, int L, bool FUSED>
__global__ void __launch_bounds__(kTcThreads, 4)
msda_bwd_tc_kernel(const VT* __restrict__ gout, const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                   const int64_t* __restrict__ starts, const AT* __restrict__ locp, const AT* __restrict__ attnp,
                   const float* __restrict__ refp, float* __restrict__ gvalue, AT* __restrict__ gloc,
                   AT* __restrict__ gattn, int N, int S, int M, int Lq, int q_per_cta, int q_tiles, int dbg) {
    constexpr int D = 32;
    __shared__ __align__(1024) unsigned char tiles[2 * kTcATile + kTcBTile];
    __shared__ __align__(8) unsigned long long bar_storage[2];
    __shared__ uint32_t tmem_slot;
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int p = lane >> 3, k = lane & 7;
    int bid = blockIdx.x;
    const int qt = bid % q_tiles;
    bid /= q_tiles;
    const int m = bid % M, n = bid / M;
    const int rowStride = M * D;

    int H[L], W[L], off[L];
    int st_last = 0;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        H[l] = static_cast<int>(__ldg(shapes + 2 * l));
        W[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const long long s0 = __ldg(starts + l);
        off[l] = static_cast<int>(s0) * rowStride;
        if (l == L - 1) st_last = static_cast<int>(s0);
        // a level outside S is skipped (W = 0: no corner passes the range test), as in msda_bwd_fast_kernel
        if (s0 < 0 || H[l] < 0 || W[l] < 0 || s0 + static_cast<long long>(H[l]) * W[l] > S) W[l] = 0;
    }
    const int tc_rows = H[L - 1] * W[L - 1];
    const bool tc_on = W[L - 1] > 0 && tc_rows <= kTcRowsMax;     // CTA-uniform

    const uint32_t a_hi = smem_addr_u32(tiles), a_lo = a_hi + kTcATile, b_tile = a_hi + 2 * kTcATile;
    const uint32_t bar_mma = smem_addr_u32(&bar_storage[0]), bar_full = smem_addr_u32(&bar_storage[1]);
    uint32_t tmem_base = 0;
    if (tc_on) {
        for (int i = threadIdx.x; i < (2 * kTcATile + kTcBTile) / 16; i += kTcThreads)
            reinterpret_cast<float4*>(tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (threadIdx.x == 0) {
            mbarrier_init(bar_mma, 1);
            mbarrier_init(bar_full, kTcWarps);
            mbarrier_init_fence();
        }
        if (warp == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(&tmem_slot)),
                         "n"(kTcTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_proxy_async_shared();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_base = tmem_slot;
    }

    const int64_t headOff = (static_cast<int64_t>(n) * S * M + m) * D + k * 4;
    const VT* vbase = value + headOff;
    float* gbase = gvalue + headOff;
    float dimf = 1.f;
#pragma unroll
    for (int l = 0; l < L; ++l)
        if ((lane >> 3) == l) dimf = static_cast<float>((lane & 1) ? H[l] : W[l]);

    const int q_begin = qt * q_per_cta, q_end = min(Lq, q_begin + q_per_cta);
    const int n_batches = (q_end - q_begin + kTcBatch - 1) / kTcBatch;
    const int n_iter = tc_on ? n_batches * kTcPerWarp : (q_end - q_begin + kTcWarps - 1) / kTcWarps;
    uint32_t prev0 = 0xffffu, prev1 = 0xffffu, prev2 = 0xffffu, prev3 = 0xffffu;   // tile entry this lane wrote in column j of the previous batch

    for (int i = 0; i < n_iter; ++i) {
        const int q = q_begin + i * kTcWarps + warp;
        const bool active = q < q_end;
        float e_px = -4.f, e_py = -4.f, e_a = 0.f;      // this lane's sample of the covered level (point p)
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            const int64_t qm = (static_cast<int64_t>(n) * Lq + q) * M + m;
            float locv = 0.f, attnv = FUSED ? -INFINITY : 0.f;
            if (lane < L * 8) locv = ld1_stream(locp + qm * (L * 8) + lane);
            if (lane < L * 4) attnv = ld1_stream(attnp + qm * (L * 4) + lane);
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
                if (lane < L * 8)
                    locv = __ldg(refp + (static_cast<int64_t>(n) * Lq + q) * (L * 2) + (lane >> 3) * 2 + (lane & 1)) + locv / dimf;
            }
            locv = pixel_coord(locv, dimf);
            float part[12], a_lvl[4];
#pragma unroll
            for (int j = 0; j < 12; ++j) part[j] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) a_lvl[j] = 0.f;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const float px = __shfl_sync(kFullMask, locv, l * 8 + p * 2);
                const float py = __shfl_sync(kFullMask, locv, l * 8 + p * 2 + 1);
                const float a = __shfl_sync(kFullMask, attnv, l * 4 + p);
                if (l == L - 1) {
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
                const int o00 = (y0 * W[l] + x0) * rowStride + off[l];
                const int o10 = o00 + W[l] * rowStride;
                const float4 v00 = ld4_or_zero(vbase + o00, y0ok & x0ok);
                const float4 v01 = ld4_or_zero(vbase + o00 + rowStride, y0ok & x1ok);
                const float4 v10 = ld4_or_zero(vbase + o10, y1ok & x0ok);
                const float4 v11 = ld4_or_zero(vbase + o10 + rowStride, y1ok & x1ok);
                const float hx = 1.f - lx, hy = 1.f - ly;
                if (l < L - 1 || !tc_on) {          // levels the tensor cores do not cover: vector REDs
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
            transpose_reduce12t(part, k, sum);
            const int lvl = k >> 1;
            const bool owner = !(k & 1) && lvl < L;
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
        if (tc_on && dbg != 1) {
            // ---- this query's column of the batch's operand tiles: column = (i % 4) * 8 + warp ----
            const int slot = i & (kTcPerWarp - 1), batch = i / kTcPerWarp;
            const int col = slot * kTcWarps + warp;
            if (slot == 0 && batch > 0 && dbg != 2) {
                if (warp == 0) {            // the previous batch: all 32 columns written -> issue its MMAs
                    mbarrier_wait(bar_full, (batch - 1) & 1);
                    if (lane == 0) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int ks = 0; ks < kTcBatch / 8; ++ks) {
                            const uint64_t adv = static_cast<uint64_t>(ks * 2);     // 8 tf32 = 32 bytes along K
                            const uint64_t db = umma_desc_sw128(b_tile) + adv;
                            umma_tf32_ss(tmem_base, umma_desc_sw128(a_hi) + adv, db, kIdescM64N64, (batch == 1 && ks == 0) ? 0u : 1u);
                            umma_tf32_ss(tmem_base, umma_desc_sw128(a_lo) + adv, db, kIdescM64N32, 1u);
                        }
                        umma_commit_to(bar_mma);
                    }
                    __syncwarp();
                }
                if (dbg != 3) mbarrier_wait(bar_mma, (batch - 1) & 1);           // the previous batch's MMAs have read the tiles
            }
            const int c = k & 3;                // lanes with k < 4: corner c of point p; lanes k >= 4 idle here
            const float xf = floorf(e_px), yf = floorf(e_py);
            const float lx = e_px - xf, ly = e_py - yf;
            const int xc = static_cast<int>(xf) + (c & 1), yc = static_cast<int>(yf) + (c >> 1);
            const bool valid = active && k < 4 && static_cast<unsigned>(xc) < static_cast<unsigned>(W[L - 1]) &&
                               static_cast<unsigned>(yc) < static_cast<unsigned>(H[L - 1]);
            const float wgt = e_a * ((c & 2) ? ly : 1.f - ly) * ((c & 1) ? lx : 1.f - lx);
            const uint32_t idx = valid ? tile_index(yc * W[L - 1] + xc, col) : 0xffffu;
            const uint32_t prev = slot == 0 ? prev0 : (slot == 1 ? prev1 : (slot == 2 ? prev2 : prev3));
            if (prev != 0xffffu) {
                sts_f32(a_hi + prev * 4, 0.f);
                sts_f32(a_lo + prev * 4, 0.f);
            }
            __syncwarp();
            // corners of different points often coincide: shared-memory float add (a CAS loop on sm_100, ATOMS.CAST.SPIN)
            if (valid) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a_hi + idx * 4), "f"(wgt) : "memory");
            __syncwarp();
            if (valid) sts_f32(a_lo + idx * 4, tf32_lo_part(lds_f32(a_hi + idx * 4)));
            if (slot == 0) prev0 = idx;
            if (slot == 1) prev1 = idx;
            if (slot == 2) prev2 = idx;
            if (slot == 3) prev3 = idx;
            {   // this query's column of the transposed grad_out operand: lane -> channel 4k + p (hi row, lo row); zeros if inactive
                const int ch = k * 4 + p;
                const float gv = p == 0 ? g.x : (p == 1 ? g.y : (p == 2 ? g.z : g.w));
                const uint32_t o = static_cast<uint32_t>(ch) * 128u + ((((static_cast<uint32_t>(col) >> 2) ^ (ch & 7)) << 4) |
                                                                       ((static_cast<uint32_t>(col) & 3u) << 2));
                sts_f32(b_tile + o, gv);
                sts_f32(b_tile + 32 * 128 + o, tf32_lo_part(gv));
            }
            if (slot == kTcPerWarp - 1) {
                fence_proxy_async_shared();
                __syncwarp();
                if (lane == 0) mbarrier_arrive(bar_full);
            }
        }
    }

    if (tc_on) {
        // ---- last batch's MMAs, then one RED per covered pixel row ----
        if (n_batches > 0 && dbg != 1 && dbg != 2) {
            if (warp == 0) {
                mbarrier_wait(bar_full, (n_batches - 1) & 1);
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int ks = 0; ks < kTcBatch / 8; ++ks) {
                        const uint64_t adv = static_cast<uint64_t>(ks * 2);
                        const uint64_t db = umma_desc_sw128(b_tile) + adv;
                        umma_tf32_ss(tmem_base, umma_desc_sw128(a_hi) + adv, db, kIdescM64N64, (n_batches == 1 && ks == 0) ? 0u : 1u);
                        umma_tf32_ss(tmem_base, umma_desc_sw128(a_lo) + adv, db, kIdescM64N32, 1u);
                    }
                    umma_commit_to(bar_mma);
                }
                __syncwarp();
            }
            if (warp < 4) {
                mbarrier_wait(bar_mma, (n_batches - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int row = warp * 16 + lane;             // M = 64: rows 16w .. 16w+15 live in lanes 0..15 of quadrant w
                float* dst = gvalue + ((static_cast<int64_t>(n) * S + st_last + row) * M + m) * D;
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 16) {       // columns c0..c0+15 (hi*hi + lo*hi) and 32+c0.. (hi*lo)
                    uint32_t r[16], t[16];
                    const uint32_t taddr = tmem_base + c0 + (static_cast<uint32_t>(warp * 32) << 16);
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
                    if (lane < 16 && row < tc_rows) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            red_add4_if(dst + c0 + j, true, __uint_as_float(r[j]) + __uint_as_float(t[j]),
                                        __uint_as_float(r[j + 1]) + __uint_as_float(t[j + 1]),
                                        __uint_as_float(r[j + 2]) + __uint_as_float(t[j + 2]),
                                        __uint_as_float(r[j + 3]) + __uint_as_float(t[j + 3]));
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        }
        __syncthreads();
        if (warp == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcTmemCols) : "memory");
    }
}

template <typename VT, typename AT, bool FUSED>
cudaError_t launch_tc_typed(const BwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    int q_per_cta = tuning(kTuneBwdQpc, 0);
    if (q_per_cta <= 0) q_per_cta = balanced_q_per_cta(static_cast<int64_t>(d.N) * d.M, d.Lq, 256, 4 * 148, kTcBatch);
    q_per_cta = (q_per_cta + kTcBatch - 1) / kTcBatch * kTcBatch;
    const int q_tiles = (d.Lq + q_per_cta - 1) / q_per_cta;
    const int64_t grid = static_cast<int64_t>(d.N) * d.M * q_tiles;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
#define CAPE_TC_CASE(LL)                                                                                                     \
    case LL:                                                                                                                 \
        msda_bwd_tc_kernel<VT, AT, LL, FUSED><<<static_cast<unsigned>(grid), kTcThreads, 0, stream>>>(                       \
            static_cast<const VT*>(a.grad_out), static_cast<const VT*>(a.value), a.shapes, a.starts,                         \
            static_cast<const AT*>(a.loc), static_cast<const AT*>(a.attn), a.ref_points, a.grad_value,                       \
            static_cast<AT*>(a.grad_loc), static_cast<AT*>(a.grad_attn), d.N, d.S, d.M, d.Lq, q_per_cta, q_tiles,            \
            tuning(kTuneProfile, 0));                                                                                        \
        break;
    switch (d.L) {
        CAPE_TC_CASE(2)
        CAPE_TC_CASE(3)
        CAPE_TC_CASE(4)
        default: return cudaErrorNotSupported;
    }
#undef CAPE_TC_CASE
    return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_tc_value(const BwdArgs& a, cudaStream_t stream) {
    if (a.fused) return launch_tc_typed<VT, float, true>(a, stream);
    if (a.aux_dtype == CAPE_DTYPE_F32) return launch_tc_typed<VT, float, false>(a, stream);
    return launch_tc_typed<VT, VT, false>(a, stream);
}

}  // namespace

// mode 5: small CTAs, coarsest level scattered on the tensor cores.  cudaErrorNotSupported outside D = 32, P = 4, 2 <= L <= 4
// or for problems too small to fill the chip (the caller then uses msda_bwd_fast_kernel).
cudaError_t launch_backward_tc(const BwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    if (d.D != 32 || d.P != 4 || d.L < 2 || d.L > 4) return cudaErrorNotSupported;
    const long long total = static_cast<long long>(d.N) * d.M * d.Lq;
    if (total < tuning(kTuneBwdTcMinQm, 148 * 1024)) return cudaErrorNotSupported;
    switch (a.value_dtype) {
        case CAPE_DTYPE_F32: return launch_tc_value<float>(a, stream);
        case CAPE_DTYPE_BF16: return launch_tc_value<__nv_bfloat16>(a, stream);
        case CAPE_DTYPE_F16: return launch_tc_value<__half>(a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cape
