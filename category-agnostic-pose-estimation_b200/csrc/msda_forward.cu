// MSDeformAttn forward (bilinear gather + weighted reduction) for sm_100a.
//
// Replaces ms_deform_attn_core_pytorch, /root/reference/models/deformable_transformer.py:115-141, and — with the
// fused prologue — the softmax / location arithmetic of MSDeformAttn.forward (:99-105) for the decode variant.
//
// Mapping (fast path, D = 32, P = 4, L <= 4 — the CAPE configuration, train_cape_episodic.py:168-188):
//   CTA   = one (n, head m) and a run of consecutive queries, so that every warp of the CTA gathers from the same
//           (n, m) value rows and neighbouring queries share their L1-resident corner rows;
//   warp  = FOUR consecutive queries at a time, one per 8-lane group: lane = (query slot g = lane>>3, channel quad
//           k = lane&7).  Each group walks its own query's 16 samples; one LDG.128 per lane fetches the same corner of
//           the 4 queries' current sample (4 x 128 B rows per warp instruction).  A group owns its whole output row, so
//           there is no cross-lane reduction at the end — the LSU wavefront pipe, which shuffles share with the row
//           gathers, is what bounds this kernel;
//   the 32 location floats and 16 weights of a (q, m) are loaded by the query's group (float4 + float2 per lane,
//           128 B + 64 B coalesced), prefetched one query quad ahead; lane k turns its 4 floats into pixel coordinates
//           once (sentinel -4 when an axis is fully out of bounds, so validity is two unsigned compares) and the
//           sample loop distributes them inside the group with 3 shuffles per sample;
//   out-of-bounds corners are predicated loads into zero-initialised registers (one asm statement per corner so
//           ptxas does not load-then-select); the loop is branch-free.
// Measured on B200 (profiles/r01_ubench_gather_scatter.txt) an SM sustains one 128 B row per ~1.7 clk from L1, so the
// 64 corner rows of a (q, m) bound this kernel well below the HBM roofline; see DESIGN.md.
// Everything else (other D / P / L) takes the generic kernel: one warp per (n, q, m), lanes strided over channels.
#include <cstdlib>
#include <type_traits>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_point.cuh"

namespace cape {

namespace {

constexpr int kFwdMaxThreads = 512;

// Build variant `--variant=fwdtable` (-DCAPE_FWD_TABLE=1): every lane converts its own two samples once and the group reads
// 16-byte records from a per-warp shared-memory table instead of 3 shuffles + the floor / bounds arithmetic per sample.
// Parity-green and 35 % fewer instructions, but SLOWER in fp32 (357 vs 313 us; bf16 294 vs 298): a 16-byte read by four
// query groups costs 4 wavefronts on the LSU pipe that bounds this kernel, the 3 shuffles cost 3
// (profiles/r02_forward_sample_table.txt).  Default stays the shuffle form.
#ifndef CAPE_FWD_TABLE
#define CAPE_FWD_TABLE 0
#endif

// Sample record written once per sample by the lane that owns it and read by the 8 lanes of its query group:
// (a, lx, ly, off | valid) — off = element offset of corner (y0, x0) inside the (n, m, quad) image, a multiple of
// rowStride >= 32, so its 4 low bits carry the in-bounds flags of the 4 corners.
__device__ __forceinline__ float4 make_sample_record(float px, float py, float a, int H, int W, int levelOff, int rowStride) {
    const float xf = floorf(px), yf = floorf(py);
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const unsigned x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W);
    const unsigned x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W);
    const unsigned y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H);
    const unsigned y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H);
    const int off = levelOff + (y0 * W + x0) * rowStride;
    const unsigned bits = (y0ok & x0ok) | ((y0ok & x1ok) << 1) | ((y1ok & x0ok) << 2) | ((y1ok & x1ok) << 3);
    return make_float4(a, px - xf, py - yf, __int_as_float(off | static_cast<int>(bits)));
}

template <typename VT>
__device__ __forceinline__ void gather_record(const VT* __restrict__ img, const float4& r, int rowStride, int downStride,
                                              float4& acc) {
    const int w = __float_as_int(r.w);
    const VT* p00 = img + (w & ~15);
    const VT* p10 = p00 + downStride;
    const float a = r.x, lx = r.y, ly = r.z;
    const float ahy = a * (1.f - ly), aly = a * ly, hx = 1.f - lx;
    const float4 v00 = ld4_or_zero(p00, w & 1);
    const float4 v01 = ld4_or_zero(p00 + rowStride, w & 2);
    const float4 v10 = ld4_or_zero(p10, w & 4);
    const float4 v11 = ld4_or_zero(p10 + rowStride, w & 8);
    fma4(ahy * hx, v00, acc);
    fma4(ahy * lx, v01, acc);
    fma4(aly * hx, v10, acc);
    fma4(aly * lx, v11, acc);
}

// The sample table of one (q, m), spread over the 8 lanes of the query's group: lane k holds location floats
// 4k .. 4k+3 (= samples 2k and 2k+1, both of level k >> 1) and weights 2k, 2k+1.
struct RawSamples {
    float4 loc;
    float2 attn;
};

__device__ __forceinline__ void load_raw(const float* locp, const float* attnp, int64_t qm, int k, int L, bool on,
                                         RawSamples& r, float pad) {
    r.loc = make_float4(0.f, 0.f, 0.f, 0.f);
    r.attn = make_float2(pad, pad);
    if (on && k < 2 * L) {
        r.loc = ld4_stream(locp + qm * (L * 8) + k * 4);
        r.attn = ld2_stream(attnp + qm * (L * 4) + k * 2);
    }
}
template <typename HT>   // bf16 / fp16 location + weight tensors
__device__ __forceinline__ void load_raw(const HT* locp, const HT* attnp, int64_t qm, int k, int L, bool on,
                                         RawSamples& r, float pad) {
    r.loc = make_float4(0.f, 0.f, 0.f, 0.f);
    r.attn = make_float2(pad, pad);
    if (on && k < 2 * L) {
        r.loc = ld4(locp + qm * (L * 8) + k * 4);
        const HT* a = attnp + qm * (L * 4) + k * 2;
        r.attn = make_float2(to_f32(a[0]), to_f32(a[1]));
    }
}

// MC: compile-time head count (row stride = MC * 32 elements) or 0 for a run-time stride.
template <typename VT, typename AT, int L, bool FUSED, int MC>
__global__ void __launch_bounds__(kFwdMaxThreads, 2)
msda_fwd_fast_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ starts, const void* __restrict__ locp,
                     const void* __restrict__ attnp, const float* __restrict__ refp, VT* __restrict__ out, int N, int S,
                     int M_rt, int Lq, int q_per_cta, int q_tiles) {
    constexpr int D = 32;
    using LT = typename std::conditional<FUSED, float, AT>::type;   // dtype of the location / weight tensors
    const int M = MC ? MC : M_rt;
    const int lane = threadIdx.x & 31, warp = uniform_warp_id(), nwarps = blockDim.x >> 5;
    const int g = lane >> 3, k = lane & 7;
    int bid = blockIdx.x;
    const int qt = bid % q_tiles;
    bid /= q_tiles;
    const int m = bid % M, n = bid / M;
    const int rowStride = M * D;
    FwdLevels<VT, L> lv;
    lv.load(shapes, starts, value + (static_cast<int64_t>(n) * S * M + m) * D + k * 4, rowStride, S);
    // dimensions of the level whose samples this lane converts (level k >> 1)
    float ownW = 1.f, ownH = 1.f;
    int ownWi = 0, ownHi = 0, ownOff = 0;
#pragma unroll
    for (int l = 0; l < L; ++l)
        if ((k >> 1) == l) {
            ownW = static_cast<float>(lv.W[l]);
            ownH = static_cast<float>(lv.H[l]);
            ownWi = lv.W[l];
            ownHi = lv.H[l];
            ownOff = lv.off[l];
        }
#if CAPE_FWD_TABLE
    // per-warp sample records: [sample 0..15][query slot g] x 16 B, so that the 4 groups' reads of one sample are 64 contiguous bytes
    __shared__ float4 sample_table[kFwdMaxThreads / 32][16][4];
    float4(*table)[4] = sample_table[warp];
#endif
    const LT* loc_t = static_cast<const LT*>(locp);
    const LT* attn_t = static_cast<const LT*>(attnp);
    const int q_end = min(Lq, (qt + 1) * q_per_cta);
    int qw = qt * q_per_cta + warp * 4;                  // first query of this warp's quad (warp-uniform)
    if (qw >= q_end) return;
    const float pad = FUSED ? -INFINITY : 0.f;
    const int grp = lane & 24;
    RawSamples cur, nxt;
    load_raw(loc_t, attn_t, (static_cast<int64_t>(n) * Lq + qw + g) * M + m, k, L, qw + g < q_end, nxt, pad);
    for (; qw < q_end; qw += nwarps * 4) {
        const int q = qw + g;                            // this group's query
        const bool on = q < q_end;
        const int64_t nq = static_cast<int64_t>(n) * Lq + q;
        cur = nxt;
        load_raw(loc_t, attn_t, (nq + nwarps * 4) * M + m, k, L, q + nwarps * 4 < q_end, nxt, pad);
        if (FUSED) {   // softmax over the group's 4L logits; loc = ref + off / (W_l, H_l)  (deformable_transformer.py:100-105)
            float mx = fmaxf(cur.attn.x, cur.attn.y);
#pragma unroll
            for (int s = 4; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, s));
            const bool has = on && k < 2 * L;
            const float e0 = has ? expf(cur.attn.x - mx) : 0.f, e1 = has ? expf(cur.attn.y - mx) : 0.f;
            float sum = e0 + e1;
#pragma unroll
            for (int s = 4; s >= 1; s >>= 1) sum += __shfl_xor_sync(kFullMask, sum, s);
            cur.attn = make_float2(e0 / sum, e1 / sum);
            if (has) {
                const float2 r = __ldg(reinterpret_cast<const float2*>(refp + nq * (L * 2)) + (k >> 1));
                cur.loc.x = r.x + cur.loc.x / ownW;
                cur.loc.y = r.y + cur.loc.y / ownH;
                cur.loc.z = r.x + cur.loc.z / ownW;
                cur.loc.w = r.y + cur.loc.w / ownH;
            }
        }
        // pixel coordinates of this lane's two samples; a group without a query samples nothing
        const float px0 = on ? pixel_coord(cur.loc.x, ownW) : -4.f, py0 = on ? pixel_coord(cur.loc.y, ownH) : -4.f;
        const float px1 = on ? pixel_coord(cur.loc.z, ownW) : -4.f, py1 = on ? pixel_coord(cur.loc.w, ownH) : -4.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#if CAPE_FWD_TABLE
        // every lane converts its own two samples once (floor, bounds, offset) instead of all 8 lanes of the group
        // repeating it for all 16; the group then reads one 16-byte record per sample
        __syncwarp();                                    // the previous quad's records have been read
        if (k < 2 * L) {
            table[2 * k][g] = make_sample_record(px0, py0, cur.attn.x, ownHi, ownWi, ownOff, rowStride);
            table[2 * k + 1][g] = make_sample_record(px1, py1, cur.attn.y, ownHi, ownWi, ownOff, rowStride);
        }
        __syncwarp();
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int downStride = lv.W[l] * rowStride;
#pragma unroll
            for (int p = 0; p < 4; ++p) gather_record(lv.img, table[l * 4 + p][g], rowStride, downStride, acc);
        }
#else
#pragma unroll
        for (int l = 0; l < L; ++l) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int s = l * 4 + p, src = grp | (s >> 1);
                const float px = __shfl_sync(kFullMask, (s & 1) ? px1 : px0, src);
                const float py = __shfl_sync(kFullMask, (s & 1) ? py1 : py0, src);
                const float a = __shfl_sync(kFullMask, (s & 1) ? cur.attn.y : cur.attn.x, src);
                gather_level(lv.img, lv.off[l], rowStride, lv.H[l], lv.W[l], px, py, a, acc);
            }
        }
#endif
        if (on) st4(out + (nq * M + m) * D + k * 4, acc);
    }
}

// Latency-oriented variant for small problems (the decode step: Lq = 1 .. a few tokens): one warp per (n, q, m), lane =
// (point p = lane>>3, channel quad k), so the 16 samples of a query are spread over the 4 lane groups and the 4 levels are
// fully unrolled — all 16 corner loads of a lane are in flight at once.  At N = 128 the value cache (713 MB) does not
// fit L2, every corner row is an HBM access, and the 4-queries-per-warp kernel above would walk its 16 samples one
// dependent round trip after the other with three quarters of its lanes idle.  Costs 8 reduction shuffles per (q, m),
// irrelevant at this size.
template <typename VT, typename AT, int L, bool FUSED>
__global__ void __launch_bounds__(128)
msda_fwd_point_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ starts, const void* __restrict__ locp,
                      const void* __restrict__ attnp, const float* __restrict__ refp, VT* __restrict__ out,
                      int64_t total_qm, int S, int M, int Lq) {
    constexpr int D = 32;
    using LT = typename std::conditional<FUSED, float, AT>::type;
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int p = lane >> 3, k = lane & 7;
    const int64_t qm = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (qm >= total_qm) return;
    const int m = static_cast<int>(qm % M);
    const int64_t nq = qm / M, n = nq / Lq;
    const float4 acc = sample_point<VT, LT, L, FUSED>(value, shapes, starts, static_cast<const LT*>(locp),
                                                       static_cast<const LT*>(attnp), refp, qm, n, m, nq, S, M, lane);
    if (p == 0) st4(out + qm * D + k * 4, acc);
}

// Generic path: any D <= 256, L <= 8, P <= 8.  One warp per (n, q, m); lanes stride over d.
template <typename VT, typename AT, bool FUSED>
__global__ void __launch_bounds__(128)
msda_fwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ starts, const void* __restrict__ locp,
                        const void* __restrict__ attnp, const float* __restrict__ refp, VT* __restrict__ out,
                        int64_t total_qm, int S, int M, int D, int Lq, int L, int P) {
    const int lane = threadIdx.x & 31;
    const int64_t qm = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qm >= total_qm) return;
    const int m = static_cast<int>(qm % M);
    const int64_t nq = qm / M;
    const int64_t n = nq / Lq;
    const int LP = L * P;
    float mx = -INFINITY, denom = 1.f;
    if (FUSED) {   // softmax statistics over the L*P logits (deformable_transformer.py:100-101)
        const float* g = static_cast<const float*>(attnp) + qm * LP;
        for (int i = 0; i < LP; ++i) mx = fmaxf(mx, __ldg(g + i));
        denom = 0.f;
        for (int i = 0; i < LP; ++i) denom += expf(__ldg(g + i) - mx);
    }
    constexpr int kChunks = 8;   // D <= 256
    float acc[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) acc[c] = 0.f;
    for (int l = 0; l < L; ++l) {
        int H = static_cast<int>(__ldg(shapes + 2 * l)), W = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const long long start64 = __ldg(starts + l);
        if (start64 < 0 || H < 0 || W < 0 || start64 + static_cast<long long>(H) * W > S) H = W = 0;   // level outside S: skipped
        const int start = static_cast<int>(start64);
        for (int p = 0; p < P; ++p) {
            float locx, locy, a;
            const int64_t si = qm * LP + l * P + p;
            if (FUSED) {
                const float* o = static_cast<const float*>(locp);
                locx = __ldg(refp + (nq * L + l) * 2) + __ldg(o + si * 2) / static_cast<float>(W);
                locy = __ldg(refp + (nq * L + l) * 2 + 1) + __ldg(o + si * 2 + 1) / static_cast<float>(H);
                a = expf(__ldg(static_cast<const float*>(attnp) + si) - mx) / denom;
            } else {
                const AT* o = static_cast<const AT*>(locp);
                locx = to_f32(o[si * 2]);
                locy = to_f32(o[si * 2 + 1]);
                a = to_f32(static_cast<const AT*>(attnp)[si]);
            }
            int x0, y0;
            float lx, ly;
            if (!sample_coords(locx, locy, H, W, x0, y0, lx, ly)) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int xi = x0 + (c & 1), yi = y0 + (c >> 1);
                if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
                const float w = a * ((c >> 1) ? ly : 1.f - ly) * ((c & 1) ? lx : 1.f - lx);
                const VT* v = value + ((n * S + start + yi * W + xi) * M + m) * D;
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    const int d = lane + ch * 32;
                    if (d < D) acc[ch] = fmaf(w, to_f32(v[d]), acc[ch]);
                }
            }
        }
    }
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
        const int d = lane + ch * 32;
        if (d < D) out[qm * D + d] = from_f32<VT>(acc[ch]);
    }
}

// Launch geometry of the fast path.  Default: 512-thread CTAs (16 warps) over 512 consecutive queries — two CTAs per
// SM at 64 registers, each sweeping a compact window of the (n, m) value rows.  The FWD_THREADS / FWD_QPC tuning knobs
// (msda_launch.h) override it for tuning runs.
struct FastGeometry {
    int threads, q_per_cta, q_tiles;
    int64_t grid;
};

FastGeometry fast_geometry(const cape_msda_dims& d, Tune knob_threads, Tune knob_qpc, int def_threads, int def_qpc) {
    int threads = tuning(knob_threads, def_threads);
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > kFwdMaxThreads) threads = kFwdMaxThreads;
    int q_per_cta = tuning(knob_qpc, 0);
    if (q_per_cta <= 0)   // default: whole waves of the 2 x 148 resident 512-thread CTAs (an explicit FWD_QPC is taken as is)
        q_per_cta = balanced_q_per_cta(static_cast<int64_t>(d.N) * d.M, d.Lq, def_qpc, 2 * 148, 4);
    // small problems (decode: Lq = 1..k): shrink the tile until the grid covers the chip a few times over
    while (q_per_cta > 1 && static_cast<int64_t>(d.N) * d.M * ((d.Lq + q_per_cta - 1) / q_per_cta) < 148 * 2) q_per_cta >>= 1;
    if (q_per_cta > d.Lq) q_per_cta = d.Lq;
    if (q_per_cta < 1) q_per_cta = 1;
    if (threads > ((q_per_cta + 3) / 4) * 32) threads = ((q_per_cta + 3) / 4) * 32;   // a warp takes 4 queries at a time
    FastGeometry g;
    g.threads = threads;
    g.q_per_cta = q_per_cta;
    g.q_tiles = (d.Lq + q_per_cta - 1) / q_per_cta;
    g.grid = static_cast<int64_t>(d.N) * d.M * g.q_tiles;
    return g;
}

template <typename VT, typename AT, bool FUSED, int L>
void launch_fast(const FwdArgs& a, const FastGeometry& g, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    const VT* value = static_cast<const VT*>(a.value);
    VT* out = static_cast<VT*>(a.out);
    const dim3 grid(static_cast<unsigned>(g.grid)), block(g.threads);
    if (d.M == 8)
        msda_fwd_fast_kernel<VT, AT, L, FUSED, 8><<<grid, block, 0, stream>>>(
            value, a.shapes, a.starts, a.loc, a.attn, a.ref_points, out, d.N, d.S, d.M, d.Lq, g.q_per_cta, g.q_tiles);
    else
        msda_fwd_fast_kernel<VT, AT, L, FUSED, 0><<<grid, block, 0, stream>>>(
            value, a.shapes, a.starts, a.loc, a.attn, a.ref_points, out, d.N, d.S, d.M, d.Lq, g.q_per_cta, g.q_tiles);
}

template <typename VT, typename AT, bool FUSED>
cudaError_t launch_typed(const FwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    const int64_t total_qm = static_cast<int64_t>(d.N) * d.Lq * d.M;
    if (total_qm == 0) return cudaSuccess;
    if (d.D == 32 && d.P == 4 && d.L >= 1 && d.L <= 4 && total_qm <= tuning(kTuneFwdPointMaxQm, 148 * 64)) {
        // small problem (decode): one warp per (n, q, m), everything in flight at once
        const int warps = 4;
        const unsigned grid = static_cast<unsigned>((total_qm + warps - 1) / warps);
        const VT* value = static_cast<const VT*>(a.value);
        VT* out = static_cast<VT*>(a.out);
#define CAPE_POINT_CASE(LL)                                                                                          \
    case LL:                                                                                                         \
        msda_fwd_point_kernel<VT, AT, LL, FUSED><<<grid, warps * 32, 0, stream>>>(                                   \
            value, a.shapes, a.starts, a.loc, a.attn, a.ref_points, out, total_qm, d.S, d.M, d.Lq);                    \
        break;
        switch (d.L) {
            CAPE_POINT_CASE(1)
            CAPE_POINT_CASE(2)
            CAPE_POINT_CASE(3)
            CAPE_POINT_CASE(4)
        }
#undef CAPE_POINT_CASE
    } else if (d.D == 32 && d.P == 4 && d.L >= 1 && d.L <= 4) {
        const FastGeometry g = fast_geometry(d, kTuneFwdThreads, kTuneFwdQpc, 512, 512);
        if (g.grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        switch (d.L) {
            case 1: launch_fast<VT, AT, FUSED, 1>(a, g, stream); break;
            case 2: launch_fast<VT, AT, FUSED, 2>(a, g, stream); break;
            case 3: launch_fast<VT, AT, FUSED, 3>(a, g, stream); break;
            default: launch_fast<VT, AT, FUSED, 4>(a, g, stream); break;
        }
    } else {
        const int warps = 4;
        const int64_t grid = (total_qm + warps - 1) / warps;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        msda_fwd_generic_kernel<VT, AT, FUSED><<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
            static_cast<const VT*>(a.value), a.shapes, a.starts, a.loc, a.attn, a.ref_points, static_cast<VT*>(a.out),
            total_qm, d.S, d.M, d.D, d.Lq, d.L, d.P);
    }
    count_launch();
    return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_value_typed(const FwdArgs& a, cudaStream_t stream) {
    if (a.fused) return launch_typed<VT, float, true>(a, stream);
    if (a.aux_dtype == CAPE_DTYPE_F32) return launch_typed<VT, float, false>(a, stream);
    return launch_typed<VT, VT, false>(a, stream);   // aux dtype == value dtype (validated by the ABI layer)
}

}  // namespace

cudaError_t launch_forward(const FwdArgs& a, cudaStream_t stream) {
    // Opt-in (FWD_STAGED = 1): coarse levels staged in shared memory by the TMA (msda_forward_staged.cu).  Measured on B200
    // at the bench shape it is 5 % SLOWER than the L1 kernel below (350 vs 331 us: the staged rows take the L1 capacity
    // that level 0 needs, and LDS rows cost the same data-stage wavefronts as L1 hits; profiles/r02_forward_staged.txt).
    if (tuning(kTuneFwdStaged, 2) == 1) {
        const cudaError_t e = launch_forward_staged(a, stream);
        if (e == cudaSuccess) count_launch();
        if (e != cudaErrorNotSupported) return e;
    }
    switch (a.value_dtype) {
        case CAPE_DTYPE_F32: return launch_value_typed<float>(a, stream);
        case CAPE_DTYPE_BF16: return launch_value_typed<__nv_bfloat16>(a, stream);
        case CAPE_DTYPE_F16: return launch_value_typed<__half>(a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cape
