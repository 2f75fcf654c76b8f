// MSDeformAttn forward (bilinear gather + weighted reduction) for sm_100a.
//
// Replaces ms_deform_attn_core_pytorch, /root/reference/models/deformable_transformer.py:115-141, and — with the
// fused prologue — the softmax / location arithmetic of MSDeformAttn.forward (:99-105) for the decode variant.
//
// Mapping (fast path, D = 32, P = 4, L <= 4 — the CAPE configuration, train_cape_episodic.py:168-188):
//   CTA   = one (n, head m) and a run of consecutive queries, so that every warp of the CTA gathers from the same
//           (n, m) value rows and neighbouring queries share their L1-resident corner rows;
//   warp  = one query at a time; its 32 lanes are (point p = lane>>3, channel quad k = lane&7): one 16-byte load per
//           lane fetches the same corner of the 4 points of one level (4 x 128 B rows per warp instruction);
//   the 32 location floats and 16 weights of a (q, m) are one coalesced 128 B + 64 B load, distributed by shuffles,
//   and prefetched one query ahead;
//   the level loop is branch-free (out-of-bounds corners are predicated loads of zero), so the shuffles need no
//   reconvergence and the 16 corner loads of a query are all in flight together;
//   the 4 point-partials are folded with two xor-shuffles and lanes 0-7 store the 128 B output row.
// Measured on B200 (profiles/r01_ubench_gather_scatter.txt) an SM sustains one 128 B row per ~1.7 clk from L1, so the
// 64 corner rows of a (q, m) bound this kernel well below the HBM roofline; see DESIGN.md.
// Everything else (other D / P / L) takes the generic kernel: one warp per (n, q, m), lanes strided over channels.
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kFwdMaxThreads = 1024;

// Per-level constants kept in registers for the whole CTA.
template <typename VT, int L>
struct FwdLevels {
    int H[L], W[L];
    const VT* base[L];   // first row of the level for this (n, m, channel quad)
    __device__ __forceinline__ void load(const int64_t* __restrict__ shapes, const int64_t* __restrict__ starts,
                                         const VT* vbase, int rowStride) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            H[l] = static_cast<int>(__ldg(shapes + 2 * l));
            W[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
            base[l] = vbase + static_cast<int64_t>(__ldg(starts + l)) * rowStride;
        }
    }
    // map dimension that loc float `lane` (= [l][p][xy]) is scaled by: W_l for x, H_l for y
    __device__ __forceinline__ float lane_dim(int lane) const {
        int dim = 1;
#pragma unroll
        for (int l = 0; l < L; ++l)
            if ((lane >> 3) == l) dim = (lane & 1) ? H[l] : W[l];
        return static_cast<float>(dim);
    }
};

// One level, 4 points (one per 8-lane group), 4 corners each.  (px, py) are pixel coordinates from pixel_coord().
template <typename VT>
__device__ __forceinline__ void gather_level(const VT* __restrict__ base, int rowStride, int H, int W, float px,
                                             float py, float a, float4& acc) {
    const float xf = floorf(px), yf = floorf(py);
    const float lx = px - xf, ly = py - yf;
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const bool x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W);
    const bool x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W);
    const bool y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H);
    const bool y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H);
    const VT* p00 = base + static_cast<int64_t>(y0 * W + x0) * rowStride;
    const VT* p10 = p00 + static_cast<int64_t>(W) * rowStride;
    const float ahy = a * (1.f - ly), aly = a * ly, hx = 1.f - lx;
    const float4 v00 = ld4_or_zero(p00, y0ok & x0ok);
    const float4 v01 = ld4_or_zero(p00 + rowStride, y0ok & x1ok);
    const float4 v10 = ld4_or_zero(p10, y1ok & x0ok);
    const float4 v11 = ld4_or_zero(p10 + rowStride, y1ok & x1ok);
    fma4(ahy * hx, v00, acc);
    fma4(ahy * lx, v01, acc);
    fma4(aly * hx, v10, acc);
    fma4(aly * lx, v11, acc);
}

// Per-(q, m) sample table held one float per lane: lane i < 8L holds loc[i] (= [l][p][xy]), lane i < 4L holds attn[i].
// finish() turns the raw prefetch into pixel coordinates (and, FUSED, first applies the softmax over the 4L logits and
// loc = ref + off / (W_l, H_l), deformable_transformer.py:100-105) — once per (q, m), one coordinate per lane.
template <typename AT, int L, bool FUSED>
struct SampleTable {
    float loc, attn;
    __device__ __forceinline__ void fetch(const void* locp, const void* attnp, int64_t qm, int lane) {
        loc = 0.f;
        attn = FUSED ? -INFINITY : 0.f;
        if (FUSED) {
            const float* o = static_cast<const float*>(locp) + qm * (L * 8);
            const float* g = static_cast<const float*>(attnp) + qm * (L * 4);
            if (lane < L * 8) loc = __ldg(o + lane);
            if (lane < L * 4) attn = __ldg(g + lane);
        } else {
            const AT* o = static_cast<const AT*>(locp) + qm * (L * 8);
            const AT* g = static_cast<const AT*>(attnp) + qm * (L * 4);
            if (lane < L * 8) loc = to_f32(o[lane]);
            if (lane < L * 4) attn = to_f32(g[lane]);
        }
    }
    __device__ __forceinline__ void finish(const float* refp, int64_t nq, int lane, float dimf) {
        if (FUSED) {
            float mx = attn;
#pragma unroll
            for (int s = 8; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, s));
            const float e = (lane < L * 4) ? expf(attn - mx) : 0.f;
            float sum = e;
#pragma unroll
            for (int s = 8; s >= 1; s >>= 1) sum += __shfl_xor_sync(kFullMask, sum, s);
            attn = e / sum;
            if (lane < L * 8) loc = __ldg(refp + nq * (L * 2) + (lane >> 3) * 2 + (lane & 1)) + loc / dimf;
        }
        loc = pixel_coord(loc, dimf);
    }
};

// MC: compile-time head count (row stride = MC * 32 elements) or 0 for a run-time stride.
template <typename VT, typename AT, int L, bool FUSED, int MC>
__global__ void __launch_bounds__(kFwdMaxThreads)
msda_fwd_fast_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ starts, const void* __restrict__ locp,
                     const void* __restrict__ attnp, const float* __restrict__ refp, VT* __restrict__ out, int N, int S,
                     int M_rt, int Lq, int q_per_cta, int q_tiles) {
    constexpr int D = 32;
    const int M = MC ? MC : M_rt;
    const int lane = threadIdx.x & 31, warp = uniform_warp_id(), nwarps = blockDim.x >> 5;
    const int p = lane >> 3, k = lane & 7;
    int bid = blockIdx.x;
    const int qt = bid % q_tiles;
    bid /= q_tiles;
    const int m = bid % M, n = bid / M;
    const int rowStride = M * D;
    FwdLevels<VT, L> lv;
    lv.load(shapes, starts, value + (static_cast<int64_t>(n) * S * M + m) * D + k * 4, rowStride);
    const float dimf = lv.lane_dim(lane);
    const int q_end = min(Lq, (qt + 1) * q_per_cta);
    int q = qt * q_per_cta + warp;
    if (q >= q_end) return;
    SampleTable<AT, L, FUSED> cur, nxt;
    nxt.fetch(locp, attnp, (static_cast<int64_t>(n) * Lq + q) * M + m, lane);
    for (; q < q_end; q += nwarps) {
        const int64_t nq = static_cast<int64_t>(n) * Lq + q;
        cur = nxt;
        if (q + nwarps < q_end) nxt.fetch(locp, attnp, (nq + nwarps) * M + m, lane);
        cur.finish(refp, nq, lane, dimf);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const float px = __shfl_sync(kFullMask, cur.loc, l * 8 + p * 2);
            const float py = __shfl_sync(kFullMask, cur.loc, l * 8 + p * 2 + 1);
            const float a = __shfl_sync(kFullMask, cur.attn, l * 4 + p);
            gather_level(lv.base[l], rowStride, lv.H[l], lv.W[l], px, py, a, acc);
        }
#pragma unroll
        for (int s = 8; s <= 16; s <<= 1) {
            acc.x += __shfl_xor_sync(kFullMask, acc.x, s);
            acc.y += __shfl_xor_sync(kFullMask, acc.y, s);
            acc.z += __shfl_xor_sync(kFullMask, acc.z, s);
            acc.w += __shfl_xor_sync(kFullMask, acc.w, s);
        }
        if (p == 0) st4(out + (nq * M + m) * D + k * 4, acc);
    }
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
        const int H = static_cast<int>(__ldg(shapes + 2 * l)), W = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const int start = static_cast<int>(__ldg(starts + l));
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
// SM at 64 registers, each sweeping a compact window of the (n, m) value rows.  CAPE_FWD_THREADS / CAPE_FWD_QPC
// override it for tuning runs.
struct FastGeometry {
    int threads, q_per_cta, q_tiles;
    int64_t grid;
};

int env_int(const char* name, int fallback) {
    const char* v = std::getenv(name);
    if (!v || !*v) return fallback;
    const int x = std::atoi(v);
    return x > 0 ? x : fallback;
}

FastGeometry fast_geometry(const cape_msda_dims& d, const char* env_threads, const char* env_qpc, int def_threads,
                           int def_qpc) {
    int threads = env_int(env_threads, def_threads);
    threads = (threads / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > kFwdMaxThreads) threads = kFwdMaxThreads;
    int q_per_cta = env_int(env_qpc, def_qpc);
    // small problems (decode: Lq = 1..k): shrink the tile until the grid covers the chip a few times over
    while (q_per_cta > 1 && static_cast<int64_t>(d.N) * d.M * ((d.Lq + q_per_cta - 1) / q_per_cta) < 148 * 4) q_per_cta >>= 1;
    if (q_per_cta > d.Lq) q_per_cta = d.Lq;
    if (q_per_cta < 1) q_per_cta = 1;
    if (threads > q_per_cta * 32) threads = q_per_cta * 32;
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
    if (d.D == 32 && d.P == 4 && d.L >= 1 && d.L <= 4) {
        const FastGeometry g = fast_geometry(d, "CAPE_FWD_THREADS", "CAPE_FWD_QPC", 512, 512);
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
    switch (a.value_dtype) {
        case CAPE_DTYPE_F32: return launch_value_typed<float>(a, stream);
        case CAPE_DTYPE_BF16: return launch_value_typed<__nv_bfloat16>(a, stream);
        case CAPE_DTYPE_F16: return launch_value_typed<__half>(a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cape
