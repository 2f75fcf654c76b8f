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
//   the 4 point-partials are folded with two xor-shuffles and lanes 0-7 store the 128 B output row.
// Everything else (other D / P / L) takes the generic kernel: one warp per (n, q, m), lanes strided over channels.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kFwdWarps = 8;

template <typename VT>
__device__ __forceinline__ void gather_level(const VT* __restrict__ vbase, int rowStride, int H, int W, int start,
                                             float locx, float locy, float a, float4& acc) {
    int x0, y0;
    float lx, ly;
    if (!sample_coords(locx, locy, H, W, x0, y0, lx, ly)) return;
    const float hx = 1.f - lx, hy = 1.f - ly;
    const bool x0ok = x0 >= 0, x1ok = x0 + 1 < W, y0ok = y0 >= 0, y1ok = y0 + 1 < H;
    const VT* p00 = vbase + static_cast<int64_t>(start + y0 * W + x0) * rowStride;
    const VT* p10 = p00 + static_cast<int64_t>(W) * rowStride;
    float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
    if (y0ok && x0ok) v00 = ld4(p00);
    if (y0ok && x1ok) v01 = ld4(p00 + rowStride);
    if (y1ok && x0ok) v10 = ld4(p10);
    if (y1ok && x1ok) v11 = ld4(p10 + rowStride);
    const float w00 = a * hy * hx, w01 = a * hy * lx, w10 = a * ly * hx, w11 = a * ly * lx;
    acc.x = fmaf(w00, v00.x, fmaf(w01, v01.x, fmaf(w10, v10.x, fmaf(w11, v11.x, acc.x))));
    acc.y = fmaf(w00, v00.y, fmaf(w01, v01.y, fmaf(w10, v10.y, fmaf(w11, v11.y, acc.y))));
    acc.z = fmaf(w00, v00.z, fmaf(w01, v01.z, fmaf(w10, v10.z, fmaf(w11, v11.z, acc.z))));
    acc.w = fmaf(w00, v00.w, fmaf(w01, v01.w, fmaf(w10, v10.w, fmaf(w11, v11.w, acc.w))));
}

// Per-(q, m) sample table held one float per lane: lane i < 8L holds loc[i] (= [l][p][xy]), lane i < 4L holds attn[i].
template <typename AT, int L, bool FUSED>
struct SampleTable {
    float loc, attn;
    // raw (not yet transformed) prefetch of the next query
    __device__ __forceinline__ void fetch(const void* locp, const void* attnp, const float* refp, int64_t qm, int64_t nq,
                                          int lane) {
        loc = 0.f;
        attn = FUSED ? -INFINITY : 0.f;
        if (FUSED) {
            const float* o = static_cast<const float*>(locp) + qm * (L * 8);
            const float* g = static_cast<const float*>(attnp) + qm * (L * 4);
            if (lane < L * 8) loc = __ldg(o + lane);
            if (lane < L * 4) attn = __ldg(g + lane);
            (void)refp;
            (void)nq;
        } else {
            const AT* o = static_cast<const AT*>(locp) + qm * (L * 8);
            const AT* g = static_cast<const AT*>(attnp) + qm * (L * 4);
            if (lane < L * 8) loc = to_f32(o[lane]);
            if (lane < L * 4) attn = to_f32(g[lane]);
        }
    }
    // FUSED: softmax over the 4L logits and loc = ref + off / (W_l, H_l)   (deformable_transformer.py:100-105)
    __device__ __forceinline__ void finish(const float* refp, int64_t nq, int lane, const Levels<L>& lv) {
        if (!FUSED) return;
        float mx = attn;
#pragma unroll
        for (int s = 8; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, s));
        float e = (lane < L * 4) ? expf(attn - mx) : 0.f;
        float sum = e;
#pragma unroll
        for (int s = 8; s >= 1; s >>= 1) sum += __shfl_xor_sync(kFullMask, sum, s);
        attn = e / sum;
        if (lane < L * 8) {
            const int l = lane >> 3, c = lane & 1;
            int dim = 1;
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (i == l) dim = c ? lv.H[i] : lv.W[i];
            const float r = __ldg(refp + nq * (L * 2) + l * 2 + c);
            loc = r + loc / static_cast<float>(dim);
        }
    }
};

template <typename VT, typename AT, int L, bool FUSED>
__global__ void __launch_bounds__(kFwdWarps * 32)
msda_fwd_fast_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ starts, const void* __restrict__ locp,
                     const void* __restrict__ attnp, const float* __restrict__ refp, VT* __restrict__ out, int N, int S,
                     int M, int Lq, int q_per_cta, int q_tiles) {
    constexpr int D = 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int p = lane >> 3, k = lane & 7;
    int bid = blockIdx.x;
    const int qt = bid % q_tiles;
    bid /= q_tiles;
    const int m = bid % M, n = bid / M;
    Levels<L> lv;
    lv.load(shapes, starts);
    const int rowStride = M * D;
    const VT* vbase = value + (static_cast<int64_t>(n) * S * M + m) * D + k * 4;
    const int q_end = min(Lq, (qt + 1) * q_per_cta);
    int q = qt * q_per_cta + warp;
    if (q >= q_end) return;
    SampleTable<AT, L, FUSED> cur, nxt;
    nxt.fetch(locp, attnp, refp, (static_cast<int64_t>(n) * Lq + q) * M + m, static_cast<int64_t>(n) * Lq + q, lane);
    for (; q < q_end; q += nwarps) {
        const int64_t nq = static_cast<int64_t>(n) * Lq + q;
        cur = nxt;
        if (q + nwarps < q_end) nxt.fetch(locp, attnp, refp, (nq + nwarps) * M + m, nq + nwarps, lane);
        cur.finish(refp, nq, lane, lv);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const float locx = __shfl_sync(kFullMask, cur.loc, l * 8 + p * 2);
            const float locy = __shfl_sync(kFullMask, cur.loc, l * 8 + p * 2 + 1);
            const float a = __shfl_sync(kFullMask, cur.attn, l * 4 + p);
            gather_level(vbase, rowStride, lv.H[l], lv.W[l], lv.start[l], locx, locy, a, acc);
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

// Generic path: any D (multiple of 4 not required here), L <= 8, P <= 8.  One warp per (n, q, m); lanes stride over d.
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

template <typename VT, typename AT, bool FUSED>
cudaError_t launch_typed(const FwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    const int64_t total_qm = static_cast<int64_t>(d.N) * d.Lq * d.M;
    if (total_qm == 0) return cudaSuccess;
    const VT* value = static_cast<const VT*>(a.value);
    VT* out = static_cast<VT*>(a.out);
    if (d.D == 32 && d.P == 4 && d.L >= 1 && d.L <= 4) {
        // queries per CTA: enough CTAs to fill 148 SMs several times over, enough queries per warp to amortise setup
        int q_per_cta = 64;
        while (q_per_cta > 8 && static_cast<int64_t>(d.N) * d.M * ((d.Lq + q_per_cta - 1) / q_per_cta) < 148 * 8) q_per_cta >>= 1;
        if (q_per_cta > d.Lq) q_per_cta = d.Lq;
        const int q_tiles = (d.Lq + q_per_cta - 1) / q_per_cta;
        const int warps = q_per_cta < kFwdWarps ? q_per_cta : kFwdWarps;
        const int64_t grid = static_cast<int64_t>(d.N) * d.M * q_tiles;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const dim3 g(static_cast<unsigned>(grid)), b(warps * 32);
#define CAPE_FWD_CASE(LL)                                                                                           \
    case LL:                                                                                                        \
        msda_fwd_fast_kernel<VT, AT, LL, FUSED><<<g, b, 0, stream>>>(value, a.shapes, a.starts, a.loc, a.attn,       \
                                                                     a.ref_points, out, d.N, d.S, d.M, d.Lq,        \
                                                                     q_per_cta, q_tiles);                           \
        break;
        switch (d.L) {
            CAPE_FWD_CASE(1)
            CAPE_FWD_CASE(2)
            CAPE_FWD_CASE(3)
            CAPE_FWD_CASE(4)
        }
#undef CAPE_FWD_CASE
    } else {
        const int warps = 4;
        const int64_t grid = (total_qm + warps - 1) / warps;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        msda_fwd_generic_kernel<VT, AT, FUSED><<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
            value, a.shapes, a.starts, a.loc, a.attn, a.ref_points, out, total_qm, d.S, d.M, d.D, d.Lq, d.L, d.P);
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
