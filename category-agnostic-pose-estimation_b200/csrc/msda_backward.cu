// MSDeformAttn backward (scatter of grad_value + grad_sampling_loc + grad_attn_weight in one pass) for sm_100a.
//
// Derivative of ms_deform_attn_core_pytorch, /root/reference/models/deformable_transformer.py:129-141, i.e. what
// autograd produces through grid_sampler_2d_backward.  Per sample with weight A, corner weights w_c = wx_c*wy_c and
// G = grad_out[n, q, m, :]:
//     grad_attn      = sum_c w_c <G, v_c>
//     grad_loc.x     = A * W_l * sum_c (dx ? +wy_c : -wy_c) <G, v_c>
//     grad_loc.y     = A * H_l * sum_c (dy ? +wx_c : -wx_c) <G, v_c>
//     grad_value[c] += A * w_c * G                 (in-bounds corners only)
//
// Fast path (D = 32, P = 4, L <= 4): CTA = one (n, head) and a run of consecutive queries; a warp takes one query at a
// time, lane = (point p = lane>>3, channel quad k = lane&7).  Per level: 4 predicated corner loads (4 x 128 B rows per
// warp instruction), 4 dot products with grad_out's 4 channels, 4 predicated vector reductions into grad_value
// (REDG.E.ADD.F32x4: 8 lanes cover one 128 B corner row).  The 12 per-lane partials (4 levels x (grad_attn, grad_loc.x,
// grad_loc.y)) are summed over the 8 lanes of a point with a 12-shuffle transpose-reduce after the level loop; lane
// (p, k even) then owns sample (level k/2, point p) and writes its grad_loc / grad_attn entries.
// The scatter is what bounds this kernel: 44 M corner rows per launch at the L2 reduction units' ~51 G rows/s
// (DESIGN.md §5).
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kBwdMaxThreads = 512;

template <typename VT, int L>
struct BwdLevels {
    int H[L], W[L];
    int off[L];   // element offset of the level's first row inside this batch element's (S, M, D) block (< 2^31)
    __device__ __forceinline__ void load(const int64_t* __restrict__ shapes, const int64_t* __restrict__ starts,
                                         int rowStride, int S) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            H[l] = static_cast<int>(__ldg(shapes + 2 * l));
            W[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
            const long long s0 = __ldg(starts + l);
            off[l] = static_cast<int>(s0) * rowStride;
            // a level that does not fit inside S (the reference asserts sum(H*W) == S, deformable_transformer.py:94) is
            // skipped: with W = 0 no corner passes the x-range test, so nothing is gathered or scattered for it
            if (s0 < 0 || H[l] < 0 || W[l] < 0 || s0 + static_cast<long long>(H[l]) * W[l] > S) W[l] = 0;
        }
    }
    __device__ __forceinline__ float lane_dim(int lane) const {
        int dim = 1;
#pragma unroll
        for (int l = 0; l < L; ++l)
            if ((lane >> 3) == l) dim = (lane & 1) ? H[l] : W[l];
        return static_cast<float>(dim);
    }
};

// One level, 4 points (one per 8-lane group).  (px, py) are pixel coordinates from pixel_coord().  Branch-free:
// out-of-bounds corners load zeros and their REDG is predicated off.
template <typename VT>
__device__ __forceinline__ void scatter_level(const VT* __restrict__ vbase, float* __restrict__ gbase, int rowStride,
                                              int levelOff, int H, int W, float px, float py, float a,
                                              const float4& g, float& ga, float& gx, float& gy) {
    const float xf = floorf(px), yf = floorf(py);
    const float lx = px - xf, ly = py - yf;
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const bool x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W);
    const bool x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W);
    const bool y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H);
    const bool y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H);
    const int o00 = (y0 * W + x0) * rowStride + levelOff;   // 32-bit element offsets: S*M*D < 2^31 (checked by the ABI)
    const int o10 = o00 + W * rowStride;
#ifdef CAPE_EXP_NO_LOAD   // profiling-only variant (tools/, never shipped): drop the corner gathers
    const float4 v00 = make_float4(px, py, a, 1.f), v01 = v00, v10 = v00, v11 = v00;
#else
    const float4 v00 = ld4_or_zero(vbase + o00, y0ok & x0ok);
    const float4 v01 = ld4_or_zero(vbase + o00 + rowStride, y0ok & x1ok);
    const float4 v10 = ld4_or_zero(vbase + o10, y1ok & x0ok);
    const float4 v11 = ld4_or_zero(vbase + o10 + rowStride, y1ok & x1ok);
#endif
    const float hx = 1.f - lx, hy = 1.f - ly;
    const float ahy = a * hy, aly = a * ly;
#ifndef CAPE_EXP_NO_RED   // profiling-only variant (tools/, never shipped) drops the scatter
#ifdef CAPE_EXP_NO_COARSE_RED
    if (gbase != nullptr) {
#else
    {
#endif
    float c = ahy * hx;
    { const float4 cg = mul4(c, g); red_add4_if(gbase + o00, y0ok & x0ok, cg.x, cg.y, cg.z, cg.w); }
    c = ahy * lx;
    { const float4 cg = mul4(c, g); red_add4_if(gbase + o00 + rowStride, y0ok & x1ok, cg.x, cg.y, cg.z, cg.w); }
    c = aly * hx;
    { const float4 cg = mul4(c, g); red_add4_if(gbase + o10, y1ok & x0ok, cg.x, cg.y, cg.z, cg.w); }
    c = aly * lx;
    { const float4 cg = mul4(c, g); red_add4_if(gbase + o10 + rowStride, y1ok & x1ok, cg.x, cg.y, cg.z, cg.w); }
    }
#else
    (void)gbase;
    (void)ahy;
    (void)aly;
#endif
    const float d00 = dot4(g, v00), d01 = dot4(g, v01), d10 = dot4(g, v10), d11 = dot4(g, v11);   // 0 for OOB corners
    ga = hy * (hx * d00 + lx * d01) + ly * (hx * d10 + lx * d11);
    gx = hy * (d01 - d00) + ly * (d11 - d10);
    gy = hx * (d10 - d00) + lx * (d11 - d01);
}

// Sum 12 per-lane partials over the 8 lanes of a point group with 12 shuffles instead of 36: two halving rounds
// (xor 4, xor 2: each lane keeps the half it will own and sends the other), then one full round (xor 1).
// v = [level][ga, gx, gy]; afterwards every lane holds the three complete sums of level (k >> 1) in out[0..2].
__device__ __forceinline__ void transpose_reduce12(const float (&v)[12], int k, float (&out)[3]) {
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

// FUSED = false: locp / attnp are sampling_locations / attention_weights, outputs grad_loc / grad_attn.
// FUSED = true : the module prologue is part of the op (deformable_transformer.py:100-105): locp / attnp are the raw
//                sampling offsets and attention logits (fp32), refp the reference points; the kernel recomputes
//                attn = softmax(logits) and loc = ref + off / (W_l, H_l), and writes grad_offsets = grad_loc / (W_l, H_l)
//                and grad_logits = attn * (grad_attn - sum_j attn_j grad_attn_j) instead, so sampling_locations,
//                attention_weights and their gradients never exist in HBM.
template <typename VT, typename AT, int L, int MC, bool FUSED>
__global__ void __launch_bounds__(kBwdMaxThreads, 2)
msda_bwd_fast_kernel(const VT* __restrict__ gout, const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ starts, const AT* __restrict__ locp, const AT* __restrict__ attnp,
                     const float* __restrict__ refp, float* __restrict__ gvalue, AT* __restrict__ gloc,
                     AT* __restrict__ gattn, int N, int S, int M_rt, int Lq, int q_per_cta, int q_tiles) {
    constexpr int D = 32;
    const int M = MC ? MC : M_rt;
    const int lane = threadIdx.x & 31, warp = uniform_warp_id(), nwarps = blockDim.x >> 5;
    const int p = lane >> 3, k = lane & 7;
    int bid = blockIdx.x;
    const int qt = bid % q_tiles;
    bid /= q_tiles;
    const int m = bid % M, n = bid / M;
    const int rowStride = M * D;
    BwdLevels<VT, L> lv;
    lv.load(shapes, starts, rowStride, S);
    const int64_t headOff = (static_cast<int64_t>(n) * S * M + m) * D + k * 4;
    const VT* vbase = value + headOff;
    float* gbase = gvalue + headOff;
    const float dimf = lv.lane_dim(lane);
    const int q_end = min(Lq, (qt + 1) * q_per_cta);
    for (int q = qt * q_per_cta + warp; q < q_end; q += nwarps) {
        const int64_t qm = (static_cast<int64_t>(n) * Lq + q) * M + m;
        float locv = 0.f, attnv = FUSED ? -INFINITY : 0.f;
        if (lane < L * 8) locv = ld1_stream(locp + qm * (L * 8) + lane);      // touched once: no L1 allocation
        if (lane < L * 4) attnv = ld1_stream(attnp + qm * (L * 4) + lane);
        const float4 g = ld4(gout + qm * D + k * 4);
        if (FUSED) {   // softmax over the 4L logits (lanes 0 .. 4L-1) and loc = ref + off / dim
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
        for (int i = 0; i < 12; ++i) part[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a_lvl[i] = 0.f;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const float px = __shfl_sync(kFullMask, locv, l * 8 + p * 2);
            const float py = __shfl_sync(kFullMask, locv, l * 8 + p * 2 + 1);
            const float a = __shfl_sync(kFullMask, attnv, l * 4 + p);
            float ga, gx, gy;
#ifdef CAPE_EXP_NO_COARSE_RED   // profiling-only variant (tools/, never shipped): levels in the mask gather but do not scatter
            scatter_level(vbase, ((CAPE_EXP_NO_COARSE_RED >> l) & 1) ? nullptr : gbase, rowStride, lv.off[l], lv.H[l], lv.W[l], px, py, a, g, ga, gx, gy);
#else
            scatter_level(vbase, gbase, rowStride, lv.off[l], lv.H[l], lv.W[l], px, py, a, g, ga, gx, gy);
#endif
            part[l * 3] = ga;
            // d loc / d offset = 1 / dim cancels the dim factor of d pixel / d loc in the fused form
            part[l * 3 + 1] = FUSED ? a * gx : a * static_cast<float>(lv.W[l]) * gx;
            part[l * 3 + 2] = FUSED ? a * gy : a * static_cast<float>(lv.H[l]) * gy;
            a_lvl[l] = a;
        }
        float sum[3];
        transpose_reduce12(part, k, sum);      // every lane now holds (ga, gx, gy) of level k >> 1 for its point
        const int lvl = k >> 1;
        const bool owner = !(k & 1) && lvl < L;
        if (FUSED) {   // softmax backward over the (q, m)'s 4L samples, held by the owner lanes
            float r_a = 0.f;
#pragma unroll
            for (int l = 0; l < L; ++l)
                if (lvl == l) r_a = a_lvl[l];
            float dot = owner ? r_a * sum[0] : 0.f;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) dot += __shfl_xor_sync(kFullMask, dot, s);
            sum[0] = r_a * (sum[0] - dot);
        }
        if (owner) {   // lane (p, k even) owns sample (level k / 2, point p)
            const int si = lvl * 4 + p;
            gattn[qm * (L * 4) + si] = from_f32<AT>(sum[0]);
            gloc[(qm * (L * 4) + si) * 2] = from_f32<AT>(sum[1]);
            gloc[(qm * (L * 4) + si) * 2 + 1] = from_f32<AT>(sum[2]);
        }
    }
}

// Generic path: one warp per (n, q, m); lanes stride over channels; scalar atomics.
template <typename VT, typename AT>
__global__ void __launch_bounds__(128)
msda_bwd_generic_kernel(const VT* __restrict__ gout, const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ starts, const AT* __restrict__ locp, const AT* __restrict__ attnp,
                        float* __restrict__ gvalue, AT* __restrict__ gloc, AT* __restrict__ gattn, int64_t total_qm,
                        int S, int M, int D, int Lq, int L, int P) {
    const int lane = threadIdx.x & 31;
    const int64_t qm = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qm >= total_qm) return;
    const int m = static_cast<int>(qm % M);
    const int64_t n = (qm / M) / Lq;
    const int LP = L * P;
    constexpr int kChunks = 8;
    float g[kChunks];
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
        const int d = lane + ch * 32;
        g[ch] = d < D ? to_f32(gout[qm * D + d]) : 0.f;
    }
    for (int l = 0; l < L; ++l) {
        int H = static_cast<int>(__ldg(shapes + 2 * l)), W = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const long long start64 = __ldg(starts + l);
        if (start64 < 0 || H < 0 || W < 0 || start64 + static_cast<long long>(H) * W > S) H = W = 0;   // level outside S: skipped
        const int start = static_cast<int>(start64);
        for (int p = 0; p < P; ++p) {
            const int64_t si = qm * LP + l * P + p;
            const float locx = to_f32(locp[si * 2]), locy = to_f32(locp[si * 2 + 1]), a = to_f32(attnp[si]);
            float ga = 0.f, gx = 0.f, gy = 0.f;
            int x0, y0;
            float lx, ly;
            if (sample_coords(locx, locy, H, W, x0, y0, lx, ly)) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int xi = x0 + (c & 1), yi = y0 + (c >> 1);
                    if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
                    const float wx = (c & 1) ? lx : 1.f - lx, wy = (c >> 1) ? ly : 1.f - ly;
                    const int64_t row = ((n * S + start + yi * W + xi) * M + m) * D;
                    float dot = 0.f;
#pragma unroll
                    for (int ch = 0; ch < kChunks; ++ch) {
                        const int d = lane + ch * 32;
                        if (d < D) {
                            dot = fmaf(g[ch], to_f32(value[row + d]), dot);
                            atomicAdd(gvalue + row + d, a * wx * wy * g[ch]);
                        }
                    }
                    ga = fmaf(wx * wy, dot, ga);
                    gx += ((c & 1) ? wy : -wy) * dot;
                    gy += ((c >> 1) ? wx : -wx) * dot;
                }
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                ga += __shfl_xor_sync(kFullMask, ga, s);
                gx += __shfl_xor_sync(kFullMask, gx, s);
                gy += __shfl_xor_sync(kFullMask, gy, s);
            }
            if (lane == 0) {
                gattn[si] = from_f32<AT>(ga);
                gloc[si * 2] = from_f32<AT>(a * static_cast<float>(W) * gx);
                gloc[si * 2 + 1] = from_f32<AT>(a * static_cast<float>(H) * gy);
            }
        }
    }
}

template <typename VT, typename AT, bool FUSED>
cudaError_t launch_typed(const BwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    const int64_t total_qm = static_cast<int64_t>(d.N) * d.Lq * d.M;
    if (total_qm == 0) return cudaSuccess;
    const VT* gout = static_cast<const VT*>(a.grad_out);
    const VT* value = static_cast<const VT*>(a.value);
    const AT* loc = static_cast<const AT*>(a.loc);
    const AT* attn = static_cast<const AT*>(a.attn);
    AT* gloc = static_cast<AT*>(a.grad_loc);
    AT* gattn = static_cast<AT*>(a.grad_attn);
    if (d.D == 32 && d.P == 4 && d.L >= 1 && d.L <= 4) {
        int threads = tuning(kTuneBwdThreads, 256);
        threads = (threads / 32) * 32;
        if (threads < 32) threads = 32;
        if (threads > kBwdMaxThreads) threads = kBwdMaxThreads;
        int q_per_cta = tuning(kTuneBwdQpc, 0);
        if (q_per_cta <= 0)   // default: whole waves of the 4 x 148 resident 256-thread CTAs (an explicit BWD_QPC is taken as is)
            q_per_cta = balanced_q_per_cta(static_cast<int64_t>(d.N) * d.M, d.Lq, 128, (kBwdMaxThreads / threads) * 2 * 148, 1);
        while (q_per_cta > 1 && static_cast<int64_t>(d.N) * d.M * ((d.Lq + q_per_cta - 1) / q_per_cta) < 148 * 4) q_per_cta >>= 1;
        if (q_per_cta > d.Lq) q_per_cta = d.Lq;
        if (q_per_cta < 1) q_per_cta = 1;
        if (threads > q_per_cta * 32) threads = q_per_cta * 32;
        const int q_tiles = (d.Lq + q_per_cta - 1) / q_per_cta;
        const int64_t grid = static_cast<int64_t>(d.N) * d.M * q_tiles;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const dim3 gdim(static_cast<unsigned>(grid)), b(threads);
#define CAPE_BWD_CASE(LL)                                                                                          \
    case LL:                                                                                                       \
        if (d.M == 8)                                                                                              \
            msda_bwd_fast_kernel<VT, AT, LL, 8, FUSED><<<gdim, b, 0, stream>>>(                                     \
                gout, value, a.shapes, a.starts, loc, attn, a.ref_points, a.grad_value, gloc, gattn, d.N, d.S, d.M,  \
                d.Lq, q_per_cta, q_tiles);                                                                         \
        else                                                                                                       \
            msda_bwd_fast_kernel<VT, AT, LL, 0, FUSED><<<gdim, b, 0, stream>>>(                                     \
                gout, value, a.shapes, a.starts, loc, attn, a.ref_points, a.grad_value, gloc, gattn, d.N, d.S, d.M,  \
                d.Lq, q_per_cta, q_tiles);                                                                         \
        break;
        switch (d.L) {
            CAPE_BWD_CASE(1)
            CAPE_BWD_CASE(2)
            CAPE_BWD_CASE(3)
            CAPE_BWD_CASE(4)
        }
#undef CAPE_BWD_CASE
    } else {
        if (FUSED) return cudaErrorNotSupported;   // the fused prologue exists on the fast path only (ABI checks dims)
        const int warps = 4;
        const int64_t grid = (total_qm + warps - 1) / warps;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        msda_bwd_generic_kernel<VT, AT><<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
            gout, value, a.shapes, a.starts, loc, attn, a.grad_value, gloc, gattn, total_qm, d.S, d.M, d.D, d.Lq, d.L,
            d.P);
    }
    count_launch();
    return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_value_typed(const BwdArgs& a, cudaStream_t stream) {
    if (a.fused) return launch_typed<VT, float, true>(a, stream);
    if (a.aux_dtype == CAPE_DTYPE_F32) return launch_typed<VT, float, false>(a, stream);
    return launch_typed<VT, VT, false>(a, stream);
}

}  // namespace

cudaError_t launch_backward(const BwdArgs& a, cudaStream_t stream) {
    const int mode = tuning(kTuneBwdMode, 1);      // 1: L1 kernel + REDs; 2: staged rows + tensor-core scatter; 3: staged rows
    if (mode == 2 || mode == 3 || mode == 4) {   // 4: profiling only (results invalid), see msda_backward_staged.cu
        const cudaError_t e = launch_backward_staged(a, mode, stream);
        if (e == cudaSuccess) count_launch();
        if (e != cudaErrorNotSupported) return e;
    }
    if (mode == 5) {   // small CTAs + tensor-core scatter of the coarsest level (msda_backward_tc.cu)
        const cudaError_t e = launch_backward_tc(a, stream);
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
