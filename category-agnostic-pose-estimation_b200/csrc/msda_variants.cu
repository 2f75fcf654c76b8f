// Variant samplers of the same kernel family — non-default paths of the reference, small
// problem sizes, so these are straightforward fp32 kernels (one warp or one thread per output, scalar atomics) rather
// than tuned ones.
//
//  * query-pooled deformable sampling: TransformerDecoderLayerV4._sample_reference_points,
//    /root/reference/models/deformable_transformer_v2.py:661-687 — the sampling of ms_deform_attn_core_pytorch, but the
//    weighted sum runs over the QUERIES (the reference softmaxes the weights over dim 1) and the result keeps one row
//    per (level, point):  out[n, l*P+p, m*D+d] = sum_q A[n,q,m,l,p] * bilinear(V_l[n,:,m,d], loc[n,q,m,l,p]).
//  * planar point sampling: MSDeformablePoints.forward, /root/reference/models/deformable_points.py:118-128 —
//    F.grid_sample(bilinear, zeros padding, align_corners=True) of a level viewed channel-first, (B*G, c, H, W), at
//    positions given as (y, x) in [-1, 1]; the output is written directly in the (B, Hk*Wk, G*c) layout of :128.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kChunks = 8;   // channels handled per lane: D <= 256

// ---- query-pooled sampling ---------------------------------------------------------------------------------------
// One warp per (n, m, l, p); lanes stride over channels; loop over the queries.
__global__ void __launch_bounds__(128)
query_pool_fwd_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ starts, const float* __restrict__ loc, const float* __restrict__ attn,
                      float* __restrict__ out, int64_t total, int S, int M, int D, int Lq, int L, int P) {
    const int lane = threadIdx.x & 31;
    const int64_t w = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= total) return;
    const int LP = L * P;
    const int lp = static_cast<int>(w % LP);
    const int m = static_cast<int>((w / LP) % M);
    const int64_t n = w / (static_cast<int64_t>(LP) * M);
    const int l = lp / P;
    const int H = static_cast<int>(__ldg(shapes + 2 * l)), W = static_cast<int>(__ldg(shapes + 2 * l + 1));
    const int start = static_cast<int>(__ldg(starts + l));
    float acc[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) acc[c] = 0.f;
    for (int q = 0; q < Lq; ++q) {
        const int64_t si = ((n * Lq + q) * M + m) * LP + lp;
        const float a = __ldg(attn + si);
        int x0, y0;
        float lx, ly;
        if (!sample_coords(__ldg(loc + si * 2), __ldg(loc + si * 2 + 1), H, W, x0, y0, lx, ly)) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int xi = x0 + (c & 1), yi = y0 + (c >> 1);
            if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
            const float wgt = a * ((c >> 1) ? ly : 1.f - ly) * ((c & 1) ? lx : 1.f - lx);
            const float* v = value + ((n * S + start + yi * W + xi) * M + m) * D;
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const int d = lane + ch * 32;
                if (d < D) acc[ch] = fmaf(wgt, __ldg(v + d), acc[ch]);
            }
        }
    }
    float* o = out + ((n * LP + lp) * M + m) * D;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
        const int d = lane + ch * 32;
        if (d < D) o[d] = acc[ch];
    }
}

// One warp per (n, q, m): same per-sample gradient formulas as the main backward, with G = grad_out[n, l*P+p, m, :].
__global__ void __launch_bounds__(128)
query_pool_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ starts, const float* __restrict__ loc, const float* __restrict__ attn,
                      float* __restrict__ gvalue, float* __restrict__ gloc, float* __restrict__ gattn, int64_t total_qm,
                      int S, int M, int D, int Lq, int L, int P) {
    const int lane = threadIdx.x & 31;
    const int64_t qm = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qm >= total_qm) return;
    const int m = static_cast<int>(qm % M);
    const int64_t n = (qm / M) / Lq;
    const int LP = L * P;
    for (int l = 0; l < L; ++l) {
        const int H = static_cast<int>(__ldg(shapes + 2 * l)), W = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const int start = static_cast<int>(__ldg(starts + l));
        for (int p = 0; p < P; ++p) {
            const int64_t si = qm * LP + l * P + p;
            const float* g = gout + ((n * LP + l * P + p) * M + m) * D;
            const float a = __ldg(attn + si);
            float ga = 0.f, gx = 0.f, gy = 0.f;
            int x0, y0;
            float lx, ly;
            if (sample_coords(__ldg(loc + si * 2), __ldg(loc + si * 2 + 1), H, W, x0, y0, lx, ly)) {
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
                            const float gd = __ldg(g + d);
                            dot = fmaf(gd, __ldg(value + row + d), dot);
                            atomicAdd(gvalue + row + d, a * wx * wy * gd);
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
                gattn[si] = ga;
                gloc[si * 2] = a * static_cast<float>(W) * gx;
                gloc[si * 2 + 1] = a * static_cast<float>(H) * gy;
            }
        }
    }
}

// ---- planar point sampling (align_corners = True) ------------------------------------------------------------------
// Pixel coordinate for align_corners=True: ((g + 1) / 2) * (size - 1), same operation order as ATen.
__device__ __forceinline__ float unnormalize_ac(float g, int size) {
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), static_cast<float>(size - 1));
}

// One thread per (bg, hk, wk, ch); ch fastest so that the (B, Hk*Wk, G*c) output row is written contiguously.
__global__ void __launch_bounds__(256)
points_sample_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pos, float* __restrict__ out,
                         int64_t total, int G, int c, int H, int W, int Hk, int Wk) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int ch = static_cast<int>(t % c);
    const int64_t s = t / c;                       // (bg, hk, wk) flattened
    const int k = static_cast<int>(s % (Hk * Wk));
    const int64_t bg = s / (Hk * Wk);
    const float py = unnormalize_ac(__ldg(pos + s * 2), H);       // pos[..., 0] is y (the reference swaps to (x, y) at :126)
    const float px = unnormalize_ac(__ldg(pos + s * 2 + 1), W);
    const float xf = floorf(px), yf = floorf(py);
    const float lx = px - xf, ly = py - yf;
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const float* plane = x + (bg * c + ch) * static_cast<int64_t>(H) * W;
    float acc = 0.f;
#pragma unroll
    for (int cn = 0; cn < 4; ++cn) {
        const int xi = x0 + (cn & 1), yi = y0 + (cn >> 1);
        if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
        acc = fmaf(((cn >> 1) ? ly : 1.f - ly) * ((cn & 1) ? lx : 1.f - lx), __ldg(plane + yi * W + xi), acc);
    }
    const int64_t b = bg / G;
    const int g = static_cast<int>(bg % G);
    out[(b * (Hk * Wk) + k) * (static_cast<int64_t>(G) * c) + g * c + ch] = acc;
}

// One thread per (bg, hk, wk): loops over the c channels, scatters grad_x with scalar atomics, writes grad_pos (y, x).
__global__ void __launch_bounds__(256)
points_sample_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x, const float* __restrict__ pos,
                         float* __restrict__ gx_out, float* __restrict__ gpos, int64_t total, int G, int c, int H, int W,
                         int Hk, int Wk) {
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const int k = static_cast<int>(s % (Hk * Wk));
    const int64_t bg = s / (Hk * Wk);
    const int64_t b = bg / G;
    const int g = static_cast<int>(bg % G);
    const float py = unnormalize_ac(__ldg(pos + s * 2), H);
    const float px = unnormalize_ac(__ldg(pos + s * 2 + 1), W);
    const float xf = floorf(px), yf = floorf(py);
    const float lx = px - xf, ly = py - yf;
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const float* go = gout + (b * (Hk * Wk) + k) * (static_cast<int64_t>(G) * c) + g * c;
    float dpx = 0.f, dpy = 0.f;
    for (int ch = 0; ch < c; ++ch) {
        const int64_t plane = (bg * c + ch) * static_cast<int64_t>(H) * W;
        const float gd = __ldg(go + ch);
#pragma unroll
        for (int cn = 0; cn < 4; ++cn) {
            const int xi = x0 + (cn & 1), yi = y0 + (cn >> 1);
            if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
            const float wx = (cn & 1) ? lx : 1.f - lx, wy = (cn >> 1) ? ly : 1.f - ly;
            const float v = __ldg(x + plane + yi * W + xi);
            atomicAdd(gx_out + plane + yi * W + xi, wx * wy * gd);
            dpx += ((cn & 1) ? wy : -wy) * v * gd;
            dpy += ((cn >> 1) ? wx : -wx) * v * gd;
        }
    }
    // d pixel / d normalised = (size - 1) / 2
    gpos[s * 2] = dpy * 0.5f * static_cast<float>(H - 1);
    gpos[s * 2 + 1] = dpx * 0.5f * static_cast<float>(W - 1);
}

}  // namespace

cudaError_t launch_query_pool_forward(const FwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    const int64_t total = static_cast<int64_t>(d.N) * d.M * d.L * d.P;
    if (total == 0) return cudaSuccess;
    const int warps = 4;
    const int64_t grid = (total + warps - 1) / warps;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    query_pool_fwd_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
        static_cast<const float*>(a.value), a.shapes, a.starts, static_cast<const float*>(a.loc),
        static_cast<const float*>(a.attn), static_cast<float*>(a.out), total, d.S, d.M, d.D, d.Lq, d.L, d.P);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_query_pool_backward(const BwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    const int64_t total_qm = static_cast<int64_t>(d.N) * d.Lq * d.M;
    if (total_qm == 0) return cudaSuccess;
    const int warps = 4;
    const int64_t grid = (total_qm + warps - 1) / warps;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    query_pool_bwd_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
        static_cast<const float*>(a.grad_out), static_cast<const float*>(a.value), a.shapes, a.starts,
        static_cast<const float*>(a.loc), static_cast<const float*>(a.attn), a.grad_value,
        static_cast<float*>(a.grad_loc), static_cast<float*>(a.grad_attn), total_qm, d.S, d.M, d.D, d.Lq, d.L, d.P);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_points_sample_forward(const float* x, const float* pos, float* out, const PointsDims& p,
                                         cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(p.B) * p.G * p.Hk * p.Wk * p.c;
    if (total == 0) return cudaSuccess;
    const int64_t grid = (total + 255) / 256;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    points_sample_fwd_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(x, pos, out, total, p.G, p.c, p.H, p.W,
                                                                             p.Hk, p.Wk);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_points_sample_backward(const float* gout, const float* x, const float* pos, float* gx, float* gpos,
                                          const PointsDims& p, cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(p.B) * p.G * p.Hk * p.Wk;
    if (total == 0) return cudaSuccess;
    const int64_t grid = (total + 255) / 256;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    points_sample_bwd_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(gout, x, pos, gx, gpos, total, p.G, p.c,
                                                                             p.H, p.W, p.Hk, p.Wk);
    count_launch();
    return cudaGetLastError();
}

// ---- padding-mask fill of the projected value (MSDeformAttn.forward, /root/reference/models/deformable_transformer.py:96-97) ----
// value[r, :] = 0 where mask[r].  A CTA owns 256 consecutive rows: it reads their 256 mask bytes and leaves at once when none
// is set — for the all-False mask CAPE always passes the launch touches N*S bytes instead of reading and re-writing the
// whole value tensor, with no host-side `mask.any()` synchronisation.
namespace {
template <int BYTES_PER_ROW_UNIT>
__global__ void __launch_bounds__(256)
zero_masked_rows_kernel(uint4* __restrict__ value, const uint8_t* __restrict__ mask, long long rows, int units_per_row) {
    const long long r0 = static_cast<long long>(blockIdx.x) * 256;
    const long long r = r0 + threadIdx.x;
    const int mine = (r < rows && mask[r]) ? 1 : 0;
    if (!__syncthreads_or(mine)) return;
    __shared__ uint8_t flags[256];
    flags[threadIdx.x] = static_cast<uint8_t>(mine);
    __syncthreads();
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = 0; i < 256; ++i) {
        if (!flags[i]) continue;
        uint4* row = value + (r0 + i) * units_per_row;
        for (int u = threadIdx.x; u < units_per_row; u += 256) row[u] = z;
    }
}
}  // namespace

cudaError_t launch_zero_masked_rows(void* value, const uint8_t* mask, long long rows, int row_bytes, cudaStream_t stream) {
    if (rows == 0 || row_bytes == 0) return cudaSuccess;
    if (row_bytes % 16 != 0) return cudaErrorInvalidValue;
    const long long grid = (rows + 255) / 256;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    zero_masked_rows_kernel<16><<<static_cast<unsigned>(grid), 256, 0, stream>>>(static_cast<uint4*>(value), mask, rows,
                                                                                row_bytes / 16);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cape
