// Kernels of the incremental-decode step around the MSDeformAttn decode op: what
// TransformerDecoderLayer.forward (/root/reference/models/deformable_transformer_v2.py:320-370) and the per-layer part of
// TransformerDecoder.forward (:1069-1121) do for ONE new token per sequence.  At that size (B <= a few hundred rows of 256
// channels) library GEMM / attention kernels are pure launch latency (10 us SIMT sgemm + 3 us epilogue kernel, 42 us
// memory-efficient attention for a single query); these do the same fp32 arithmetic in a few microseconds each:
//
//  * decode_attn_kernel   — the KV-cache attention of the new token: appends the token's K/V rows to the cache at the
//                           device-resident position and attends positions 0..pos (self-attention, :322-341), or attends
//                           a fixed key set under a key-padding bias (support cross-attention, :350-357);
//  * skinny_linear_kernel — y = epilogue(x W^T + b) for a handful of rows: bias | bias+ReLU | bias(+residual)+LayerNorm,
//                           optional second addend on the input (tgt + query_pos) or the sine embedding of the reference
//                           points as the input (TransformerDecoder.get_query_pos_embed, :1005-1018);
//  * tiny_linear_kernel   — y = x W^T + b for N <= 8 outputs (class head, last coordinate layer) with the reference-point
//                           refinement sigmoid(y + inverse_sigmoid(ref)) (:1096-1102) as an optional epilogue.
//
// All fp32, FMA accumulation; weights are read from L2 (the whole 6-layer decoder is 32 MB).
#include <cfloat>

#include <cooperative_groups.h>

#include "msda_common.cuh"
#include "msda_launch.h"

namespace cg = cooperative_groups;

namespace cape {

namespace {

constexpr int kAttnWarps = 4;
constexpr int kAttnMaxKeys = 1024;

// One warp per (b, head); head dim 32.
__global__ void __launch_bounds__(kAttnWarps * 32)
decode_attn_kernel(const float* __restrict__ q, const float* __restrict__ k_new, const float* __restrict__ v_new,
                   float* __restrict__ k_cache, float* __restrict__ v_cache, const int64_t* __restrict__ pos_dev,
                   const float* __restrict__ key_bias, float* __restrict__ out, int B, int T, int H, int q_stride,
                   int new_stride) {
    constexpr int D = 32;
    __shared__ float probs[kAttnWarps][kAttnMaxKeys];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bh = blockIdx.x * kAttnWarps + warp;
    if (bh >= B * H) return;
    const int b = bh / H, h = bh % H;
    const int C = H * D;
    int pos = -1, n_keys = T;
    if (pos_dev) {
        const int64_t p = *pos_dev;
        if (p < 0 || p >= T) return;                       // out of range: leave `out` untouched
        pos = static_cast<int>(p);
        n_keys = pos + 1;
    }
    const float* qrow = q + static_cast<int64_t>(b) * q_stride + h * D;
    float4 qv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = __ldg(reinterpret_cast<const float4*>(qrow) + i);
    const float scale = rsqrtf(static_cast<float>(D));
    const float* kb = k_cache + static_cast<int64_t>(b) * T * C + h * D;
    const float* vb = v_cache + static_cast<int64_t>(b) * T * C + h * D;
    const float* knew = k_new ? k_new + static_cast<int64_t>(b) * new_stride + h * D : nullptr;
    const float* vnew = v_new ? v_new + static_cast<int64_t>(b) * new_stride + h * D : nullptr;
    // scores: lane = key
    float mx = -INFINITY;
    for (int t0 = 0; t0 < n_keys; t0 += 32) {
        const int t = t0 + lane;
        float s = -INFINITY;
        if (t < n_keys) {
            const float* krow = (t == pos && knew) ? knew : kb + static_cast<int64_t>(t) * C;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 kv = *reinterpret_cast<const float4*>(krow + 4 * i);
                acc = fmaf(qv[i].x, kv.x, acc);
                acc = fmaf(qv[i].y, kv.y, acc);
                acc = fmaf(qv[i].z, kv.z, acc);
                acc = fmaf(qv[i].w, kv.w, acc);
            }
            s = acc * scale;
            if (key_bias) s += key_bias[static_cast<int64_t>(b) * T + t];
            probs[warp][t] = s;
        }
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
    float sum = 0.f;
    for (int t = lane; t < n_keys; t += 32) {
        const float e = expf(probs[warp][t] - mx);
        probs[warp][t] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
    __syncwarp();
    // weighted sum of the values: lane = channel
    float acc0 = 0.f, acc1 = 0.f;
    int t = 0;
    for (; t + 1 < n_keys; t += 2) {
        const float v0 = (t == pos && vnew) ? vnew[lane] : vb[static_cast<int64_t>(t) * C + lane];
        const float v1 = (t + 1 == pos && vnew) ? vnew[lane] : vb[static_cast<int64_t>(t + 1) * C + lane];
        acc0 = fmaf(probs[warp][t], v0, acc0);
        acc1 = fmaf(probs[warp][t + 1], v1, acc1);
    }
    if (t < n_keys) {
        const float v0 = (t == pos && vnew) ? vnew[lane] : vb[static_cast<int64_t>(t) * C + lane];
        acc0 = fmaf(probs[warp][t], v0, acc0);
    }
    out[static_cast<int64_t>(b) * C + h * D + lane] = (acc0 + acc1) / sum;
    if (pos >= 0 && knew) {                                 // append the new token's rows to the cache
        k_cache[(static_cast<int64_t>(b) * T + pos) * C + h * D + lane] = knew[lane];
        v_cache[(static_cast<int64_t>(b) * T + pos) * C + h * D + lane] = vnew[lane];
    }
}

// ---- skinny linear ---------------------------------------------------------------------------------------------------
// Latency-oriented: a decode-step GEMM is 128 rows x 256 x 256 (17 MFLOP) — what matters is how many bytes one SM has to
// pull from L2 and how many dependent round trips the kernel makes.  Tile = 4 rows x 64 columns, so a 256 x 256 layer
// spreads over 32 x 4 = 128 CTAs and each SM reads a 64 KB weight slab instead of the whole 256 KB matrix.
// 256 threads = 16 k-groups x 16 column threads (4 columns each): a thread's share of the reduction dimension is K / 16
// rows of the transposed weight, fetched as up to 16 independent LDG.128 issued BEFORE the input rows are staged, so the
// weight round trip overlaps the input round trip.  The k-groups are summed through shared memory.  The LayerNorm
// epilogue needs whole rows: the (up to 4) CTAs that share a row tile form a thread-block cluster and exchange their
// per-row sums through distributed shared memory (two-pass mean / variance, like ATen).
constexpr int kSkRows = 4;        // rows per CTA
constexpr int kSkCols = 64;       // output columns per CTA
constexpr int kSkGroups = 16;     // k-groups
constexpr int kSkBatch = 16;      // weight rows a thread keeps in flight

template <int EPI>   // 0 bias, 1 bias + ReLU, 2 bias (+ residual) + LayerNorm over the N (<= 256) outputs
__global__ void __launch_bounds__(256)
skinny_linear_kernel(SkinnyArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                       // [kSkRows][K]
    float* part = smem + kSkRows * a.K;                     // [kSkGroups][kSkRows][kSkCols]
    __shared__ float stat[2][kSkRows];                      // this CTA's per-row partial sums (read by cluster peers)
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * kSkRows;
    const int c0 = blockIdx.y * kSkCols;
    const int rows = min(kSkRows, a.rows - r0);
    const int kg = tid >> 4, ct = tid & 15;
    const int col = c0 + ct * 4;
    const int kper = a.K / kSkGroups;                       // K % 16 == 0
    const bool col_live = col < a.N;
    const float* wp = a.wt + static_cast<int64_t>(kg * kper) * a.N + col;
    float acc[kSkRows][4];
#pragma unroll
    for (int r = 0; r < kSkRows; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    // first batch of weight rows: in flight while the inputs are staged
    float4 w[kSkBatch];
#pragma unroll
    for (int j = 0; j < kSkBatch; ++j)
        w[j] = (col_live && j < kper) ? __ldg(reinterpret_cast<const float4*>(wp + static_cast<int64_t>(j) * a.N))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    // stage the input rows (optionally x + x2, or the sine embedding of the reference points)
    for (int i = tid; i < kSkRows * a.K; i += 256) {
        const int r = i / a.K, k = i - r * a.K;
        float v = 0.f;
        if (r < rows) {
            if (a.sine_dim_t) {   // K = 256: [coordinate 0: sin, cos interleaved over 128 | coordinate 1: same]  (:1013-1017)
                const int half = k >> 7, j = k & 127;
                const float coord = a.x[static_cast<int64_t>(r0 + r) * a.x_stride + half];
                const float arg = (coord * 6.283185307179586f) / a.sine_dim_t[j];
                v = (j & 1) ? cosf(arg) : sinf(arg);
            } else {
                v = a.x[static_cast<int64_t>(r0 + r) * a.x_stride + k];
                if (a.x2) v += a.x2[static_cast<int64_t>(r0 + r) * a.x2_stride + k];
            }
        }
        xs[i] = v;
    }
    __syncthreads();
    const float* xp = xs + kg * kper;
    for (int k0 = 0; k0 < kper; k0 += kSkBatch) {
#pragma unroll
        for (int j = 0; j < kSkBatch; ++j) {
            if (k0 + j < kper) {
#pragma unroll
                for (int r = 0; r < kSkRows; ++r) {
                    const float xk = xp[r * a.K + k0 + j];
                    acc[r][0] = fmaf(xk, w[j].x, acc[r][0]);
                    acc[r][1] = fmaf(xk, w[j].y, acc[r][1]);
                    acc[r][2] = fmaf(xk, w[j].z, acc[r][2]);
                    acc[r][3] = fmaf(xk, w[j].w, acc[r][3]);
                }
            }
        }
        if (k0 + kSkBatch < kper) {
#pragma unroll
            for (int j = 0; j < kSkBatch; ++j)
                w[j] = (col_live && k0 + kSkBatch + j < kper)
                           ? __ldg(reinterpret_cast<const float4*>(wp + static_cast<int64_t>(k0 + kSkBatch + j) * a.N))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int r = 0; r < kSkRows; ++r)
        *reinterpret_cast<float4*>(part + (kg * kSkRows + r) * kSkCols + ct * 4) =
            make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    __syncthreads();
    // thread t finishes (row t / 64, column c0 + t % 64)
    const int r = tid >> 6, cc = tid & 63;
    const int c = c0 + cc;
    const bool live = c < a.N && r < rows;
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < kSkGroups; ++g) v += part[(g * kSkRows + r) * kSkCols + cc];
    if (live && a.bias) v += __ldg(a.bias + c);
    if (EPI == 1) v = fmaxf(v, 0.f);
    if (EPI == 2) {
        if (live && a.res) v += a.res[static_cast<int64_t>(r0 + r) * a.res_stride + c];
        if (!live) v = 0.f;
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned peers = cluster.num_blocks();
        const int lane = tid & 31, half = (tid >> 5) & 1;            // a row is two warps
        __shared__ float warp_part[kSkRows][2];
        float mean = 0.f, rstd = 0.f;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            float s = pass == 0 ? v : ((c < a.N) ? (v - mean) * (v - mean) : 0.f);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
            if (lane == 0) warp_part[r][half] = s;
            __syncthreads();
            if (tid < kSkRows) stat[pass][tid] = warp_part[tid][0] + warp_part[tid][1];
            cluster.sync();                                          // every CTA of the row tile has published its partial
            float total = 0.f;
            for (unsigned p = 0; p < peers; ++p) total += cluster.map_shared_rank(&stat[pass][0], p)[r];
            if (pass == 0) mean = total / static_cast<float>(a.N);
            else rstd = rsqrtf(total / static_cast<float>(a.N) + a.eps);
        }
        if (live) v = (v - mean) * rstd * __ldg(a.gamma + c) + __ldg(a.beta + c);
        cluster.sync();                                              // nobody leaves while a peer may still read its stats
    }
    if (live) a.y[static_cast<int64_t>(r0 + r) * a.y_stride + c] = v;
}

// One warp per row; N <= 8 outputs; w is (N, K) row-major (a Linear's own layout).
__global__ void __launch_bounds__(128)
tiny_linear_kernel(const float* __restrict__ x, int x_stride, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ refine_ref, float* __restrict__ y, int rows, int K, int N) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * x_stride + k);
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (n < N) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(n) * K + k));
                acc[n] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[n]))));
            }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[n] += __shfl_xor_sync(kFullMask, acc[n], o);
    if (lane < N) {
        float v = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (lane == n) v = acc[n];
        if (bias) v += __ldg(bias + lane);
        if (refine_ref) {   // sigmoid(v + inverse_sigmoid(ref)), util/misc.py:436-440 with eps = 1e-5
            const float p = fminf(fmaxf(refine_ref[static_cast<int64_t>(r) * N + lane], 0.f), 1.f);
            const float inv = logf(fmaxf(p, 1e-5f) / fmaxf(1.f - p, 1e-5f));
            v = 1.f / (1.f + expf(-(v + inv)));
        }
        y[static_cast<int64_t>(r) * N + lane] = v;
    }
}

}  // namespace

cudaError_t launch_decode_attention(const float* q, const float* k_new, const float* v_new, float* k_cache, float* v_cache,
                                    const int64_t* pos_dev, const float* key_bias, float* out, int B, int T, int H,
                                    int q_stride, int new_stride, cudaStream_t stream) {
    const int warps = B * H;
    if (warps == 0) return cudaSuccess;
    decode_attn_kernel<<<(warps + kAttnWarps - 1) / kAttnWarps, kAttnWarps * 32, 0, stream>>>(
        q, k_new, v_new, k_cache, v_cache, pos_dev, key_bias, out, B, T, H, q_stride, new_stride);
    count_launch();
    return cudaGetLastError();
}

size_t skinny_smem_bytes(int K) { return (static_cast<size_t>(kSkRows) * K + kSkGroups * kSkRows * kSkCols) * sizeof(float); }

cudaError_t launch_skinny_linear(const SkinnyArgs& a, int epilogue, cudaStream_t stream) {
    if (a.rows == 0) return cudaSuccess;
    const unsigned col_tiles = (a.N + kSkCols - 1) / kSkCols;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.rows + kSkRows - 1) / kSkRows, col_tiles);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = skinny_smem_bytes(a.K);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = epilogue == 2 ? col_tiles : 1;    // LayerNorm: the CTAs of a row tile share statistics
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    switch (epilogue) {
        case 0: e = cudaLaunchKernelEx(&cfg, skinny_linear_kernel<0>, a); break;
        case 1: e = cudaLaunchKernelEx(&cfg, skinny_linear_kernel<1>, a); break;
        case 2: e = cudaLaunchKernelEx(&cfg, skinny_linear_kernel<2>, a); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_tiny_linear(const float* x, int x_stride, const float* w, const float* bias, const float* refine_ref,
                               float* y, int rows, int K, int N, cudaStream_t stream) {
    if (rows == 0) return cudaSuccess;
    tiny_linear_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(x, x_stride, w, bias, refine_ref, y, rows, K, N);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cape
