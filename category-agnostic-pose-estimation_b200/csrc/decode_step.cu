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
#include "msda_point.cuh"

namespace cg = cooperative_groups;

namespace cape {

namespace {

// Launch configuration of a step kernel with an optional cluster shape.
// (Programmatic dependent launch between the step kernels was measured — weight prefetch ahead of griddepcontrol.wait —
// and changed nothing inside the replayed graph: 559 vs 557 us per token.  Not kept.)
struct StepLaunch {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    StepLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_y = 1) {
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        unsigned n = 0;
        if (cluster_y > 1) {
            attr[n].id = cudaLaunchAttributeClusterDimension;
            attr[n].val.clusterDim.x = 1;
            attr[n].val.clusterDim.y = cluster_y;
            attr[n].val.clusterDim.z = 1;
            ++n;
        }
        cfg.attrs = attr;
        cfg.numAttrs = n;
    }
};

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
    // scores: lane = key.  Cached rows first (plain strided loads), the new token's own key last (index n_cached).
    const int n_cached = knew ? pos : n_keys;
    float mx = -INFINITY;
    for (int t0 = 0; t0 < n_cached; t0 += 32) {
        const int t = t0 + lane;
        float s = -INFINITY;
        if (t < n_cached) {
            const float4* krow = reinterpret_cast<const float4*>(kb + static_cast<int64_t>(t) * C);
            float4 kv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) kv[i] = krow[i];
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                a0 = fmaf(qv[i].x, kv[i].x, fmaf(qv[i].y, kv[i].y, fmaf(qv[i].z, kv[i].z, fmaf(qv[i].w, kv[i].w, a0))));
                a1 = fmaf(qv[i + 1].x, kv[i + 1].x,
                          fmaf(qv[i + 1].y, kv[i + 1].y, fmaf(qv[i + 1].z, kv[i + 1].z, fmaf(qv[i + 1].w, kv[i + 1].w, a1))));
            }
            s = (a0 + a1) * scale;
            if (key_bias) s += key_bias[static_cast<int64_t>(b) * T + t];
            probs[warp][t] = s;
        }
        mx = fmaxf(mx, s);
    }
    float s_new = -INFINITY;
    if (knew) {   // every lane computes the new token's score (same value), lane 0 records it
        float a0 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 kv = *reinterpret_cast<const float4*>(knew + 4 * i);
            a0 = fmaf(qv[i].x, kv.x, fmaf(qv[i].y, kv.y, fmaf(qv[i].z, kv.z, fmaf(qv[i].w, kv.w, a0))));
        }
        s_new = a0 * scale;
        mx = fmaxf(mx, s_new);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
    float sum = 0.f;
    for (int t = lane; t < n_cached; t += 32) {
        const float e = expf(probs[warp][t] - mx);
        probs[warp][t] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
    const float p_new = knew ? expf(s_new - mx) : 0.f;
    sum += p_new;
    __syncwarp();
    // weighted sum of the values: lane = channel; 8 cached rows in flight per step
    float acc0 = 0.f, acc1 = 0.f;
    int t0 = 0;
    for (; t0 + 8 <= n_cached; t0 += 8) {
        float vv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) vv[j] = vb[static_cast<int64_t>(t0 + j) * C + lane];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            acc0 = fmaf(probs[warp][t0 + j], vv[j], acc0);
            acc1 = fmaf(probs[warp][t0 + j + 1], vv[j + 1], acc1);
        }
    }
    for (; t0 < n_cached; ++t0) acc0 = fmaf(probs[warp][t0], vb[static_cast<int64_t>(t0) * C + lane], acc0);
    if (knew) acc1 = fmaf(p_new, vnew[lane], acc1);
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
constexpr int kSkBatch = 16;      // weight rows a thread keeps in flight

// EPI: 0 bias, 1 bias + ReLU, 2 bias (+ residual) + LayerNorm over the N (<= 256) outputs,
//      3 bias + ReLU followed by the coordinate head's last layer and the reference-point refinement
//        ref' = sigmoid(relu(x W^T + b) . W3^T + b3 + inverse_sigmoid(ref))     (deformable_transformer_v2.py:1096-1102)
//        ref_levels[row, l, :] = ref' * valid_ratios[row, l, :]                  (:1072, the next layer's sampling centres)
//      — the hidden row never leaves the chip; the 2 dot products are reduced over the cluster like the LayerNorm sums.
// GROUPS: k-groups (16 threads each): 16 for K = 256, 32 for long reductions (K = 1024: two weight batches instead of four).
template <int EPI, int GROUPS>
__global__ void __launch_bounds__(GROUPS * 16)
skinny_linear_kernel(SkinnyArgs a) {
    constexpr int kThreads = GROUPS * 16;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                       // [kSkRows][K]
    float* part = smem + kSkRows * a.K;                     // [GROUPS][kSkRows][kSkCols]
    __shared__ float stat[3][kSkRows];                      // this CTA's per-row statistics (read by cluster peers)
    __shared__ float warp_part[2][kSkRows][2];
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * kSkRows;
    const int c0 = blockIdx.y * kSkCols;
    const int rows = min(kSkRows, a.rows - r0);
    const int kg = tid >> 4, ct = tid & 15;
    const int col = c0 + ct * 4;
    const int kper = a.K / GROUPS;                          // multiple of 4 (checked by the launcher)
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
    // stage the input rows (optionally x + x2, the sine embedding of the reference points, or — output_proj of the decode
    // step's MSDeformAttn, deformable_transformer.py:112-113 — the sampled rows themselves)
    if (EPI == 2 && a.msda_value) {
        // The CTAs of a row tile (one cluster) share the sampling: CTA `rank` takes every peers-th (row, head) pair, one warp
        // per pair with all 16 corner loads in flight (sample_point), and writes the 32 channels into the input tile of
        // EVERY CTA of the cluster through distributed shared memory; same arithmetic as cape_msda_decode.
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned peers = cluster.num_blocks(), rank = cluster.block_rank();
        const int lane = tid & 31, warp = tid >> 5;
        const int M = a.msda_M;
        cluster.sync();                                     // every peer's shared memory exists before anybody writes into it
        for (int i = static_cast<int>(rank) + static_cast<int>(peers) * warp; i < kSkRows * M; i += static_cast<int>(peers) * (kThreads >> 5)) {
            const int r = i / M, m = i - r * M;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
                const int64_t nq = r0 + r, qm = nq * M + m;
                acc = sample_point<float, float, 4, true>(a.msda_value, a.msda_shapes, a.msda_starts, a.msda_off, a.msda_logits,
                                                          a.msda_ref, qm, nq / a.msda_Lq, m, nq, a.msda_S, M, lane);
            }
            if (lane < 8)
                for (unsigned p = 0; p < peers; ++p)
                    *reinterpret_cast<float4*>(cluster.map_shared_rank(xs, p) + r * a.K + m * 32 + lane * 4) = acc;
        }
        cluster.sync();                                     // all peers' channels have landed
    } else if (a.sine_dim_t) {   // K = 256: [coordinate 0: sin, cos interleaved over 128 | coordinate 1: same]  (:1013-1017)
        for (int i = tid; i < kSkRows * 256; i += kThreads) {
            const int r = i >> 8, k = i & 255;
            float v = 0.f;
            if (r < rows) {
                const int half = k >> 7, j = k & 127;
                const float coord = a.x[static_cast<int64_t>(r0 + r) * a.x_stride + half];
                const float arg = (coord * 6.283185307179586f) / __ldg(a.sine_dim_t + j);
                v = (j & 1) ? cosf(arg) : sinf(arg);
            }
            xs[i] = v;
        }
    } else {              // 16-byte vectors: x / x2 rows are 16-byte aligned (checked by the ABI)
        const int k4 = a.K >> 2;
        for (int i = tid; i < kSkRows * k4; i += kThreads) {
            const int r = i / k4, k = (i - r * k4) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
                v = *reinterpret_cast<const float4*>(a.x + static_cast<int64_t>(r0 + r) * a.x_stride + k);
                if (a.x2) {
                    const float4 u = *reinterpret_cast<const float4*>(a.x2 + static_cast<int64_t>(r0 + r) * a.x2_stride + k);
                    v.x += u.x, v.y += u.y, v.z += u.z, v.w += u.w;
                }
            }
            *reinterpret_cast<float4*>(xs + r * a.K + k) = v;
        }
    }
    __syncthreads();
    const float* xp = xs + kg * kper;
    for (int k0 = 0; k0 < kper; k0 += kSkBatch) {
#pragma unroll
        for (int j = 0; j < kSkBatch; j += 4) {
            if (k0 + j < kper) {
#pragma unroll
                for (int r = 0; r < kSkRows; ++r) {
                    const float4 xv = *reinterpret_cast<const float4*>(xp + r * a.K + k0 + j);
                    const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        acc[r][0] = fmaf(xk[jj], w[j + jj].x, acc[r][0]);
                        acc[r][1] = fmaf(xk[jj], w[j + jj].y, acc[r][1]);
                        acc[r][2] = fmaf(xk[jj], w[j + jj].z, acc[r][2]);
                        acc[r][3] = fmaf(xk[jj], w[j + jj].w, acc[r][3]);
                    }
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
    // the first 256 threads finish one output each: (row t / 64, column c0 + t % 64); the others only keep the barriers
    const bool fin = tid < kSkRows * kSkCols;
    const int r = (tid >> 6) & (kSkRows - 1), cc = tid & 63;
    const int c = c0 + cc;
    const bool live = fin && c < a.N && r < rows;
    float v = 0.f;
    if (fin) {
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) v += part[(g * kSkRows + r) * kSkCols + cc];
    }
    if (live && a.bias) v += __ldg(a.bias + c);
    if (EPI == 1 || EPI == 3) v = fmaxf(v, 0.f);
    const int lane = tid & 31, half = (tid >> 5) & 1;                  // a row is two warps
    if (EPI == 2) {
        // LayerNorm over the row, which is spread over the cluster: every CTA publishes (mean, M2) of its own columns and
        // the parts are combined with the parallel-variance formula — one exchange, no cancellation
        if (live && a.res) v += a.res[static_cast<int64_t>(r0 + r) * a.res_stride + c];
        const bool in_row = fin && c < a.N;
        if (!in_row) v = 0.f;
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned peers = cluster.num_blocks();
        const float n_local = static_cast<float>(max(0, min(kSkCols, a.N - c0)));
        float s = v;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
        if (fin && lane == 0) warp_part[0][r][half] = s;
        __syncthreads();
        const float mean_local = (warp_part[0][r][0] + warp_part[0][r][1]) / fmaxf(n_local, 1.f);
        float d = in_row ? (v - mean_local) * (v - mean_local) : 0.f;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(kFullMask, d, o);
        if (fin && lane == 0) warp_part[1][r][half] = d;
        __syncthreads();
        if (tid < kSkRows) {
            stat[0][tid] = (warp_part[0][tid][0] + warp_part[0][tid][1]) / fmaxf(n_local, 1.f);
            stat[1][tid] = warp_part[1][tid][0] + warp_part[1][tid][1];
            stat[2][tid] = n_local;
        }
        cluster.sync();                                                // every CTA of the row tile has published
        float mean = 0.f;
        for (unsigned p = 0; p < peers; ++p) {
            const float* st = cluster.map_shared_rank(&stat[0][0], p);
            mean += st[2 * kSkRows + r] * st[r];
        }
        mean /= static_cast<float>(a.N);
        float m2 = 0.f;
        for (unsigned p = 0; p < peers; ++p) {
            const float* st = cluster.map_shared_rank(&stat[0][0], p);
            const float dm = st[r] - mean;
            m2 += st[kSkRows + r] + st[2 * kSkRows + r] * dm * dm;
        }
        const float rstd = rsqrtf(m2 / static_cast<float>(a.N) + a.eps);
        if (live) v = (v - mean) * rstd * __ldg(a.gamma + c) + __ldg(a.beta + c);
        cluster.sync();                                                // nobody leaves while a peer may still read its stats
    }
    if (EPI == 3) {
        if (!live) v = 0.f;
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned peers = cluster.num_blocks();
        float s0 = live ? v * __ldg(a.w3 + c) : 0.f;                       // W3 is (2, N) row-major
        float s1 = live ? v * __ldg(a.w3 + a.N + c) : 0.f;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            s0 += __shfl_xor_sync(kFullMask, s0, o);
            s1 += __shfl_xor_sync(kFullMask, s1, o);
        }
        if (fin && lane == 0) {
            warp_part[0][r][half] = s0;
            warp_part[1][r][half] = s1;
        }
        __syncthreads();
        if (tid < 2 * kSkRows) stat[tid / kSkRows][tid % kSkRows] = warp_part[tid / kSkRows][tid % kSkRows][0] +
                                                                    warp_part[tid / kSkRows][tid % kSkRows][1];
        cluster.sync();
        if (cluster.block_rank() == 0 && tid < 2 * kSkRows) {
            const int o = tid / kSkRows, rr = tid % kSkRows;              // output coordinate o of row rr
            if (rr < rows) {
                float total = __ldg(a.b3 + o);
                for (unsigned p = 0; p < peers; ++p) total += cluster.map_shared_rank(&stat[o][0], p)[rr];
                const float pr = fminf(fmaxf(a.ref_in[static_cast<int64_t>(r0 + rr) * 2 + o], 0.f), 1.f);
                const float inv = logf(fmaxf(pr, 1e-5f) / fmaxf(1.f - pr, 1e-5f));      // util/misc.py:436-440
                const float nr = 1.f / (1.f + expf(-(total + inv)));
                a.ref_out[static_cast<int64_t>(r0 + rr) * 2 + o] = nr;
                for (int l = 0; l < a.n_levels; ++l)
                    a.ref_levels[(static_cast<int64_t>(r0 + rr) * a.n_levels + l) * 2 + o] =
                        nr * a.valid_ratios[(static_cast<int64_t>(r0 + rr) * a.n_levels + l) * 2 + o];
            }
        }
        cluster.sync();
        return;                                                            // the hidden row is not written
    }
    if (live) {
        if (a.y2 && c >= a.split) a.y2[static_cast<int64_t>(r0 + r) * a.y2_stride + (c - a.split)] = v;
        else a.y[static_cast<int64_t>(r0 + r) * a.y_stride + c] = v;
    }
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
    StepLaunch l(dim3((warps + kAttnWarps - 1) / kAttnWarps), dim3(kAttnWarps * 32), 0, stream);
    const cudaError_t e = cudaLaunchKernelEx(&l.cfg, decode_attn_kernel, q, k_new, v_new, k_cache, v_cache, pos_dev, key_bias,
                                             out, B, T, H, q_stride, new_stride);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int GROUPS>
cudaError_t launch_skinny_groups(const SkinnyArgs& a, int epilogue, cudaStream_t stream) {
    const unsigned col_tiles = (a.N + kSkCols - 1) / kSkCols;
    // LayerNorm / head epilogues: the CTAs of a row tile form a cluster and share their sums
    StepLaunch l(dim3((a.rows + kSkRows - 1) / kSkRows, col_tiles), dim3(GROUPS * 16),
                 (static_cast<size_t>(kSkRows) * a.K + static_cast<size_t>(GROUPS) * kSkRows * kSkCols) * sizeof(float), stream,
                 epilogue >= 2 ? col_tiles : 1);
    cudaLaunchConfig_t& cfg = l.cfg;
    static unsigned long long configured = 0;   // per GROUPS instantiation and device: opt in to > 48 KB of shared memory
    if (first_use_on_device(&configured)) {
        cudaFuncSetAttribute(skinny_linear_kernel<0, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(skinny_linear_kernel<1, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(skinny_linear_kernel<2, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(skinny_linear_kernel<3, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }
    switch (epilogue) {
        case 0: return cudaLaunchKernelEx(&cfg, skinny_linear_kernel<0, GROUPS>, a);
        case 1: return cudaLaunchKernelEx(&cfg, skinny_linear_kernel<1, GROUPS>, a);
        case 2: return cudaLaunchKernelEx(&cfg, skinny_linear_kernel<2, GROUPS>, a);
        case 3: return cudaLaunchKernelEx(&cfg, skinny_linear_kernel<3, GROUPS>, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_skinny_linear(const SkinnyArgs& a, int epilogue, cudaStream_t stream) {
    if (a.rows == 0) return cudaSuccess;
    // 32 k-groups (512 threads) for long reductions; needs K % 128 == 0 so that a group's share stays a multiple of 4
    const cudaError_t e = (a.K >= 512 && a.K % 128 == 0 && a.K <= 1024) ? launch_skinny_groups<32>(a, epilogue, stream)
                                                                       : launch_skinny_groups<16>(a, epilogue, stream);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_tiny_linear(const float* x, int x_stride, const float* w, const float* bias, const float* refine_ref,
                               float* y, int rows, int K, int N, cudaStream_t stream) {
    if (rows == 0) return cudaSuccess;
    StepLaunch l(dim3((rows + 3) / 4), dim3(128), 0, stream);
    const cudaError_t e = cudaLaunchKernelEx(&l.cfg, tiny_linear_kernel, x, x_stride, w, bias, refine_ref, y, rows, K, N);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace cape
