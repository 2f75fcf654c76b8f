// Pieces of the MSDeformAttn forward shared by the sampling kernels (msda_forward.cu) and by the decode-step linear that
// samples its own input rows (decode_step.cu): per-level constants, the 4-corner bilinear gather of one sample per
// 8-lane group, and the one-warp-per-(n, q, m) sampler of the latency-oriented path.
#pragma once

#include "msda_common.cuh"

namespace cape {

// Per-level constants kept in registers for the whole CTA.
template <typename VT, int L>
struct FwdLevels {
    int H[L], W[L];
    const VT* img;       // row 0 of this (n, m, channel quad)
    int off[L];          // element offset of the level's first row (S*M*D < 2^31, checked by the ABI): 32-bit address math
    __device__ __forceinline__ void load(const int64_t* __restrict__ shapes, const int64_t* __restrict__ starts,
                                         const VT* vbase, int rowStride, int S) {
        img = vbase;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            H[l] = static_cast<int>(__ldg(shapes + 2 * l));
            W[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
            const long long s0 = __ldg(starts + l);
            // a level that does not fit inside S (the reference asserts sum(H*W) == S, deformable_transformer.py:94)
            // contributes nothing instead of being read out of bounds
            if (s0 < 0 || H[l] < 0 || W[l] < 0 || s0 + static_cast<long long>(H[l]) * W[l] > S) H[l] = W[l] = 0;
            off[l] = H[l] > 0 ? static_cast<int>(s0) * rowStride : 0;
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

// One sample per 8-lane group (4 per warp instruction), 4 corners each.  (px, py) are pixel coordinates from pixel_coord().
template <typename VT>
__device__ __forceinline__ void gather_level(const VT* __restrict__ img, int levelOff, int rowStride, int H, int W, float px,
                                             float py, float a, float4& acc) {
    const float xf = floorf(px), yf = floorf(py);
    const float lx = px - xf, ly = py - yf;
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const bool x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W);
    const bool x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W);
    const bool y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H);
    const bool y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H);
    const VT* p00 = img + (levelOff + (y0 * W + x0) * rowStride);
    const VT* p10 = p00 + W * rowStride;
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

// One warp samples ONE (n, q, m): lane = (point p = lane >> 3, channel quad k = lane & 7), the 4 levels unrolled so that all
// 16 corner loads of a lane are in flight at once; the 4 point groups are summed with 8 shuffles, after which EVERY lane
// holds channels 4k .. 4k+3 of the result.  FUSED: locp / attnp are the raw sampling offsets / attention logits (fp32) and
// refp the reference points — softmax and loc = ref + off / (W_l, H_l) happen here (deformable_transformer.py:100-105).
template <typename VT, typename LT, int L, bool FUSED>
__device__ __forceinline__ float4 sample_point(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                                               const int64_t* __restrict__ starts, const LT* __restrict__ locp,
                                               const LT* __restrict__ attnp, const float* __restrict__ refp, int64_t qm,
                                               int64_t n, int m, int64_t nq, int S, int M, int lane) {
    constexpr int D = 32;
    const int p = lane >> 3, k = lane & 7;
    const int rowStride = M * D;
    FwdLevels<VT, L> lv;
    lv.load(shapes, starts, value + (n * S * M + m) * D + k * 4, rowStride, S);
    const float dimf = lv.lane_dim(lane);
    float loc = 0.f, attn = FUSED ? -INFINITY : 0.f;
    if (lane < L * 8) loc = to_f32(locp[qm * (L * 8) + lane]);
    if (lane < L * 4) attn = to_f32(attnp[qm * (L * 4) + lane]);
    if (FUSED) {   // softmax over the 4L logits and loc = ref + off / (W_l, H_l)  (deformable_transformer.py:100-105)
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
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const float px = __shfl_sync(kFullMask, loc, l * 8 + p * 2);
        const float py = __shfl_sync(kFullMask, loc, l * 8 + p * 2 + 1);
        const float a = __shfl_sync(kFullMask, attn, l * 4 + p);
        gather_level(lv.img, lv.off[l], rowStride, lv.H[l], lv.W[l], px, py, a, acc);
    }
#pragma unroll
    for (int s = 8; s <= 16; s <<= 1) {
        acc.x += __shfl_xor_sync(kFullMask, acc.x, s);
        acc.y += __shfl_xor_sync(kFullMask, acc.y, s);
        acc.z += __shfl_xor_sync(kFullMask, acc.z, s);
        acc.w += __shfl_xor_sync(kFullMask, acc.w, s);
    }
    return acc;
}

}  // namespace cape
