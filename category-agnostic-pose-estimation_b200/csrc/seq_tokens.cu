// Sequence side of the decoder: the two data formats either side of
// the MSDeformAttn decode step, kept on the device so a generated token never visits the host.
//
//  * bilinear token embedding — TransformerDecoder._seq_embed, /root/reference/models/deformable_transformer_v2.py:984-997:
//        out = e11*dx2*dy2 + e21*dx1*dy2 + e12*dx2*dy1 + e22*dx1*dy1,   e_ab = token_embed(seq_ab)
//    evaluated in the reference's order (left-to-right products, left-to-right sums, no FMA contraction) so the fp32
//    result is the one the eager expression gives.  Backward scatters into the embedding table's gradient and leaves
//    the padding row untouched, as nn.Embedding(padding_idx=...) does.
//  * token bookkeeping of the autoregressive loop — RoomFormerV2.forward_inference, /root/reference/models/roomformer_v2.py:
//    :548-597: argmax over the class logits, coordinate -> 4 neighbouring bin tokens + bilinear deltas, <sep>/<cls>/<eos>/
//    <pad> handling, the per-sample "unfinished" flag; all in fp32 exactly as the numpy scalars of the reference behave.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

// One warp per token; lanes stride over channel quads (C % 4 == 0).
__global__ void __launch_bounds__(128)
seq_embed_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ s11, const int64_t* __restrict__ s12,
                     const int64_t* __restrict__ s21, const int64_t* __restrict__ s22, const float* __restrict__ dx1,
                     const float* __restrict__ dx2, const float* __restrict__ dy1, const float* __restrict__ dy2,
                     float* __restrict__ out, int64_t tokens, int C, int V) {
    const int lane = threadIdx.x & 31;
    const int64_t t = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= tokens) return;
    const int64_t i11 = __ldg(s11 + t), i12 = __ldg(s12 + t), i21 = __ldg(s21 + t), i22 = __ldg(s22 + t);
    const float x1 = __ldg(dx1 + t), x2 = __ldg(dx2 + t), y1 = __ldg(dy1 + t), y2 = __ldg(dy2 + t);
    const bool ok = static_cast<uint64_t>(i11) < static_cast<uint64_t>(V) && static_cast<uint64_t>(i12) < static_cast<uint64_t>(V) &&
                    static_cast<uint64_t>(i21) < static_cast<uint64_t>(V) && static_cast<uint64_t>(i22) < static_cast<uint64_t>(V);
    for (int c = lane * 4; c < C; c += 128) {
        float4 r;
        if (ok) {
            const float4 e11 = ld4(table + i11 * C + c), e21 = ld4(table + i21 * C + c);
            const float4 e12 = ld4(table + i12 * C + c), e22 = ld4(table + i22 * C + c);
#define CAPE_SEQ_TERM(f)                                                                                                 \
    __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(e11.f, x2), y2), __fmul_rn(__fmul_rn(e21.f, x1), y2)),             \
                        __fmul_rn(__fmul_rn(e12.f, x2), y1)),                                                            \
              __fmul_rn(__fmul_rn(e22.f, x1), y1))
            r = make_float4(CAPE_SEQ_TERM(x), CAPE_SEQ_TERM(y), CAPE_SEQ_TERM(z), CAPE_SEQ_TERM(w));
#undef CAPE_SEQ_TERM
        } else {   // token id outside the table: poison the row instead of reading out of bounds
            const float nan = __int_as_float(0x7fc00000);
            r = make_float4(nan, nan, nan, nan);
        }
        st4(out + t * C + c, r);
    }
}

__global__ void __launch_bounds__(128)
seq_embed_bwd_kernel(const float* __restrict__ gout, const int64_t* __restrict__ s11, const int64_t* __restrict__ s12,
                     const int64_t* __restrict__ s21, const int64_t* __restrict__ s22, const float* __restrict__ dx1,
                     const float* __restrict__ dx2, const float* __restrict__ dy1, const float* __restrict__ dy2,
                     float* __restrict__ gtable, int64_t tokens, int C, int V, int64_t padding_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t t = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= tokens) return;
    const int64_t idx[4] = {__ldg(s11 + t), __ldg(s21 + t), __ldg(s12 + t), __ldg(s22 + t)};
    const float x1 = __ldg(dx1 + t), x2 = __ldg(dx2 + t), y1 = __ldg(dy1 + t), y2 = __ldg(dy2 + t);
    const float w[4] = {x2 * y2, x1 * y2, x2 * y1, x1 * y1};
    for (int c = lane * 4; c < C; c += 128) {
        const float4 g = ld4(gout + t * C + c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool live = static_cast<uint64_t>(idx[k]) < static_cast<uint64_t>(V) && idx[k] != padding_idx;
            red_add4_if(gtable + idx[k] * C + c, live, w[k] * g.x, w[k] * g.y, w[k] * g.z, w[k] * g.w);
        }
    }
}

// One thread per sample.  Mirrors the body of the reference's `for j in range(bs)` loop (roomformer_v2.py:548-597).
__global__ void __launch_bounds__(128)
token_step_kernel(const float* __restrict__ cls_logits, const float* __restrict__ reg, const int64_t* __restrict__ step_dev,
                  cape_token_state st, cape_tokenizer tk, int B, int n_classes) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    const int64_t i = *step_dev;
    if (i < 0 || i >= st.max_len) return;
    // record this step's head outputs (the reference appends them to output_cls_list / output_reg_list, :523-535)
    int best = 0;
    float best_v = cls_logits[j * n_classes];
    st.pred_logits[(static_cast<int64_t>(j) * st.max_len + i) * n_classes] = best_v;
    for (int c = 1; c < n_classes; ++c) {
        const float v = cls_logits[j * n_classes + c];
        st.pred_logits[(static_cast<int64_t>(j) * st.max_len + i) * n_classes + c] = v;
        if (v > best_v) {      // first maximum wins, like torch.argmax
            best_v = v;
            best = c;
        }
    }
    const float rx = reg[j * 2], ry = reg[j * 2 + 1];
    st.pred_coords[(static_cast<int64_t>(j) * st.max_len + i) * 2] = rx;
    st.pred_coords[(static_cast<int64_t>(j) * st.max_len + i) * 2 + 1] = ry;

    int64_t t11, t12, t21, t22;
    float dx = 0.f, dy = 0.f, gx = 0.f, gy = 0.f;
    int kind = -1;
    if (st.unfinished[j]) {
        if (best == tk.type_coord || (best == tk.type_eos && i < tk.min_len)) {
            const float x = fminf(rx, 1.f), y = fminf(ry, 1.f);                   // :552-553
            gx = x;
            gy = y;
            kind = 0;
            const float xs = __fmul_rn(x, static_cast<float>(tk.num_bins - 1));  // :557-558 (np.float32 * int)
            const float ys = __fmul_rn(y, static_cast<float>(tk.num_bins - 1));
            const float xf = floorf(xs), yf = floorf(ys), xc = ceilf(xs), yc = ceilf(ys);
            t11 = static_cast<int64_t>(xf) * tk.num_bins + static_cast<int64_t>(yf);   // :566-569
            t12 = static_cast<int64_t>(xf) * tk.num_bins + static_cast<int64_t>(yc);
            t21 = static_cast<int64_t>(xc) * tk.num_bins + static_cast<int64_t>(yf);
            t22 = static_cast<int64_t>(xc) * tk.num_bins + static_cast<int64_t>(yc);
            dx = __fsub_rn(xs, xf);
            dy = __fsub_rn(ys, yf);
        } else if (best == tk.type_sep) {
            kind = 2;
            t11 = t12 = t21 = t22 = tk.sep;
        } else if (best == tk.type_cls) {
            t11 = t12 = t21 = t22 = tk.cls;
        } else {                                                                  // <eos> at i >= min_len: done
            st.unfinished[j] = 0;
            st.finish_step[j] = i;
            t11 = t12 = t21 = t22 = tk.eos;
        }
    } else {
        t11 = t12 = t21 = t22 = tk.pad;
    }
    st.gen_kind[static_cast<int64_t>(j) * st.max_len + i] = kind;
    st.gen_xy[(static_cast<int64_t>(j) * st.max_len + i) * 2] = gx;
    st.gen_xy[(static_cast<int64_t>(j) * st.max_len + i) * 2 + 1] = gy;
    // inputs of step i + 1 (the reference appends to prev_output_token_* / delta_* and slices column i+1 next time)
    st.seq11[j] = t11;
    st.seq12[j] = t12;
    st.seq21[j] = t21;
    st.seq22[j] = t22;
    st.delta_x1[j] = dx;
    st.delta_y1[j] = dy;
    st.delta_x2[j] = __fsub_rn(1.f, dx);
    st.delta_y2[j] = __fsub_rn(1.f, dy);
}

// Last kernel of a step: advance the device-resident step counter (after every sample has read it).
__global__ void advance_step_kernel(int64_t* step_dev) { *step_dev += 1; }

}  // namespace

cudaError_t launch_seq_embed_forward(const SeqEmbedArgs& a, cudaStream_t stream) {
    if (a.tokens == 0) return cudaSuccess;
    const int warps = 4;
    const int64_t grid = (a.tokens + warps - 1) / warps;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    seq_embed_fwd_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
        a.table, a.seq11, a.seq12, a.seq21, a.seq22, a.dx1, a.dx2, a.dy1, a.dy2, a.out, a.tokens, a.C, a.V);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_seq_embed_backward(const SeqEmbedArgs& a, cudaStream_t stream) {
    if (a.tokens == 0) return cudaSuccess;
    const int warps = 4;
    const int64_t grid = (a.tokens + warps - 1) / warps;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    seq_embed_bwd_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, stream>>>(
        a.grad_out, a.seq11, a.seq12, a.seq21, a.seq22, a.dx1, a.dx2, a.dy1, a.dy2, a.grad_table, a.tokens, a.C, a.V,
        a.padding_idx);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_token_step(const float* cls_logits, const float* reg, int64_t* step_dev, const cape_token_state& st,
                              const cape_tokenizer& tk, int B, int n_classes, cudaStream_t stream) {
    if (B == 0) return cudaSuccess;
    token_step_kernel<<<(B + 127) / 128, 128, 0, stream>>>(cls_logits, reg, step_dev, st, tk, B, n_classes);
    advance_step_kernel<<<1, 1, 0, stream>>>(step_dev);
    count_launch();
    count_launch();
    return cudaGetLastError();
}

}  // namespace cape
