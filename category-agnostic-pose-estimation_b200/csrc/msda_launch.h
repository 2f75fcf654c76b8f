// Internal launch interface between the C ABI (cape_abi.cu) and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cape_msda.h"

namespace cape {

constexpr int kMaxLevels = 8;
constexpr int kMaxPoints = 8;

// What the sampling kernels read their per-sample (x, y, weight) from.
//   direct: sampling_locations + attention_weights (aux dtype)          — cape_msda_forward / backward
//   fused : reference_points + raw offsets + raw logits (fp32), softmax and the location arithmetic of
//           MSDeformAttn.forward (deformable_transformer.py:99-105) done in the kernel — cape_msda_decode
struct FwdArgs {
    const void* value;
    const int64_t* shapes;
    const int64_t* starts;
    const void* loc;          // direct: sampling_locations; fused: sampling_offsets
    const void* attn;         // direct: attention_weights;  fused: attention_logits
    const float* ref_points;  // fused only
    void* out;
    cape_msda_dims d;
    int value_dtype;
    int aux_dtype;
    bool fused;
};

struct BwdArgs {
    const void* grad_out;
    const void* value;
    const int64_t* shapes;
    const int64_t* starts;
    const void* loc;          // direct: sampling_locations; fused: sampling_offsets
    const void* attn;         // direct: attention_weights;  fused: attention_logits
    const float* ref_points;  // fused only
    float* grad_value;
    void* grad_loc;           // fused: grad_sampling_offsets
    void* grad_attn;          // fused: grad_attention_logits
    cape_msda_dims d;
    int value_dtype;
    int aux_dtype;
    bool fused;
};

// Planar point sampling (MSDeformablePoints): x viewed (B*G, c, H, W), positions (B*G, Hk, Wk, 2), out (B, Hk*Wk, G*c).
struct PointsDims {
    int B, G, c, H, W, Hk, Wk;
};

// Bilinear token embedding (TransformerDecoder._seq_embed); forward uses table / out, backward grad_out / grad_table.
struct SeqEmbedArgs {
    const float* table;
    const float* grad_out;
    const int64_t *seq11, *seq12, *seq21, *seq22;
    const float *dx1, *dx2, *dy1, *dy2;
    float* out;
    float* grad_table;
    int64_t tokens;
    int C, V;
    int64_t padding_idx;
};

// Each returns cudaGetLastError() after the launch and bumps the launch counter.
cudaError_t launch_forward(const FwdArgs& a, cudaStream_t stream);
cudaError_t launch_backward(const BwdArgs& a, cudaStream_t stream);
// shared-memory staged variants (persistent, TMA): return cudaErrorNotSupported for configurations they do not cover
cudaError_t launch_forward_staged(const FwdArgs& a, cudaStream_t stream);
cudaError_t launch_backward_staged(const BwdArgs& a, int mode, cudaStream_t stream);   // mode 2: + tensor-core scatter, 3: staged only
// small CTAs, coarsest level scattered by tcgen05 (msda_backward_tc.cu); cudaErrorNotSupported outside its configurations
cudaError_t launch_backward_tc(const BwdArgs& a, cudaStream_t stream);
cudaError_t read_backward_staged_cycles(long long* out16, bool reset);   // PROFILE knob: cycle counters of CTA 0
cudaError_t launch_query_pool_forward(const FwdArgs& a, cudaStream_t stream);     // fp32 only
cudaError_t launch_query_pool_backward(const BwdArgs& a, cudaStream_t stream);    // fp32 only
cudaError_t launch_points_sample_forward(const float* x, const float* pos, float* out, const PointsDims& p,
                                         cudaStream_t stream);
cudaError_t launch_points_sample_backward(const float* gout, const float* x, const float* pos, float* gx, float* gpos,
                                          const PointsDims& p, cudaStream_t stream);
// y = epilogue(x W^T + b) for a few rows (decode step).  wt is the weight TRANSPOSED, (K, N) row-major.
struct SkinnyArgs {
    const float* x;            // (rows, K) with row stride x_stride; in sine mode the (rows, 2) reference points
    const float* x2;           // optional addend on the input, row stride x2_stride
    const float* wt;
    const float* bias;         // (N) or NULL
    const float* res;          // epilogue 2: optional residual (rows, N), row stride res_stride
    const float* gamma;        // epilogue 2: LayerNorm weight / bias (N)
    const float* beta;
    const float* sine_dim_t;   // non-NULL: input = sine embedding of x with these 128 divisors (K must be 256)
    float* y;                  // (rows, N), row stride y_stride
    float* y2;                 // optional second output: columns >= split go to y2[:, c - split] (row stride y2_stride)
    int rows, K, N;
    int x_stride, x2_stride, res_stride, y_stride, y2_stride, split;
    float eps;
    // epilogue 3 (coordinate head + refinement)
    const float* w3;           // (2, N)
    const float* b3;           // (2)
    const float* ref_in;       // (rows, 2)
    const float* valid_ratios; // (rows, n_levels, 2)
    float* ref_out;            // (rows, 2)
    float* ref_levels;         // (rows, n_levels, 2)
    int n_levels;
    // input = MSDeformAttn sampling (fused softmax / location prologue) on a projected-value cache instead of x: row r of the
    // input is the sampled (M x 32)-vector of query r.  LayerNorm epilogue only; K = msda_M * 32, 4 levels x 4 points.
    const float* msda_value;   // (B, S, M, 32) fp32
    const int64_t* msda_shapes;
    const int64_t* msda_starts;
    const float* msda_ref;     // (rows, L, 2)
    const float* msda_off;     // (rows, M, L, P, 2) raw offsets
    const float* msda_logits;  // (rows, M, L * P) raw logits
    int msda_S, msda_M, msda_Lq;
};

cudaError_t launch_decode_attention(const float* q, const float* k_new, const float* v_new, float* k_cache, float* v_cache,
                                    const int64_t* pos_dev, const float* key_bias, float* out, int B, int T, int H,
                                    int q_stride, int new_stride, cudaStream_t stream);
cudaError_t launch_skinny_linear(const SkinnyArgs& a, int epilogue, cudaStream_t stream);
cudaError_t launch_tiny_linear(const float* x, int x_stride, const float* w, const float* bias, const float* refine_ref,
                               float* y, int rows, int K, int N, cudaStream_t stream);
cudaError_t launch_tf32_split_lo(const float* x, float* lo, int64_t n, cudaStream_t stream);
cudaError_t launch_linear_tf32x3(const float* x, const float* w, const float* w_lo, const float* bias, float* y, int M, int N,
                                 int K, int act, int split_k, cudaStream_t stream);
cudaError_t launch_wgrad_tf32x3(const float* grad_out, const float* x, float* grad_w, int rows, int N, int K, cudaStream_t stream);
cudaError_t launch_transpose_lo(const float* in, float* out, float* out_lo, int R, int C, cudaStream_t stream);
cudaError_t launch_seq_embed_forward(const SeqEmbedArgs& a, cudaStream_t stream);
cudaError_t launch_seq_embed_backward(const SeqEmbedArgs& a, cudaStream_t stream);
cudaError_t launch_token_step(const float* cls_logits, const float* reg, int64_t* step_dev, const cape_token_state& st,
                              const cape_tokenizer& tk, int B, int n_classes, cudaStream_t stream);

cudaError_t launch_zero_masked_rows(void* value, const uint8_t* mask, long long rows, int row_bytes, cudaStream_t stream);

void count_launch();

// Tile size (queries per CTA) that fills whole waves: with `slots` CTAs resident on the chip, a grid of N*M*ceil(Lq/qpc)
// CTAs just over a multiple of `slots` leaves the last wave nearly empty (N = 2, Lq = 5440: 688 CTAs on 592 slots ran the
// backward at 1.46x its pro-rata time).  Keep the wave count the default tile would need and size the tiles so the grid
// fills those waves.  `align` = queries a warp sweep covers (4 in the forward).
inline int balanced_q_per_cta(int64_t nm, int lq, int q_default, int slots, int align) {
    if (nm <= 0 || lq <= 0) return q_default;
    const int64_t tiles0 = (lq + q_default - 1) / q_default;
    const int64_t waves = (nm * tiles0 + slots - 1) / slots;
    int64_t tiles = (waves * slots) / nm;                   // tiles per (image, head) that fit `waves` full waves
    if (tiles < 1) tiles = 1;
    if (tiles > lq) tiles = lq;
    int qpc = static_cast<int>((lq + tiles - 1) / tiles);
    if (align > 1) qpc = (qpc + align - 1) / align * align;
    return qpc < 1 ? 1 : qpc;
}

// Tuning knobs: read once per process from the environment variable "CAPE_<NAME>" (std::call_once), changeable at run
// time through cape_set_tuning() (tools/tune.py).  Values <= 0 mean "default".
enum Tune {
    kTuneFwdThreads, kTuneFwdQpc, kTuneFwdPointMaxQm, kTuneFwdStaged, kTuneFwdStagedMinQm, kTuneFwdStagedKb,
    kTuneBwdThreads, kTuneBwdQpc, kTuneBwdMode, kTuneBwdStagedKb, kTuneProfile, kTuneBwdTcMinQm, kTuneHostChunks, kTuneWgradTranspose, kTuneCount
};
int tuning(Tune knob, int fallback);

// True the first time it is called for the current device with this `flags` word (one word per kernel family): function
// attributes such as the dynamic shared-memory opt-in are per device, so a once-per-process flag is not enough.
bool first_use_on_device(unsigned long long* flags);

}  // namespace cape
