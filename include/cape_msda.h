/*
 * cape_msda.h — C ABI of libcape_msda.so: multi-scale deformable attention for NVIDIA B200 (sm_100a).
 *
 * The reference (nkkrnkl/category-agnostic-pose-estimation) has no native code and no FFI; its seam for
 * this path is the Python call
 *     output = ms_deform_attn_core_pytorch(value, input_spatial_shapes, sampling_locations, attention_weights)
 * at models/deformable_transformer.py:112 (function body :115-141), reached from MSDeformAttn.forward
 * (:76-114).  The entry points below are what a ctypes binding on that seam calls; argument order mirrors
 * upstream Deformable-DETR's MSDeformAttnFunction
 *     (value, spatial_shapes, level_start_index, sampling_locations, attention_weights).
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - Every pointer named *_dev / value / loc / attn / out / grad_* is DEVICE memory on the current CUDA device,
 *     C-contiguous, 16-byte aligned.  The caller owns every buffer (inputs, outputs, workspaces); the library
 *     never allocates, frees or retains a pointer beyond the call.
 *   - All calls are asynchronous: work is enqueued on `stream` (a cudaStream_t / CUstream; NULL = legacy default
 *     stream); nothing synchronises the device.  Re-entrant and thread-safe (autograd calls backward from its own
 *     worker thread).
 *   - Return value: 0 on success; > 0 a cudaError_t; < 0 a CAPE_ERR_* code.  cape_last_error() returns a
 *     thread-local message for the last failing call on the calling thread.
 *   - There is no CPU implementation: host pointers passed where device pointers are expected are rejected with
 *     CAPE_ERR_NOT_DEVICE_PTR.
 *
 * Layouts (reference shapes, models/deformable_transformer.py:76-141)
 *   value               (N, S, M, D)           S = sum_l H_l*W_l; level l is rows start_l .. start_l + H_l*W_l
 *   spatial_shapes      (L, 2) int64           rows (H_l, W_l)
 *   level_start_index   (L,)   int64
 *   sampling_locations  (N, Lq, M, L, P, 2)    last dim (x, y), normalised to [0,1], unclamped
 *   attention_weights   (N, Lq, M, L, P)
 *   output              (N, Lq, M*D)
 */
#ifndef CAPE_MSDA_H_
#define CAPE_MSDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CAPE_API __attribute__((visibility("default")))
#else
#define CAPE_API
#endif

#define CAPE_ABI_VERSION 1

/* element types */
#define CAPE_DTYPE_F32 0
#define CAPE_DTYPE_BF16 1
#define CAPE_DTYPE_F16 2

/* library error codes (negative; positive values are cudaError_t) */
#define CAPE_ERR_BAD_DIMS (-1)        /* a dimension is <0, L>8, P>8, D not a multiple of 4 or >256, N*S*M*D overflow */
#define CAPE_ERR_BAD_DTYPE (-2)       /* unknown or unsupported dtype combination */
#define CAPE_ERR_NULL_PTR (-3)        /* a required pointer is NULL */
#define CAPE_ERR_MISALIGNED (-4)      /* a tensor base pointer is not 16-byte aligned */
#define CAPE_ERR_NOT_DEVICE_PTR (-5)  /* a tensor pointer is not device-accessible memory */
#define CAPE_ERR_WORKSPACE (-6)       /* workspace too small */

/* Problem dimensions, shared by every entry point. */
typedef struct cape_msda_dims {
    int32_t N;   /* batch (episodes x queries per episode)        */
    int32_t S;   /* value rows per batch element = sum_l H_l*W_l  */
    int32_t M;   /* heads                                         */
    int32_t D;   /* channels per head (32 in CAPE)                */
    int32_t Lq;  /* queries                                       */
    int32_t L;   /* levels  (<= 8)                                */
    int32_t P;   /* points per level (<= 8)                       */
} cape_msda_dims;

CAPE_API int cape_abi_version(void);
CAPE_API const char* cape_last_error(void);

/*
 * Forward.  Replaces ms_deform_attn_core_pytorch (models/deformable_transformer.py:115-141).
 *   value_dtype : CAPE_DTYPE_* of value and out.
 *   aux_dtype   : CAPE_DTYPE_* of sampling_locations and attention_weights (F32, or the same as value_dtype).
 */
CAPE_API int cape_msda_forward(const void* value, const int64_t* spatial_shapes_dev, const int64_t* level_start_index_dev,
                      const void* sampling_locations, const void* attention_weights, void* out,
                      const cape_msda_dims* dims, int value_dtype, int aux_dtype, void* stream);

/*
 * Backward (what autograd derives from :129-141 through grid_sampler_2d_backward).
 *   grad_out          (N, Lq, M*D)   value_dtype
 *   grad_value        (N, S, M, D)   ALWAYS fp32, accumulated with atomics: the caller zero-fills it first
 *                                    (or passes zero_grad_value = 1 to have the library enqueue the memset).
 *   grad_loc          (N, Lq, M, L, P, 2)   aux_dtype   fully overwritten
 *   grad_attn         (N, Lq, M, L, P)      aux_dtype   fully overwritten
 */
CAPE_API int cape_msda_backward(const void* grad_out, const void* value, const int64_t* spatial_shapes_dev,
                       const int64_t* level_start_index_dev, const void* sampling_locations,
                       const void* attention_weights, float* grad_value, void* grad_loc, void* grad_attn,
                       const cape_msda_dims* dims, int value_dtype, int aux_dtype, int zero_grad_value,
                       void* stream);

/*
 * Incremental decode: the prologue of MSDeformAttn.forward fused with the core, on a cached projected value
 * (the role of the reference's dead VCache, models/kv_cache.py:37-70; call site deformable_transformer_v2.py:360-363).
 *   value_cache        (B, S, M, D)      value_dtype — value_proj(memory), written once per sequence
 *   reference_points   (B, k, L, 2)      fp32        — deformable_transformer.py:102-105 (2-d branch)
 *   sampling_offsets   (B, k, M, L, P, 2) fp32       — raw output of the sampling_offsets Linear (:99)
 *   attention_logits   (B, k, M, L*P)    fp32        — raw output of the attention_weights Linear (:100), pre-softmax
 *   out                (B, k, M*D)       value_dtype
 * dims->N = B, dims->Lq = k (the newly appended tokens only).
 */
CAPE_API int cape_msda_decode(const void* value_cache, const int64_t* spatial_shapes_dev, const int64_t* level_start_index_dev,
                     const float* reference_points, const float* sampling_offsets, const float* attention_logits,
                     void* out, const cape_msda_dims* dims, int value_dtype, void* stream);

/*
 * Fused module path for training: cape_msda_decode is also the FORWARD of the op
 *     out = core(value, softmax(attention_logits), reference_points + sampling_offsets / (W_l, H_l))
 * for any number of queries (deformable_transformer.py:100-105 + :112), and this is its backward.  sampling_locations,
 * attention_weights and their gradients never round-trip HBM.
 *   grad_offsets   (N, Lq, M, L, P, 2) fp32   d out / d sampling_offsets         fully overwritten
 *   grad_logits    (N, Lq, M, L*P)     fp32   d out / d attention_logits (through the softmax)   fully overwritten
 *   grad_value     (N, S, M, D)        fp32   accumulated with atomics (zero_grad_value as in cape_msda_backward)
 * The gradient w.r.t. reference_points is sum over (m, p) of grad_offsets * (W_l, H_l); callers that need it reduce
 * grad_offsets themselves.  Fast-path dimensions only (cape_msda_fused_supported: D = 32, P = 4, L <= 4); other
 * dimensions return CAPE_ERR_BAD_DIMS and the caller composes softmax / location arithmetic with cape_msda_backward.
 */
CAPE_API int cape_msda_fused_supported(const cape_msda_dims* dims);
CAPE_API int cape_msda_fused_backward(const void* grad_out, const void* value, const int64_t* spatial_shapes_dev,
                                      const int64_t* level_start_index_dev, const float* reference_points,
                                      const float* sampling_offsets, const float* attention_logits, float* grad_value,
                                      float* grad_offsets, float* grad_logits, const cape_msda_dims* dims,
                                      int value_dtype, int zero_grad_value, void* stream);

/*
 * Variant samplers (fp32 only; non-default paths of the reference, small problems).
 *
 * Query-pooled deformable sampling — TransformerDecoderLayerV4._sample_reference_points,
 * models/deformable_transformer_v2.py:661-687: the sampling of :115-141, but the weighted sum runs over the queries:
 *     out[n, l*P+p, m*D+d] = sum_q attention_weights[n,q,m,l,p] * bilinear(V_l[n,:,m,d], sampling_locations[n,q,m,l,p])
 *   out / grad_out (N, L*P, M*D); every other tensor as in cape_msda_forward / cape_msda_backward.
 *
 * Planar point sampling — MSDeformablePoints.forward, models/deformable_points.py:118-128:
 * F.grid_sample(bilinear, zeros padding, align_corners=True) of a level whose contiguous (B, H*W, C) block is VIEWED as
 * (B*G, c, H, W) (the reference's reshape at :125, C = G*c), at positions pos (B*G, Hk, Wk, 2) given as (y, x) in
 * [-1, 1]; out (B, Hk*Wk, G*c) is the layout of :128.  grad_x (same shape as x) is accumulated with atomics
 * (zero_grad_x = 1 lets the library zero it first); grad_pos (B*G, Hk, Wk, 2) is fully overwritten.
 */
CAPE_API int cape_msda_query_pool_forward(const float* value, const int64_t* spatial_shapes_dev,
                                          const int64_t* level_start_index_dev, const float* sampling_locations,
                                          const float* attention_weights, float* out, const cape_msda_dims* dims,
                                          void* stream);
CAPE_API int cape_msda_query_pool_backward(const float* grad_out, const float* value, const int64_t* spatial_shapes_dev,
                                           const int64_t* level_start_index_dev, const float* sampling_locations,
                                           const float* attention_weights, float* grad_value, float* grad_loc,
                                           float* grad_attn, const cape_msda_dims* dims, int zero_grad_value,
                                           void* stream);
CAPE_API int cape_points_sample_forward(const float* x, const float* pos, float* out, int B, int G, int c, int H, int W,
                                        int Hk, int Wk, void* stream);
CAPE_API int cape_points_sample_backward(const float* grad_out, const float* x, const float* pos, float* grad_x,
                                         float* grad_pos, int B, int G, int c, int H, int W, int Hk, int Wk,
                                         int zero_grad_x, void* stream);

/*
 * Padding-mask fill of the projected value — `value = value.masked_fill(input_padding_mask[..., None], 0)`,
 * models/deformable_transformer.py:96-97 — in place: value (rows, row_bytes / elem) row-major, mask (rows,) bytes
 * (torch.bool), row_bytes % 16 == 0.  Rows are inspected on the device in blocks of 256; a block without a masked row
 * returns without touching `value`, so the all-False mask CAPE always passes costs one pass over the MASK only and no
 * host synchronisation.
 */
CAPE_API int cape_zero_masked_rows(void* value, const uint8_t* mask, int64_t rows, int row_bytes, void* stream);

/*
 * Sequence side of the decoder: the data formats either side of the decode step, kept on the device.
 *
 * Bilinear token embedding — TransformerDecoder._seq_embed, models/deformable_transformer_v2.py:984-997:
 *     out[t, :] = E[seq11[t]]*dx2[t]*dy2[t] + E[seq21[t]]*dx1[t]*dy2[t] + E[seq12[t]]*dx2[t]*dy1[t] + E[seq22[t]]*dx1[t]*dy1[t]
 *   table (V, C) fp32, C % 4 == 0; seq* (tokens,) int64; d* (tokens,) fp32; out (tokens, C) fp32.
 *   A token id outside [0, V) yields a NaN row (the reference raises an IndexError; a kernel cannot).
 *   Backward accumulates into grad_table (V, C) with atomics (zero_grad_table = 1 zero-fills it first) and skips
 *   padding_idx (pass -1 for none), like nn.Embedding(padding_idx=...).
 */
CAPE_API int cape_seq_embed_forward(const float* table, const int64_t* seq11, const int64_t* seq12, const int64_t* seq21,
                                    const int64_t* seq22, const float* delta_x1, const float* delta_x2,
                                    const float* delta_y1, const float* delta_y2, float* out, int64_t tokens, int C,
                                    int V, void* stream);
CAPE_API int cape_seq_embed_backward(const float* grad_out, const int64_t* seq11, const int64_t* seq12,
                                     const int64_t* seq21, const int64_t* seq22, const float* delta_x1,
                                     const float* delta_x2, const float* delta_y1, const float* delta_y2,
                                     float* grad_table, int64_t tokens, int C, int V, int64_t padding_idx,
                                     int zero_grad_table, void* stream);

/*
 * Token bookkeeping of one autoregressive step — the `for j in range(bs)` body of RoomFormerV2.forward_inference,
 * models/roomformer_v2.py:548-597, for every sample at once and without leaving the device:
 *   cls_logits (B, n_classes), reg (B, 2): this step's head outputs (fp32);
 *   step_dev: device-resident step counter i, incremented by the call (so a captured CUDA graph can be replayed);
 *   state: per-sample buffers the call reads and updates (all device memory, caller-owned):
 *     unfinished (B) int32 in/out; finish_step (B) int64, written when a sample emits its terminating <eos>;
 *     seq11..seq22 (B) int64 and delta_x1..delta_y2 (B) fp32: OUTPUT, the next step's _seq_embed inputs;
 *     pred_logits (B, max_len, n_classes), pred_coords (B, max_len, 2): column i receives cls_logits / reg;
 *     gen_kind (B, max_len) int32 and gen_xy (B, max_len, 2): the reference's gen_out entry of column i
 *       (kind 0 = [x, y], 2 = separator, -1 = everything else);
 *   tokenizer: DiscreteTokenizer constants (datasets/discrete_tokenizer.py:7-28) and TokenType values
 *     (datasets/token_types.py), min_len = 6 (roomformer_v2.py:456).
 * A step counter outside [0, max_len) makes the call a no-op (the counter is still advanced).
 */
typedef struct cape_tokenizer {
    int32_t num_bins;
    int32_t min_len;
    int64_t bos, eos, sep, pad, cls;
    int32_t type_coord, type_sep, type_eos, type_cls;
} cape_tokenizer;

typedef struct cape_token_state {
    int32_t* unfinished;
    int64_t* finish_step;
    int64_t *seq11, *seq12, *seq21, *seq22;
    float *delta_x1, *delta_x2, *delta_y1, *delta_y2;
    float* pred_logits;
    float* pred_coords;
    int32_t* gen_kind;
    float* gen_xy;
    int64_t max_len;
} cape_token_state;

CAPE_API int cape_token_step(const float* cls_logits, const float* reg, int64_t* step_dev, const cape_token_state* state,
                             const cape_tokenizer* tokenizer, int B, int n_classes, void* stream);

/*
 * Kernels of the incremental-decode step around cape_msda_decode (one new token per sequence; fp32).
 *
 * cape_decode_attention — attention of the new token over a K/V cache, head dimension 32
 *   (TransformerDecoderLayer.forward, models/deformable_transformer_v2.py:322-341 with models/kv_cache.py:3-36, and the
 *   support cross-attention :350-357).  q (B, H*32) is the in-projected query; q_stride / new_stride are the row strides
 *   (elements, multiples of 4) of q and of k_new / v_new, so slices of one fused projection output can be passed.
 *   Self-attention form: pos_dev != NULL, k_new / v_new (B, H*32) are the token's in-projected key / value; they are
 *   written to k_cache / v_cache (B, T, H*32) at row *pos_dev and positions 0..*pos_dev are attended.  A position
 *   outside [0, T) makes the call a no-op.
 *   Cross-attention form: pos_dev = k_new = v_new = NULL, all T cached keys are attended; key_bias (B, T) or NULL is
 *   added to the scores (-inf at padded keys).  T <= 1024.  out (B, H*32), before the output projection.
 *
 * cape_skinny_linear — y = epilogue(x W^T + b) for `rows` rows: wt (K, N) is the weight TRANSPOSED; K % 16 == 0, K <= 2048, N % 4 == 0.
 *   epilogue 0: bias; 1: bias + ReLU; 2: LayerNorm_N(residual + x W^T + b) * gamma + beta (N <= 256, residual may be NULL).
 *   x2 (optional) is added to x first (tgt + query_pos).  With sine_dim_t != NULL (128 divisors, K = 256) the input is
 *   the sine embedding of the (rows, 2) reference points in x (TransformerDecoder.get_query_pos_embed, :1005-1018).
 *   Strides are row strides in elements.
 *
 * cape_tiny_linear — y (rows, N) = x W^T + b with w (N, K) as a Linear stores it, N <= 8, K % 4 == 0; with refine_ref
 *   (rows, N) != NULL the result is sigmoid(y + inverse_sigmoid(refine_ref)) (iterative refinement, :1096-1102).
 */
CAPE_API int cape_decode_attention(const float* q, int q_stride, const float* k_new, const float* v_new, int new_stride,
                                   float* k_cache, float* v_cache, const int64_t* pos_dev, const float* key_bias,
                                   float* out, int B, int T, int H, int D, void* stream);
CAPE_API int cape_skinny_linear(const float* x, int x_stride, const float* x2, int x2_stride, const float* wt,
                                const float* bias, const float* residual, int residual_stride, const float* gamma,
                                const float* beta, float eps, const float* sine_dim_t, float* y, int y_stride, int rows,
                                int K, int N, int epilogue, void* stream);
/*
 * Two fusions of the decode step on top of cape_skinny_linear (same constraints on x / wt / K / N):
 *   cape_skinny_linear_split — epilogue 0 with the output columns split over two tensors: columns < split go to y,
 *     the rest to y2 (MSDeformAttn's sampling_offsets | attention_weights projections of one query, :99-100).
 *   cape_coord_head_refine — the last two layers of the coordinate MLP and the iterative refinement in one launch:
 *     h = relu(x wt + bias) (N <= 256), ref_out = sigmoid(h w3^T + b3 + inverse_sigmoid(ref_in)) with w3 (2, N), and
 *     ref_levels[r, l, :] = ref_out[r, :] * valid_ratios[r, l, :] (deformable_transformer_v2.py:1072, 1096-1102).
 */
CAPE_API int cape_skinny_linear_split(const float* x, int x_stride, const float* x2, int x2_stride, const float* wt,
                                      const float* bias, float* y, int y_stride, float* y2, int y2_stride, int split,
                                      int rows, int K, int N, void* stream);
CAPE_API int cape_coord_head_refine(const float* x, int x_stride, const float* wt, const float* bias, const float* w3,
                                    const float* b3, const float* ref_in, const float* valid_ratios, float* ref_out,
                                    float* ref_levels, int rows, int K, int N, int n_levels, void* stream);
CAPE_API int cape_tiny_linear(const float* x, int x_stride, const float* w, const float* bias, const float* refine_ref,
                              float* y, int rows, int K, int N, void* stream);
/*
 * cape_msda_output_proj — MSDeformAttn's sampling AND its output projection, residual and LayerNorm in one launch
 *   (MSDeformAttn.forward, models/deformable_transformer.py:99-113, followed by the decoder layer's norm1,
 *   deformable_transformer_v2.py:360-364):  y = LayerNorm_N(residual + sample(value_cache, ...) wt + bias) * gamma + beta.
 *   The sampled (N * Lq, M * 32) rows are the input tile of the linear and never exist in HBM; arguments and arithmetic of
 *   the sampling are those of cape_msda_decode (fp32 value cache, raw offsets and logits, reference points), of the linear
 *   those of cape_skinny_linear with epilogue 2 (wt (M * 32, n_out) transposed weight, n_out <= 256).  D = 32, P = 4, L = 4.
 */
CAPE_API int cape_msda_output_proj(const float* value_cache, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                   const float* reference_points, const float* sampling_offsets,
                                   const float* attention_logits, const cape_msda_dims* dims, const float* wt,
                                   const float* bias, const float* residual, int residual_stride, const float* gamma,
                                   const float* beta, float eps, float* y, int y_stride, int n_out, void* stream);

/*
 * fp32-accurate linear layer on the tensor cores ("3xTF32", tcgen05.mma.kind::tf32 with TMA-fed operands) for the
 * projections around the sampling op (value_proj / output_proj / FFN, models/deformable_transformer.py:95,113,219-224):
 *     y (M, N) = act(x (M, K) . w (N, K)^T + bias)        act: 0 none, 1 ReLU
 * with every operand used as hi + lo (hi = the 19 bits kind::tf32 reads, lo = the exact remainder) and
 * x_lo w_hi + x_hi w_lo + x_hi w_hi accumulated in fp32.  w_lo (N, K) comes from cape_tf32_split_lo(w) once per weight;
 * x_lo is produced inside the kernel.  x, w, w_lo, y row-major fp32, 16-byte aligned; K % 32 == 0, N % 128 == 0.
 */
CAPE_API int cape_tf32_split_lo(const float* x, float* lo, int64_t n, void* stream);
CAPE_API int cape_linear_tf32x3(const float* x, const float* w, const float* w_lo, const float* bias, float* y, int M, int N,
                                int K, int act, void* stream);
/*
 * Weight gradient of that layer, grad_w (N, K) = grad_out (rows, N)^T . x (rows, K), on the same kernel: both operands are
 * read where they are as MN-major tiles (the row index is the reduction dimension; 128-byte swizzle with 32-byte atoms, the
 * one layout the tensor core takes for 32-bit MN-major operands), their lo parts are split in shared memory, the reduction
 * is cut into chains of <= 1024 rows spread over the SMs and the partial tiles are added into grad_w by the TMA
 * (cp.reduce.async.bulk.tensor ... add); grad_w is zero-filled by the call.  N % 32 == 0, K % 128 == 0; `workspace` may be
 * NULL.  (Tuning knob WGRAD_TRANSPOSE=1 selects the earlier form on transposed copies: workspace of (N + 2 K) * rows floats,
 * rows % 32 == 0.)
 */
CAPE_API int cape_linear_tf32x3_wgrad(const float* grad_out, const float* x, float* grad_w, float* workspace, int rows, int N,
                                      int K, void* stream);

/*
 * Host-buffer round trip used for end-to-end measurement and for callers without device buffers:
 * copies the inputs from (ideally pinned) HOST memory into the caller-provided device workspace, runs forward and,
 * when grad_out_host != NULL, backward, and copies the results back to HOST memory — all enqueued on `stream`.
 * fp32 only.  cape_msda_host_workspace_bytes() gives the device workspace size for `dims`.
 * Large batches are split into up to 32 chunks along N (one image each at N <= 32) and software-pipelined: the H2D copy of chunk c+1 and the D2H copy
 * of chunk c-1 run on two library-owned helper streams (created once per device) while chunk c's kernels run on
 * `stream`; `stream` is joined to both before the call's work counts as complete, so the caller still only
 * synchronises `stream`.  This is the one place the library owns CUDA resources (two streams per device; concurrent
 * callers on different streams of one device serialise through them).  The host buffers must be PINNED for the copies to
 * overlap the kernels (pageable buffers work but every cudaMemcpyAsync then blocks).  On any error the events created for
 * the call are destroyed and the helper streams are joined back into `stream` before the error code is returned.
 */
CAPE_API size_t cape_msda_host_workspace_bytes(const cape_msda_dims* dims, int with_backward);
CAPE_API int cape_msda_forward_backward_host(const float* value_host, const int64_t* spatial_shapes_host,
                                    const int64_t* level_start_index_host, const float* loc_host,
                                    const float* attn_host, const float* grad_out_host, float* out_host,
                                    float* grad_value_host, float* grad_loc_host, float* grad_attn_host,
                                    const cape_msda_dims* dims, void* workspace_dev, size_t workspace_bytes,
                                    void* stream);

/* Counts kernels this library has launched since load (bench.py reports it as gpu_launches). */
CAPE_API uint64_t cape_launch_count(void);

/*
 * Launch-geometry / kernel-selection overrides for tuning and profiling runs.  Every knob is also read ONCE per process
 * from the environment variable of the same name (CAPE_FWD_THREADS, CAPE_FWD_QPC, CAPE_FWD_POINT_MAX_QM, CAPE_FWD_STAGED,
 * CAPE_BWD_THREADS, CAPE_BWD_QPC, CAPE_BWD_MODE ...); this call changes it at run time (value <= 0: back to the
 * default).  Returns 0, or CAPE_ERR_BAD_DIMS for an unknown name.  Not part of the reference-facing interface.
 * BWD_MODE: 1 (default) small CTAs, vector REDs for every level; 2 persistent CTAs, TMA-staged coarse rows + tcgen05 scatter of
 * the two coarse levels; 3 staged rows only; 5 small CTAs + tcgen05 scatter of the coarsest level (4 is a profiling aid with
 * invalid results).  PROFILE is read by modes 2-5 only and must stay 0 outside timing experiments.
 */
CAPE_API int cape_set_tuning(const char* name, int value);
CAPE_API int cape_get_tuning(const char* name);
/* With the PROFILE knob set, CTA 0 of the staged backward kernel accumulates clock64() cycles per phase of its builder warp
 * (slots 0..8), the batch count (9) and the sampling time of warp 0 (10).  Copies the 16 counters to the host (synchronises). */
CAPE_API int cape_debug_counters(long long* out16, int reset);

#ifdef __cplusplus
}
#endif
#endif /* CAPE_MSDA_H_ */
