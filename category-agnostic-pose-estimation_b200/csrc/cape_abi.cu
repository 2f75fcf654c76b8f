// C ABI of libcape_msda.so (declared in include/cape_msda.h).  Validation, dtype dispatch and stream plumbing only —
// the kernels live in msda_forward.cu / msda_backward.cu.  No allocation, no device synchronisation, no CPU path.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "msda_launch.h"

namespace cape {

namespace {

std::atomic<uint64_t> g_launches{0};
thread_local char t_error[512] = "";

int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}

int fail_cuda(cudaError_t e, const char* what) {
    snprintf(t_error, sizeof(t_error), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
}

size_t dtype_size(int dt) {
    switch (dt) {
        case CAPE_DTYPE_F32: return 4;
        case CAPE_DTYPE_BF16:
        case CAPE_DTYPE_F16: return 2;
    }
    return 0;
}

int check_dims(const cape_msda_dims* d) {
    if (!d) return fail(CAPE_ERR_NULL_PTR, "dims is NULL");
    if (d->N < 0 || d->S < 0 || d->Lq < 0 || d->M <= 0 || d->D <= 0 || d->L <= 0 || d->P <= 0)
        return fail(CAPE_ERR_BAD_DIMS, "negative or zero dimension (N=%d S=%d M=%d D=%d Lq=%d L=%d P=%d)", d->N, d->S, d->M,
                    d->D, d->Lq, d->L, d->P);
    if (d->L > kMaxLevels || d->P > kMaxPoints)
        return fail(CAPE_ERR_BAD_DIMS, "L=%d (max %d) or P=%d (max %d) out of range", d->L, kMaxLevels, d->P, kMaxPoints);
    if (d->D > 256 || d->D % 4 != 0)
        return fail(CAPE_ERR_BAD_DIMS, "D=%d must be a multiple of 4 and <= 256", d->D);
    if (static_cast<int64_t>(d->S) * d->M * d->D > 0x7fffffffLL)
        return fail(CAPE_ERR_BAD_DIMS, "S*M*D=%lld exceeds 2^31-1", static_cast<long long>(d->S) * d->M * d->D);
    return 0;
}

int check_dtypes(int value_dtype, int aux_dtype) {
    if (dtype_size(value_dtype) == 0) return fail(CAPE_ERR_BAD_DTYPE, "unknown value dtype %d", value_dtype);
    if (aux_dtype != CAPE_DTYPE_F32 && aux_dtype != value_dtype)
        return fail(CAPE_ERR_BAD_DTYPE, "aux dtype %d must be fp32 or equal to the value dtype %d", aux_dtype, value_dtype);
    return 0;
}

// Device pointer check: NULL, alignment, and that CUDA knows the allocation as device (or managed) memory.
int check_ptr(const void* p, const char* name, bool empty_ok, unsigned align = 16) {
    if (!p) return empty_ok ? 0 : fail(CAPE_ERR_NULL_PTR, "%s is NULL", name);
    if (reinterpret_cast<uintptr_t>(p) % align != 0)
        return fail(CAPE_ERR_MISALIGNED, "%s (%p) is not %u-byte aligned", name, p, align);
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CAPE_ERR_NOT_DEVICE_PTR, "%s (%p): %s", name, p, cudaGetErrorString(e));
    }
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)
        return fail(CAPE_ERR_NOT_DEVICE_PTR, "%s (%p) is host memory; this library has no CPU path", name, p);
    return 0;
}

}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
const char* const kTuneNames[kTuneCount] = {"FWD_THREADS", "FWD_QPC", "FWD_POINT_MAX_QM", "FWD_STAGED", "FWD_STAGED_MIN_QM",
                                            "FWD_STAGED_KB", "BWD_THREADS", "BWD_QPC", "BWD_MODE", "BWD_STAGED_KB", "PROFILE",
                                            "BWD_TC_MIN_QM", "HOST_CHUNKS", "WGRAD_TRANSPOSE"};
std::atomic<int> g_tune[kTuneCount];
std::once_flag g_tune_once;

void load_tuning_from_env() {
    for (int i = 0; i < kTuneCount; ++i) {
        char name[64];
        snprintf(name, sizeof(name), "CAPE_%s", kTuneNames[i]);
        const char* v = std::getenv(name);
        g_tune[i].store(v && *v ? std::atoi(v) : 0, std::memory_order_relaxed);
    }
}

int tune_index(const char* name) {
    if (!name) return -1;
    if (!std::strncmp(name, "CAPE_", 5)) name += 5;
    for (int i = 0; i < kTuneCount; ++i)
        if (!std::strcmp(name, kTuneNames[i])) return i;
    return -1;
}
}  // namespace

int tuning(Tune knob, int fallback) {
    std::call_once(g_tune_once, load_tuning_from_env);
    const int v = g_tune[knob].load(std::memory_order_relaxed);
    return v > 0 ? v : fallback;
}

bool first_use_on_device(unsigned long long* flags) {
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;   // unknown device: just redo the set-up
    std::lock_guard<std::mutex> lock(mu);
    const unsigned long long bit = 1ull << dev;
    if (*flags & bit) return false;
    *flags |= bit;
    return true;
}

}  // namespace cape

using namespace cape;

extern "C" {

int cape_abi_version(void) { return CAPE_ABI_VERSION; }

const char* cape_last_error(void) { return t_error; }

uint64_t cape_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int cape_set_tuning(const char* name, int value) {
    const int i = tune_index(name);
    if (i < 0) return fail(CAPE_ERR_BAD_DIMS, "unknown tuning knob %s", name ? name : "(null)");
    std::call_once(g_tune_once, load_tuning_from_env);
    g_tune[i].store(value, std::memory_order_relaxed);
    return 0;
}

int cape_debug_counters(long long* out16, int reset) {
    if (!out16) return fail(CAPE_ERR_NULL_PTR, "out16 is NULL");
    const cudaError_t e = read_backward_staged_cycles(out16, reset != 0);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_debug_counters");
}

int cape_get_tuning(const char* name) {
    const int i = tune_index(name);
    if (i < 0) return -1;
    std::call_once(g_tune_once, load_tuning_from_env);
    return g_tune[i].load(std::memory_order_relaxed);
}

int cape_msda_forward(const void* value, const int64_t* spatial_shapes_dev, const int64_t* level_start_index_dev,
                      const void* sampling_locations, const void* attention_weights, void* out,
                      const cape_msda_dims* dims, int value_dtype, int aux_dtype, void* stream) {
    int rc;
    if ((rc = check_dims(dims)) || (rc = check_dtypes(value_dtype, aux_dtype))) return rc;
    const bool empty = dims->N == 0 || dims->Lq == 0;
    if ((rc = check_ptr(value, "value", empty || dims->S == 0)) || (rc = check_ptr(spatial_shapes_dev, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index_dev, "level_start_index", false, 8)) ||
        (rc = check_ptr(sampling_locations, "sampling_locations", empty)) ||
        (rc = check_ptr(attention_weights, "attention_weights", empty)) || (rc = check_ptr(out, "out", empty)))
        return rc;
    if (empty) return 0;
    FwdArgs a{};
    a.value = value;
    a.shapes = spatial_shapes_dev;
    a.starts = level_start_index_dev;
    a.loc = sampling_locations;
    a.attn = attention_weights;
    a.ref_points = nullptr;
    a.out = out;
    a.d = *dims;
    a.value_dtype = value_dtype;
    a.aux_dtype = aux_dtype;
    a.fused = false;
    const cudaError_t e = launch_forward(a, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_forward launch");
}

int cape_msda_backward(const void* grad_out, const void* value, const int64_t* spatial_shapes_dev,
                       const int64_t* level_start_index_dev, const void* sampling_locations,
                       const void* attention_weights, float* grad_value, void* grad_loc, void* grad_attn,
                       const cape_msda_dims* dims, int value_dtype, int aux_dtype, int zero_grad_value, void* stream) {
    int rc;
    if ((rc = check_dims(dims)) || (rc = check_dtypes(value_dtype, aux_dtype))) return rc;
    const bool empty = dims->N == 0 || dims->Lq == 0;
    const bool no_value = dims->N == 0 || dims->S == 0;
    if ((rc = check_ptr(grad_out, "grad_out", empty)) || (rc = check_ptr(value, "value", empty || no_value)) ||
        (rc = check_ptr(spatial_shapes_dev, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index_dev, "level_start_index", false, 8)) ||
        (rc = check_ptr(sampling_locations, "sampling_locations", empty)) ||
        (rc = check_ptr(attention_weights, "attention_weights", empty)) || (rc = check_ptr(grad_value, "grad_value", no_value)) ||
        (rc = check_ptr(grad_loc, "grad_loc", empty)) || (rc = check_ptr(grad_attn, "grad_attn", empty)))
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (zero_grad_value && !no_value) {
        const size_t bytes = static_cast<size_t>(dims->N) * dims->S * dims->M * dims->D * sizeof(float);
        const cudaError_t e = cudaMemsetAsync(grad_value, 0, bytes, s);
        if (e != cudaSuccess) return fail_cuda(e, "cape_msda_backward memset(grad_value)");
    }
    if (empty) return 0;
    BwdArgs a{};
    a.grad_out = grad_out;
    a.value = value;
    a.shapes = spatial_shapes_dev;
    a.starts = level_start_index_dev;
    a.loc = sampling_locations;
    a.attn = attention_weights;
    a.grad_value = grad_value;
    a.grad_loc = grad_loc;
    a.grad_attn = grad_attn;
    a.d = *dims;
    a.value_dtype = value_dtype;
    a.aux_dtype = aux_dtype;
    a.fused = false;
    a.ref_points = nullptr;
    const cudaError_t e = launch_backward(a, s);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_backward launch");
}

int cape_msda_decode(const void* value_cache, const int64_t* spatial_shapes_dev, const int64_t* level_start_index_dev,
                     const float* reference_points, const float* sampling_offsets, const float* attention_logits,
                     void* out, const cape_msda_dims* dims, int value_dtype, void* stream) {
    int rc;
    if ((rc = check_dims(dims)) || (rc = check_dtypes(value_dtype, CAPE_DTYPE_F32))) return rc;
    const bool empty = dims->N == 0 || dims->Lq == 0;
    if ((rc = check_ptr(value_cache, "value_cache", empty || dims->S == 0)) ||
        (rc = check_ptr(spatial_shapes_dev, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index_dev, "level_start_index", false, 8)) ||
        (rc = check_ptr(reference_points, "reference_points", empty, 8)) ||
        (rc = check_ptr(sampling_offsets, "sampling_offsets", empty)) ||
        (rc = check_ptr(attention_logits, "attention_logits", empty)) || (rc = check_ptr(out, "out", empty)))
        return rc;
    if (empty) return 0;
    FwdArgs a{};
    a.value = value_cache;
    a.shapes = spatial_shapes_dev;
    a.starts = level_start_index_dev;
    a.loc = sampling_offsets;
    a.attn = attention_logits;
    a.ref_points = reference_points;
    a.out = out;
    a.d = *dims;
    a.value_dtype = value_dtype;
    a.aux_dtype = CAPE_DTYPE_F32;
    a.fused = true;
    const cudaError_t e = launch_forward(a, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_decode launch");
}

int cape_msda_fused_supported(const cape_msda_dims* dims) {
    return dims && dims->D == 32 && dims->P == 4 && dims->L >= 1 && dims->L <= 4;
}

int cape_msda_fused_backward(const void* grad_out, const void* value, const int64_t* spatial_shapes_dev,
                             const int64_t* level_start_index_dev, const float* reference_points,
                             const float* sampling_offsets, const float* attention_logits, float* grad_value,
                             float* grad_offsets, float* grad_logits, const cape_msda_dims* dims, int value_dtype,
                             int zero_grad_value, void* stream) {
    int rc;
    if ((rc = check_dims(dims)) || (rc = check_dtypes(value_dtype, CAPE_DTYPE_F32))) return rc;
    if (!cape_msda_fused_supported(dims))
        return fail(CAPE_ERR_BAD_DIMS, "the fused prologue needs D=32, P=4, L<=4 (got D=%d P=%d L=%d)", dims->D, dims->P,
                    dims->L);
    const bool empty = dims->N == 0 || dims->Lq == 0;
    const bool no_value = dims->N == 0 || dims->S == 0;
    if ((rc = check_ptr(grad_out, "grad_out", empty)) || (rc = check_ptr(value, "value", empty || no_value)) ||
        (rc = check_ptr(spatial_shapes_dev, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index_dev, "level_start_index", false, 8)) ||
        (rc = check_ptr(reference_points, "reference_points", empty, 8)) ||
        (rc = check_ptr(sampling_offsets, "sampling_offsets", empty)) ||
        (rc = check_ptr(attention_logits, "attention_logits", empty)) || (rc = check_ptr(grad_value, "grad_value", no_value)) ||
        (rc = check_ptr(grad_offsets, "grad_offsets", empty)) || (rc = check_ptr(grad_logits, "grad_logits", empty)))
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (zero_grad_value && !no_value) {
        const size_t bytes = static_cast<size_t>(dims->N) * dims->S * dims->M * dims->D * sizeof(float);
        const cudaError_t e = cudaMemsetAsync(grad_value, 0, bytes, s);
        if (e != cudaSuccess) return fail_cuda(e, "cape_msda_fused_backward memset(grad_value)");
    }
    if (empty) return 0;
    BwdArgs a{};
    a.grad_out = grad_out;
    a.value = value;
    a.shapes = spatial_shapes_dev;
    a.starts = level_start_index_dev;
    a.loc = sampling_offsets;
    a.attn = attention_logits;
    a.ref_points = reference_points;
    a.grad_value = grad_value;
    a.grad_loc = grad_offsets;
    a.grad_attn = grad_logits;
    a.d = *dims;
    a.value_dtype = value_dtype;
    a.aux_dtype = CAPE_DTYPE_F32;
    a.fused = true;
    const cudaError_t e = launch_backward(a, s);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_fused_backward launch");
}

// ---- variant samplers (fp32) ---------------------------------------------------------------------------------------

int cape_msda_query_pool_forward(const float* value, const int64_t* spatial_shapes_dev,
                                 const int64_t* level_start_index_dev, const float* sampling_locations,
                                 const float* attention_weights, float* out, const cape_msda_dims* dims, void* stream) {
    int rc;
    if ((rc = check_dims(dims))) return rc;
    const bool empty = dims->N == 0;
    if ((rc = check_ptr(value, "value", empty || dims->S == 0)) ||
        (rc = check_ptr(spatial_shapes_dev, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index_dev, "level_start_index", false, 8)) ||
        (rc = check_ptr(sampling_locations, "sampling_locations", empty || dims->Lq == 0)) ||
        (rc = check_ptr(attention_weights, "attention_weights", empty || dims->Lq == 0)) ||
        (rc = check_ptr(out, "out", empty)))
        return rc;
    if (empty) return 0;
    FwdArgs a{};
    a.value = value;
    a.shapes = spatial_shapes_dev;
    a.starts = level_start_index_dev;
    a.loc = sampling_locations;
    a.attn = attention_weights;
    a.out = out;
    a.d = *dims;
    const cudaError_t e = launch_query_pool_forward(a, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_query_pool_forward launch");
}

int cape_msda_query_pool_backward(const float* grad_out, const float* value, const int64_t* spatial_shapes_dev,
                                  const int64_t* level_start_index_dev, const float* sampling_locations,
                                  const float* attention_weights, float* grad_value, float* grad_loc, float* grad_attn,
                                  const cape_msda_dims* dims, int zero_grad_value, void* stream) {
    int rc;
    if ((rc = check_dims(dims))) return rc;
    const bool empty = dims->N == 0 || dims->Lq == 0;
    const bool no_value = dims->N == 0 || dims->S == 0;
    if ((rc = check_ptr(grad_out, "grad_out", dims->N == 0)) || (rc = check_ptr(value, "value", empty || no_value)) ||
        (rc = check_ptr(spatial_shapes_dev, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index_dev, "level_start_index", false, 8)) ||
        (rc = check_ptr(sampling_locations, "sampling_locations", empty)) ||
        (rc = check_ptr(attention_weights, "attention_weights", empty)) || (rc = check_ptr(grad_value, "grad_value", no_value)) ||
        (rc = check_ptr(grad_loc, "grad_loc", empty)) || (rc = check_ptr(grad_attn, "grad_attn", empty)))
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (zero_grad_value && !no_value) {
        const size_t bytes = static_cast<size_t>(dims->N) * dims->S * dims->M * dims->D * sizeof(float);
        const cudaError_t e = cudaMemsetAsync(grad_value, 0, bytes, s);
        if (e != cudaSuccess) return fail_cuda(e, "cape_msda_query_pool_backward memset(grad_value)");
    }
    if (empty) return 0;
    BwdArgs a{};
    a.grad_out = grad_out;
    a.value = value;
    a.shapes = spatial_shapes_dev;
    a.starts = level_start_index_dev;
    a.loc = sampling_locations;
    a.attn = attention_weights;
    a.grad_value = grad_value;
    a.grad_loc = grad_loc;
    a.grad_attn = grad_attn;
    a.d = *dims;
    const cudaError_t e = launch_query_pool_backward(a, s);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_query_pool_backward launch");
}

namespace {
int check_points_dims(int B, int G, int c, int H, int W, int Hk, int Wk) {
    if (B < 0 || G <= 0 || c <= 0 || H <= 0 || W <= 0 || Hk < 0 || Wk < 0)
        return fail(CAPE_ERR_BAD_DIMS, "bad point-sampling dimensions (B=%d G=%d c=%d H=%d W=%d Hk=%d Wk=%d)", B, G, c, H, W,
                    Hk, Wk);
    if (static_cast<int64_t>(B) * G * c * H * W > 0x7fffffffffffLL) return fail(CAPE_ERR_BAD_DIMS, "tensor too large");
    return 0;
}
}  // namespace

int cape_points_sample_forward(const float* x, const float* pos, float* out, int B, int G, int c, int H, int W, int Hk,
                               int Wk, void* stream) {
    int rc;
    if ((rc = check_points_dims(B, G, c, H, W, Hk, Wk))) return rc;
    const bool empty = B == 0 || Hk == 0 || Wk == 0;
    if ((rc = check_ptr(x, "x", B == 0, 4)) || (rc = check_ptr(pos, "pos", empty, 8)) || (rc = check_ptr(out, "out", empty, 4)))
        return rc;
    if (empty) return 0;
    const PointsDims p{B, G, c, H, W, Hk, Wk};
    const cudaError_t e = launch_points_sample_forward(x, pos, out, p, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_points_sample_forward launch");
}

int cape_points_sample_backward(const float* grad_out, const float* x, const float* pos, float* grad_x, float* grad_pos,
                                int B, int G, int c, int H, int W, int Hk, int Wk, int zero_grad_x, void* stream) {
    int rc;
    if ((rc = check_points_dims(B, G, c, H, W, Hk, Wk))) return rc;
    const bool empty = B == 0 || Hk == 0 || Wk == 0;
    if ((rc = check_ptr(grad_out, "grad_out", empty, 4)) || (rc = check_ptr(x, "x", B == 0, 4)) ||
        (rc = check_ptr(pos, "pos", empty, 8)) || (rc = check_ptr(grad_x, "grad_x", B == 0, 4)) ||
        (rc = check_ptr(grad_pos, "grad_pos", empty, 8)))
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (zero_grad_x && B > 0) {
        const cudaError_t e = cudaMemsetAsync(grad_x, 0, static_cast<size_t>(B) * G * c * H * W * sizeof(float), s);
        if (e != cudaSuccess) return fail_cuda(e, "cape_points_sample_backward memset(grad_x)");
    }
    if (empty) return 0;
    const PointsDims p{B, G, c, H, W, Hk, Wk};
    const cudaError_t e = launch_points_sample_backward(grad_out, x, pos, grad_x, grad_pos, p, s);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_points_sample_backward launch");
}

int cape_zero_masked_rows(void* value, const uint8_t* mask, int64_t rows, int row_bytes, void* stream) {
    if (rows < 0 || row_bytes < 0 || row_bytes % 16 != 0)
        return fail(CAPE_ERR_BAD_DIMS, "rows=%lld row_bytes=%d: row_bytes must be a non-negative multiple of 16",
                    static_cast<long long>(rows), row_bytes);
    int rc;
    const bool empty = rows == 0 || row_bytes == 0;
    if ((rc = check_ptr(value, "value", empty)) || (rc = check_ptr(mask, "mask", empty, 1))) return rc;
    const cudaError_t e = launch_zero_masked_rows(value, mask, rows, row_bytes, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_zero_masked_rows launch");
}

// ---- sequence side of the decoder ---------------------------------------------------------------------------------

namespace {
int check_seq_embed(const int64_t* const* seqs, const float* const* deltas, int64_t tokens, int C, int V) {
    if (tokens < 0 || C <= 0 || C % 4 != 0 || V <= 0)
        return fail(CAPE_ERR_BAD_DIMS, "bad token-embedding dimensions (tokens=%lld C=%d V=%d; C must be a multiple of 4)",
                    static_cast<long long>(tokens), C, V);
    static const char* const seq_names[4] = {"seq11", "seq12", "seq21", "seq22"};
    static const char* const delta_names[4] = {"delta_x1", "delta_x2", "delta_y1", "delta_y2"};
    int rc;
    for (int k = 0; k < 4; ++k) {
        if ((rc = check_ptr(seqs[k], seq_names[k], tokens == 0, 8))) return rc;
        if ((rc = check_ptr(deltas[k], delta_names[k], tokens == 0, 4))) return rc;
    }
    return 0;
}
}  // namespace

int cape_seq_embed_forward(const float* table, const int64_t* seq11, const int64_t* seq12, const int64_t* seq21,
                           const int64_t* seq22, const float* delta_x1, const float* delta_x2, const float* delta_y1,
                           const float* delta_y2, float* out, int64_t tokens, int C, int V, void* stream) {
    const int64_t* seqs[4] = {seq11, seq12, seq21, seq22};
    const float* deltas[4] = {delta_x1, delta_x2, delta_y1, delta_y2};
    int rc;
    if ((rc = check_seq_embed(seqs, deltas, tokens, C, V))) return rc;
    if ((rc = check_ptr(table, "table", false)) || (rc = check_ptr(out, "out", tokens == 0))) return rc;
    if (tokens == 0) return 0;
    SeqEmbedArgs a{};
    a.table = table;
    a.seq11 = seq11, a.seq12 = seq12, a.seq21 = seq21, a.seq22 = seq22;
    a.dx1 = delta_x1, a.dx2 = delta_x2, a.dy1 = delta_y1, a.dy2 = delta_y2;
    a.out = out;
    a.tokens = tokens, a.C = C, a.V = V;
    const cudaError_t e = launch_seq_embed_forward(a, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_seq_embed_forward launch");
}

int cape_seq_embed_backward(const float* grad_out, const int64_t* seq11, const int64_t* seq12, const int64_t* seq21,
                            const int64_t* seq22, const float* delta_x1, const float* delta_x2, const float* delta_y1,
                            const float* delta_y2, float* grad_table, int64_t tokens, int C, int V, int64_t padding_idx,
                            int zero_grad_table, void* stream) {
    const int64_t* seqs[4] = {seq11, seq12, seq21, seq22};
    const float* deltas[4] = {delta_x1, delta_x2, delta_y1, delta_y2};
    int rc;
    if ((rc = check_seq_embed(seqs, deltas, tokens, C, V))) return rc;
    if ((rc = check_ptr(grad_out, "grad_out", tokens == 0)) || (rc = check_ptr(grad_table, "grad_table", false))) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (zero_grad_table) {
        const cudaError_t e = cudaMemsetAsync(grad_table, 0, static_cast<size_t>(V) * C * sizeof(float), s);
        if (e != cudaSuccess) return fail_cuda(e, "cape_seq_embed_backward memset(grad_table)");
    }
    if (tokens == 0) return 0;
    SeqEmbedArgs a{};
    a.grad_out = grad_out;
    a.seq11 = seq11, a.seq12 = seq12, a.seq21 = seq21, a.seq22 = seq22;
    a.dx1 = delta_x1, a.dx2 = delta_x2, a.dy1 = delta_y1, a.dy2 = delta_y2;
    a.grad_table = grad_table;
    a.tokens = tokens, a.C = C, a.V = V, a.padding_idx = padding_idx;
    const cudaError_t e = launch_seq_embed_backward(a, s);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_seq_embed_backward launch");
}

int cape_token_step(const float* cls_logits, const float* reg, int64_t* step_dev, const cape_token_state* st,
                    const cape_tokenizer* tk, int B, int n_classes, void* stream) {
    if (!st || !tk) return fail(CAPE_ERR_NULL_PTR, "%s is NULL", !st ? "state" : "tokenizer");
    if (B < 0 || n_classes <= 0 || st->max_len <= 0 || tk->num_bins < 2)
        return fail(CAPE_ERR_BAD_DIMS, "bad token-step dimensions (B=%d n_classes=%d max_len=%lld num_bins=%d)", B, n_classes,
                    static_cast<long long>(st->max_len), tk->num_bins);
    int rc;
    const bool empty = B == 0;
    if ((rc = check_ptr(step_dev, "step_dev", false, 8)) || (rc = check_ptr(cls_logits, "cls_logits", empty, 4)) ||
        (rc = check_ptr(reg, "reg", empty, 4)) || (rc = check_ptr(st->unfinished, "state.unfinished", empty, 4)) ||
        (rc = check_ptr(st->finish_step, "state.finish_step", empty, 8)) ||
        (rc = check_ptr(st->seq11, "state.seq11", empty, 8)) || (rc = check_ptr(st->seq12, "state.seq12", empty, 8)) ||
        (rc = check_ptr(st->seq21, "state.seq21", empty, 8)) || (rc = check_ptr(st->seq22, "state.seq22", empty, 8)) ||
        (rc = check_ptr(st->delta_x1, "state.delta_x1", empty, 4)) || (rc = check_ptr(st->delta_x2, "state.delta_x2", empty, 4)) ||
        (rc = check_ptr(st->delta_y1, "state.delta_y1", empty, 4)) || (rc = check_ptr(st->delta_y2, "state.delta_y2", empty, 4)) ||
        (rc = check_ptr(st->pred_logits, "state.pred_logits", empty, 4)) ||
        (rc = check_ptr(st->pred_coords, "state.pred_coords", empty, 4)) ||
        (rc = check_ptr(st->gen_kind, "state.gen_kind", empty, 4)) || (rc = check_ptr(st->gen_xy, "state.gen_xy", empty, 4)))
        return rc;
    const cudaError_t e = launch_token_step(cls_logits, reg, step_dev, *st, *tk, B, n_classes, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_token_step launch");
}

// ---- decode-step kernels ------------------------------------------------------------------------------------------

int cape_decode_attention(const float* q, int q_stride, const float* k_new, const float* v_new, int new_stride,
                          float* k_cache, float* v_cache, const int64_t* pos_dev, const float* key_bias, float* out, int B,
                          int T, int H, int D, void* stream) {
    if (B < 0 || T <= 0 || H <= 0 || D != 32 || T > 1024 || q_stride % 4 != 0 || new_stride % 4 != 0)
        return fail(CAPE_ERR_BAD_DIMS, "bad attention dimensions (B=%d T=%d H=%d D=%d; D must be 32, T <= 1024)", B, T, H, D);
    if ((k_new == nullptr) != (v_new == nullptr) || (pos_dev != nullptr) != (k_new != nullptr))
        return fail(CAPE_ERR_NULL_PTR, "k_new, v_new and pos_dev must be given together (self-attention) or all be NULL");
    int rc;
    const bool empty = B == 0;
    if ((rc = check_ptr(q, "q", empty)) || (rc = check_ptr(k_cache, "k_cache", empty)) ||
        (rc = check_ptr(v_cache, "v_cache", empty)) || (rc = check_ptr(out, "out", empty, 4)) ||
        (rc = check_ptr(k_new, "k_new", true)) || (rc = check_ptr(v_new, "v_new", true)) ||
        (rc = check_ptr(pos_dev, "pos_dev", true, 8)) || (rc = check_ptr(key_bias, "key_bias", true, 4)))
        return rc;
    const cudaError_t e = launch_decode_attention(q, k_new, v_new, k_cache, v_cache, pos_dev, key_bias, out, B, T, H,
                                                  q_stride, new_stride, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_decode_attention launch");
}

int cape_skinny_linear(const float* x, int x_stride, const float* x2, int x2_stride, const float* wt, const float* bias,
                       const float* residual, int residual_stride, const float* gamma, const float* beta, float eps,
                       const float* sine_dim_t, float* y, int y_stride, int rows, int K, int N, int epilogue, void* stream) {
    if (rows < 0 || K <= 0 || N <= 0 || K % 16 != 0 || N % 4 != 0 || K > 2048)
        return fail(CAPE_ERR_BAD_DIMS, "bad linear dimensions (rows=%d K=%d N=%d; K %% 16 == 0, K <= 2048, N %% 4 == 0)", rows, K, N);
    if (epilogue < 0 || epilogue > 2) return fail(CAPE_ERR_BAD_DIMS, "unknown epilogue %d", epilogue);
    if (epilogue == 2 && (N > 256 || !gamma || !beta))
        return fail(CAPE_ERR_BAD_DIMS, "the LayerNorm epilogue needs N <= 256 (got %d) and gamma / beta", N);
    if (sine_dim_t && K != 256) return fail(CAPE_ERR_BAD_DIMS, "the sine-embedding input has K = 256, got %d", K);
    int rc;
    const bool empty = rows == 0;
    if (!sine_dim_t && (x_stride % 4 != 0 || (x2 && x2_stride % 4 != 0)))
        return fail(CAPE_ERR_MISALIGNED, "row strides of x / x2 must be multiples of 4 elements (got %d, %d)", x_stride, x2_stride);
    if ((rc = check_ptr(x, "x", empty, sine_dim_t ? 4 : 16)) || (rc = check_ptr(x2, "x2", true, 16)) ||
        (rc = check_ptr(wt, "wt", false)) || (rc = check_ptr(bias, "bias", true, 4)) ||
        (rc = check_ptr(residual, "residual", true, 4)) || (rc = check_ptr(gamma, "gamma", true, 4)) ||
        (rc = check_ptr(beta, "beta", true, 4)) || (rc = check_ptr(sine_dim_t, "sine_dim_t", true, 4)) ||
        (rc = check_ptr(y, "y", empty, 4)))
        return rc;
    SkinnyArgs a{};
    a.x = x, a.x2 = x2, a.wt = wt, a.bias = bias, a.res = residual, a.gamma = gamma, a.beta = beta;
    a.sine_dim_t = sine_dim_t, a.y = y;
    a.rows = rows, a.K = K, a.N = N;
    a.x_stride = x_stride, a.x2_stride = x2_stride, a.res_stride = residual_stride, a.y_stride = y_stride;
    a.eps = eps;
    const cudaError_t e = launch_skinny_linear(a, epilogue, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_skinny_linear launch");
}

int cape_msda_output_proj(const float* value_cache, const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const float* reference_points, const float* sampling_offsets, const float* attention_logits,
                          const cape_msda_dims* dims, const float* wt, const float* bias, const float* residual,
                          int residual_stride, const float* gamma, const float* beta, float eps, float* y, int y_stride,
                          int n_out, void* stream) {
    int rc;
    if ((rc = check_dims(dims))) return rc;
    const cape_msda_dims& d = *dims;
    if (d.D != 32 || d.P != 4 || d.L != 4 || d.M < 1 || d.M * 32 > 2048 || (d.M * 32) % 16 != 0)
        return fail(CAPE_ERR_BAD_DIMS, "cape_msda_output_proj needs D = 32, P = 4, L = 4 (got D=%d P=%d L=%d M=%d)", d.D, d.P, d.L, d.M);
    if (n_out <= 0 || n_out > 256 || n_out % 4 != 0 || !gamma || !beta)
        return fail(CAPE_ERR_BAD_DIMS, "the LayerNorm epilogue needs 0 < N <= 256, N %% 4 == 0 (got %d) and gamma / beta", n_out);
    const long long rows64 = static_cast<long long>(d.N) * d.Lq;
    if (rows64 > 0x7fffffffLL) return fail(CAPE_ERR_BAD_DIMS, "too many rows");
    const int rows = static_cast<int>(rows64);
    const bool empty = rows == 0;
    if ((rc = check_ptr(value_cache, "value_cache", empty)) || (rc = check_ptr(spatial_shapes, "spatial_shapes", false, 8)) ||
        (rc = check_ptr(level_start_index, "level_start_index", false, 8)) ||
        (rc = check_ptr(reference_points, "reference_points", empty, 4)) ||
        (rc = check_ptr(sampling_offsets, "sampling_offsets", empty, 4)) ||
        (rc = check_ptr(attention_logits, "attention_logits", empty, 4)) || (rc = check_ptr(wt, "wt", false)) ||
        (rc = check_ptr(bias, "bias", true, 4)) || (rc = check_ptr(residual, "residual", true, 4)) ||
        (rc = check_ptr(gamma, "gamma", false, 4)) || (rc = check_ptr(beta, "beta", false, 4)) || (rc = check_ptr(y, "y", empty, 4)))
        return rc;
    SkinnyArgs a{};
    a.wt = wt, a.bias = bias, a.res = residual, a.gamma = gamma, a.beta = beta, a.y = y;
    a.rows = rows, a.K = d.M * 32, a.N = n_out;
    a.res_stride = residual_stride, a.y_stride = y_stride;
    a.eps = eps;
    a.msda_value = value_cache, a.msda_shapes = spatial_shapes, a.msda_starts = level_start_index;
    a.msda_ref = reference_points, a.msda_off = sampling_offsets, a.msda_logits = attention_logits;
    a.msda_S = d.S, a.msda_M = d.M, a.msda_Lq = d.Lq;
    const cudaError_t e = launch_skinny_linear(a, 2, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_msda_output_proj launch");
}

int cape_skinny_linear_split(const float* x, int x_stride, const float* x2, int x2_stride, const float* wt,
                             const float* bias, float* y, int y_stride, float* y2, int y2_stride, int split, int rows, int K,
                             int N, void* stream) {
    if (rows < 0 || K <= 0 || N <= 0 || K % 16 != 0 || N % 4 != 0 || K > 2048 || split <= 0 || split >= N)
        return fail(CAPE_ERR_BAD_DIMS, "bad split-linear dimensions (rows=%d K=%d N=%d split=%d)", rows, K, N, split);
    int rc;
    const bool empty = rows == 0;
    if (x_stride % 4 != 0 || (x2 && x2_stride % 4 != 0))
        return fail(CAPE_ERR_MISALIGNED, "row strides of x / x2 must be multiples of 4 elements (got %d, %d)", x_stride, x2_stride);
    if ((rc = check_ptr(x, "x", empty)) || (rc = check_ptr(x2, "x2", true, 16)) || (rc = check_ptr(wt, "wt", false)) ||
        (rc = check_ptr(bias, "bias", true, 4)) || (rc = check_ptr(y, "y", empty, 4)) || (rc = check_ptr(y2, "y2", empty, 4)))
        return rc;
    SkinnyArgs a{};
    a.x = x, a.x2 = x2, a.wt = wt, a.bias = bias, a.y = y, a.y2 = y2;
    a.rows = rows, a.K = K, a.N = N;
    a.x_stride = x_stride, a.x2_stride = x2_stride, a.y_stride = y_stride, a.y2_stride = y2_stride, a.split = split;
    const cudaError_t e = launch_skinny_linear(a, 0, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_skinny_linear_split launch");
}

int cape_coord_head_refine(const float* x, int x_stride, const float* wt, const float* bias, const float* w3, const float* b3,
                           const float* ref_in, const float* valid_ratios, float* ref_out, float* ref_levels, int rows,
                           int K, int N, int n_levels, void* stream) {
    if (rows < 0 || K <= 0 || N <= 0 || K % 16 != 0 || N % 4 != 0 || K > 2048 || N > 256 || n_levels <= 0 || n_levels > 8)
        return fail(CAPE_ERR_BAD_DIMS, "bad coordinate-head dimensions (rows=%d K=%d N=%d levels=%d; N <= 256)", rows, K, N,
                    n_levels);
    int rc;
    const bool empty = rows == 0;
    if (x_stride % 4 != 0) return fail(CAPE_ERR_MISALIGNED, "row stride of x must be a multiple of 4 elements (got %d)", x_stride);
    if ((rc = check_ptr(x, "x", empty)) || (rc = check_ptr(wt, "wt", false)) || (rc = check_ptr(bias, "bias", true, 4)) ||
        (rc = check_ptr(w3, "w3", false, 4)) || (rc = check_ptr(b3, "b3", false, 4)) ||
        (rc = check_ptr(ref_in, "ref_in", empty, 4)) || (rc = check_ptr(valid_ratios, "valid_ratios", empty, 4)) ||
        (rc = check_ptr(ref_out, "ref_out", empty, 4)) || (rc = check_ptr(ref_levels, "ref_levels", empty, 4)))
        return rc;
    SkinnyArgs a{};
    a.x = x, a.wt = wt, a.bias = bias;
    a.rows = rows, a.K = K, a.N = N, a.x_stride = x_stride;
    a.w3 = w3, a.b3 = b3, a.ref_in = ref_in, a.valid_ratios = valid_ratios, a.ref_out = ref_out, a.ref_levels = ref_levels;
    a.n_levels = n_levels;
    const cudaError_t e = launch_skinny_linear(a, 3, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_coord_head_refine launch");
}

int cape_tiny_linear(const float* x, int x_stride, const float* w, const float* bias, const float* refine_ref, float* y,
                     int rows, int K, int N, void* stream) {
    if (rows < 0 || K <= 0 || K % 4 != 0 || N <= 0 || N > 8 || x_stride % 4 != 0)
        return fail(CAPE_ERR_BAD_DIMS, "bad tiny-linear dimensions (rows=%d K=%d N=%d stride=%d; N <= 8, K %% 4 == 0)", rows, K, N,
                    x_stride);
    int rc;
    const bool empty = rows == 0;
    if ((rc = check_ptr(x, "x", empty)) || (rc = check_ptr(w, "w", false)) || (rc = check_ptr(bias, "bias", true, 4)) ||
        (rc = check_ptr(refine_ref, "refine_ref", true, 4)) || (rc = check_ptr(y, "y", empty, 4)))
        return rc;
    const cudaError_t e = launch_tiny_linear(x, x_stride, w, bias, refine_ref, y, rows, K, N, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_tiny_linear launch");
}

// ---- fp32-accurate tensor-core linear ----------------------------------------------------------------------------

int cape_tf32_split_lo(const float* x, float* lo, int64_t n, void* stream) {
    if (n < 0) return fail(CAPE_ERR_BAD_DIMS, "negative element count");
    int rc;
    if ((rc = check_ptr(x, "x", n == 0, 4)) || (rc = check_ptr(lo, "lo", n == 0, 4))) return rc;
    const cudaError_t e = launch_tf32_split_lo(x, lo, n, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_tf32_split_lo launch");
}

int cape_linear_tf32x3(const float* x, const float* w, const float* w_lo, const float* bias, float* y, int M, int N, int K,
                       int act, void* stream) {
    if (M < 0 || N <= 0 || K <= 0 || K % 32 != 0 || N % 128 != 0 || act < 0 || act > 1)
        return fail(CAPE_ERR_BAD_DIMS, "bad linear dimensions (M=%d N=%d K=%d act=%d; K %% 32 == 0, N %% 128 == 0)", M, N, K, act);
    int rc;
    const bool empty = M == 0;
    if ((rc = check_ptr(x, "x", empty)) || (rc = check_ptr(w, "w", false)) || (rc = check_ptr(w_lo, "w_lo", false)) ||
        (rc = check_ptr(bias, "bias", true)) || (rc = check_ptr(y, "y", empty)))
        return rc;
    const cudaError_t e = launch_linear_tf32x3(x, w, w_lo, bias, y, M, N, K, act, 0, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_linear_tf32x3 launch");
}

int cape_linear_tf32x3_wgrad(const float* grad_out, const float* x, float* grad_w, float* workspace, int rows, int N, int K,
                             void* stream) {
    const bool transposed = cape::tuning(cape::kTuneWgradTranspose, 0) == 1;
    if (rows <= 0 || N <= 0 || K <= 0 || N % 32 != 0 || K % 128 != 0 || (transposed && rows % 32 != 0))
        return fail(CAPE_ERR_BAD_DIMS, "bad weight-gradient dimensions (rows=%d N=%d K=%d; N %% 32 == 0, K %% 128 == 0)", rows, N, K);
    int rc;
    if ((rc = check_ptr(grad_out, "grad_out", false)) || (rc = check_ptr(x, "x", false)) || (rc = check_ptr(grad_w, "grad_w", false)) ||
        (rc = check_ptr(workspace, "workspace", !transposed)))
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!transposed) {   // default: operands read in place (MN-major tiles), no transposed copies
        const cudaError_t e0 = launch_wgrad_tf32x3(grad_out, x, grad_w, rows, N, K, s);
        return e0 == cudaSuccess ? 0 : fail_cuda(e0, "cape_linear_tf32x3_wgrad launch");
    }
    // WGRAD_TRANSPOSE=1 (the earlier formulation, kept for A/B runs): workspace = g^T (N, rows) | x^T (K, rows) | lo(x^T) (K, rows)
    float* gt = workspace;
    float* xt = gt + static_cast<size_t>(N) * rows;
    float* xt_lo = xt + static_cast<size_t>(K) * rows;
    cudaError_t e;
    if ((e = launch_transpose_lo(grad_out, gt, nullptr, rows, N, s)) != cudaSuccess) return fail_cuda(e, "transpose(grad_out)");
    if ((e = launch_transpose_lo(x, xt, xt_lo, rows, K, s)) != cudaSuccess) return fail_cuda(e, "transpose(x)");
    // grad_w (N, K) = g^T (N, rows) . (x^T (K, rows))^T : the same K-major GEMM with the row index as the reduction dimension
    e = launch_linear_tf32x3(gt, xt, xt_lo, nullptr, grad_w, N, K, rows, 0, 1, s);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cape_linear_tf32x3_wgrad launch");
}

// ---- host-buffer round trip ------------------------------------------------------------------------------------

namespace {
struct HostPlan {
    size_t value, loc, attn, out, gout, gvalue, gloc, gattn, shapes, starts, total;
};
size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
HostPlan plan_host(const cape_msda_dims& d, bool bwd) {
    HostPlan p{};
    const size_t nv = static_cast<size_t>(d.N) * d.S * d.M * d.D * 4;
    const size_t no = static_cast<size_t>(d.N) * d.Lq * d.M * d.D * 4;
    const size_t na = static_cast<size_t>(d.N) * d.Lq * d.M * d.L * d.P * 4;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += align256(bytes);
        return at;
    };
    p.shapes = take(static_cast<size_t>(d.L) * 2 * 8);
    p.starts = take(static_cast<size_t>(d.L) * 8);
    p.value = take(nv);
    p.loc = take(na * 2);
    p.attn = take(na);
    p.out = take(no);
    if (bwd) {
        p.gout = take(no);
        p.gvalue = take(nv);
        p.gloc = take(na * 2);
        p.gattn = take(na);
    }
    p.total = off;
    return p;
}
}  // namespace

namespace {
// Two non-blocking copy streams per device, created on first use and kept for the life of the process.
struct HelperStreams {
    cudaStream_t in, out;
};
bool get_helper_streams(HelperStreams* hs) {
    static std::mutex mu;
    static HelperStreams per_device[64] = {};
    static bool made[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    std::lock_guard<std::mutex> lock(mu);
    if (!made[dev]) {
        if (cudaStreamCreateWithFlags(&per_device[dev].in, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&per_device[dev].out, cudaStreamNonBlocking) != cudaSuccess) return false;
        made[dev] = true;
    }
    *hs = per_device[dev];
    return true;
}
}  // namespace

size_t cape_msda_host_workspace_bytes(const cape_msda_dims* dims, int with_backward) {
    if (check_dims(dims)) return 0;
    return plan_host(*dims, with_backward != 0).total;
}

int cape_msda_forward_backward_host(const float* value_host, const int64_t* spatial_shapes_host,
                                    const int64_t* level_start_index_host, const float* loc_host,
                                    const float* attn_host, const float* grad_out_host, float* out_host,
                                    float* grad_value_host, float* grad_loc_host, float* grad_attn_host,
                                    const cape_msda_dims* dims, void* workspace_dev, size_t workspace_bytes,
                                    void* stream) {
    int rc;
    if ((rc = check_dims(dims))) return rc;
    const bool bwd = grad_out_host != nullptr;
    if (!value_host || !spatial_shapes_host || !level_start_index_host || !loc_host || !attn_host || !out_host)
        return fail(CAPE_ERR_NULL_PTR, "a required host pointer is NULL");
    if (bwd && (!grad_value_host || !grad_loc_host || !grad_attn_host))
        return fail(CAPE_ERR_NULL_PTR, "grad_out_host given but a gradient output pointer is NULL");
    const HostPlan p = plan_host(*dims, bwd);
    if ((rc = check_ptr(workspace_dev, "workspace_dev", false))) return rc;
    if (workspace_bytes < p.total)
        return fail(CAPE_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, p.total);
    const cape_msda_dims& d = *dims;
    // per batch element sizes (all tensors are N-major, so a chunk of the batch is a contiguous slice of each)
    const size_t ev = static_cast<size_t>(d.S) * d.M * d.D * 4;
    const size_t eo = static_cast<size_t>(d.Lq) * d.M * d.D * 4;
    const size_t ea = static_cast<size_t>(d.Lq) * d.M * d.L * d.P * 4;
    char* ws = static_cast<char*>(workspace_dev);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
#define CAPE_TRY(call, what) \
    if ((e = (call)) != cudaSuccess) return fail_cuda(e, what);
    CAPE_TRY(cudaMemcpyAsync(ws + p.shapes, spatial_shapes_host, static_cast<size_t>(d.L) * 16, cudaMemcpyHostToDevice, s),
             "copy spatial_shapes")
    CAPE_TRY(cudaMemcpyAsync(ws + p.starts, level_start_index_host, static_cast<size_t>(d.L) * 8, cudaMemcpyHostToDevice, s),
             "copy level_start_index")
    if (d.N == 0) return 0;

    // Software pipeline over chunks of the batch: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of chunk c
    // (PCIe is full duplex).  Copies ride on two library-owned helper streams forked from / joined to `stream` with
    // events; kernels stay on the caller's stream.
    const size_t bytes_in = static_cast<size_t>(d.N) * (ev + 3 * ea + (bwd ? eo : 0));
    constexpr int kMaxChunks = 32;
    int want = cape::tuning(cape::kTuneHostChunks, kMaxChunks);   // N = 20: 20 chunks of one image 9.67 ms, 8 chunks 10.23 ms (tools/host_chunks_time.py)
    if (want > kMaxChunks) want = kMaxChunks;
    int chunks = bytes_in > (static_cast<size_t>(32) << 20) ? (d.N < want ? d.N : want) : 1;
    HelperStreams hs;
    if (chunks > 1 && !get_helper_streams(&hs)) chunks = 1;
    // Events are owned by a guard so that every exit path (including the CAPE_TRY early returns) destroys them and joins
    // the helper streams back into the caller's stream.
    struct EventGuard {
        cudaEvent_t ev[2 + 2 * kMaxChunks] = {};
        int n = 0;
        cudaStream_t caller = nullptr, in = nullptr, out = nullptr;
        bool forked = false;
        cudaError_t make(cudaEvent_t* e) {
            const cudaError_t rc_ = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
            if (rc_ == cudaSuccess) ev[n++] = *e;
            return rc_;
        }
        ~EventGuard() {
            if (forked && n > 0) {               // error path: whatever the helper streams got must finish before the caller's stream goes on
                if (cudaEventRecord(ev[0], in) == cudaSuccess) cudaStreamWaitEvent(caller, ev[0], 0);
                if (n > 1 && cudaEventRecord(ev[1], out) == cudaSuccess) cudaStreamWaitEvent(caller, ev[1], 0);
            }
            for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
        }
    } guard;
    cudaEvent_t fork = nullptr, in_done[kMaxChunks] = {}, k_done[kMaxChunks] = {}, out_done = nullptr;
    if (chunks > 1) {
        guard.caller = s, guard.in = hs.in, guard.out = hs.out;
        CAPE_TRY(guard.make(&fork), "event")
        CAPE_TRY(guard.make(&out_done), "event")
        for (int c = 0; c < chunks; ++c) {
            CAPE_TRY(guard.make(&in_done[c]), "event")
            CAPE_TRY(guard.make(&k_done[c]), "event")
        }
        CAPE_TRY(cudaEventRecord(fork, s), "fork")
        CAPE_TRY(cudaStreamWaitEvent(hs.in, fork, 0), "fork")
        CAPE_TRY(cudaStreamWaitEvent(hs.out, fork, 0), "fork")
        guard.forked = true;
    }
    cudaStream_t s_in = chunks > 1 ? hs.in : s, s_out = chunks > 1 ? hs.out : s;
    const int per = (d.N + chunks - 1) / chunks;
    for (int c = 0; c < chunks; ++c) {
        const int n0 = c * per, nc = (n0 + per <= d.N ? per : d.N - n0);
        if (nc <= 0) break;
        const size_t o = static_cast<size_t>(n0);
        CAPE_TRY(cudaMemcpyAsync(ws + p.value + o * ev, reinterpret_cast<const char*>(value_host) + o * ev, nc * ev,
                                 cudaMemcpyHostToDevice, s_in), "H2D value")
        CAPE_TRY(cudaMemcpyAsync(ws + p.loc + o * 2 * ea, reinterpret_cast<const char*>(loc_host) + o * 2 * ea,
                                 nc * 2 * ea, cudaMemcpyHostToDevice, s_in), "H2D sampling_locations")
        CAPE_TRY(cudaMemcpyAsync(ws + p.attn + o * ea, reinterpret_cast<const char*>(attn_host) + o * ea, nc * ea,
                                 cudaMemcpyHostToDevice, s_in), "H2D attention_weights")
        if (bwd)
            CAPE_TRY(cudaMemcpyAsync(ws + p.gout + o * eo, reinterpret_cast<const char*>(grad_out_host) + o * eo, nc * eo,
                                     cudaMemcpyHostToDevice, s_in), "H2D grad_out")
        if (chunks > 1) {
            CAPE_TRY(cudaEventRecord(in_done[c], s_in), "record")
            CAPE_TRY(cudaStreamWaitEvent(s, in_done[c], 0), "wait")
        }
        cape_msda_dims dc = d;
        dc.N = nc;
        rc = cape_msda_forward(ws + p.value + o * ev, reinterpret_cast<const int64_t*>(ws + p.shapes),
                               reinterpret_cast<const int64_t*>(ws + p.starts), ws + p.loc + o * 2 * ea,
                               ws + p.attn + o * ea, ws + p.out + o * eo, &dc, CAPE_DTYPE_F32, CAPE_DTYPE_F32, stream);
        if (rc) return rc;
        if (bwd) {
            rc = cape_msda_backward(ws + p.gout + o * eo, ws + p.value + o * ev,
                                    reinterpret_cast<const int64_t*>(ws + p.shapes),
                                    reinterpret_cast<const int64_t*>(ws + p.starts), ws + p.loc + o * 2 * ea,
                                    ws + p.attn + o * ea, reinterpret_cast<float*>(ws + p.gvalue + o * ev),
                                    ws + p.gloc + o * 2 * ea, ws + p.gattn + o * ea, &dc, CAPE_DTYPE_F32, CAPE_DTYPE_F32,
                                    /*zero_grad_value=*/1, stream);
            if (rc) return rc;
        }
        if (chunks > 1) {
            CAPE_TRY(cudaEventRecord(k_done[c], s), "record")
            CAPE_TRY(cudaStreamWaitEvent(s_out, k_done[c], 0), "wait")
        }
        CAPE_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(out_host) + o * eo, ws + p.out + o * eo, nc * eo,
                                 cudaMemcpyDeviceToHost, s_out), "D2H out")
        if (bwd) {
            CAPE_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(grad_value_host) + o * ev, ws + p.gvalue + o * ev, nc * ev,
                                     cudaMemcpyDeviceToHost, s_out), "D2H grad_value")
            CAPE_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(grad_loc_host) + o * 2 * ea, ws + p.gloc + o * 2 * ea,
                                     nc * 2 * ea, cudaMemcpyDeviceToHost, s_out), "D2H grad_loc")
            CAPE_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(grad_attn_host) + o * ea, ws + p.gattn + o * ea, nc * ea,
                                     cudaMemcpyDeviceToHost, s_out), "D2H grad_attn")
        }
    }
    if (chunks > 1) {   // join: the caller's stream completes only after the last D2H and the last H2D
        CAPE_TRY(cudaEventRecord(out_done, s_out), "join")
        CAPE_TRY(cudaStreamWaitEvent(s, out_done, 0), "join")
        guard.forked = false;                    // joined: the guard only destroys the events
    }
#undef CAPE_TRY
    return 0;
}

}  // extern "C"
