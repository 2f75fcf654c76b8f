// MSDeformAttn forward with the coarse pyramid levels of one (batch element, head) staged in shared memory by the TMA.
//
// Same arithmetic and lane mapping as msda_fwd_fast_kernel (msda_forward.cu; replaces ms_deform_attn_core_pytorch,
// /root/reference/models/deformable_transformer.py:115-141): a warp takes four consecutive queries, one per 8-lane group,
// lane = (query slot, channel quad).  What changes is where the corner rows come from:
//
//   * the value rows of head m of image n are a strided 2-D tensor (pixel stride M*D elements, D = 32 contiguous
//     channels).  All levels from `l0` upwards — as many as fit the shared-memory budget, for the CAPE pyramid levels
//     1..3 = 1344 rows = 168 KB in fp32 — are copied into shared memory with cp.async.bulk.tensor (boxes of 64 pixels x
//     32 channels, completion on an mbarrier), so their corners are LDS.128 reads: one 128 B row is one conflict-free
//     quarter-warp phase, against ~1.7 clk per row for an L1-resident LDG.128 gather (4 rows per instruction replay in
//     the L1 tag stage).  Level 0 (512 KB per head) stays on the LDG path and has the whole L1 to itself;
//   * persistent grid, one CTA of 32 warps per SM: the (image, head, query) space is cut into equal contiguous ranges,
//     so a CTA refills its shared memory at most three times per launch (~2 % of its time) and there is no tail wave.
//
// Out-of-bounds corners are predicated loads into zeroed registers, as in the L1 kernel.  A level whose
// (start, H, W) does not fit inside S is skipped (contributes zeros) instead of reading out of bounds.
#include "async_copy.cuh"
#include "msda_common.cuh"
#include "msda_launch.h"

namespace cape {

namespace {

constexpr int kStThreads = 1024;
constexpr int kBoxRows = 64;          // pixels per TMA box
constexpr int kMaxDynSmem = 227 * 1024 - 2048;   // opt-in limit minus this kernel's static shared memory (barrier + padding)

// 4 channels from shared memory (zeros when !pred); `addr` is a shared-state-space byte address.
template <typename VT>
__device__ __forceinline__ float4 lds4_or_zero(uint32_t addr, bool pred);
template <>
__device__ __forceinline__ float4 lds4_or_zero<float>(uint32_t addr, bool pred) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "r"(addr), "r"(static_cast<int>(pred)));
    return v;
}
__device__ __forceinline__ uint2 lds2u_or_zero(uint32_t addr, bool pred) {
    uint2 r = make_uint2(0u, 0u);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.shared.v2.b32 {%0, %1}, [%2];\n\t}"
        : "+r"(r.x), "+r"(r.y)
        : "r"(addr), "r"(static_cast<int>(pred)));
    return r;
}
template <>
__device__ __forceinline__ float4 lds4_or_zero<__nv_bfloat16>(uint32_t addr, bool pred) {
    const uint2 r = lds2u_or_zero(addr, pred);
    float4 f;
    f.x = __uint_as_float(r.x << 16);
    f.y = __uint_as_float(r.x & 0xffff0000u);
    f.z = __uint_as_float(r.y << 16);
    f.w = __uint_as_float(r.y & 0xffff0000u);
    return f;
}
template <>
__device__ __forceinline__ float4 lds4_or_zero<__half>(uint32_t addr, bool pred) {
    const uint2 r = lds2u_or_zero(addr, pred);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

struct RawSamples4 {
    float4 loc;
    float2 attn;
};

__device__ __forceinline__ void load_raw4(const float* locp, const float* attnp, int64_t qm, int k, bool on, RawSamples4& r,
                                          float pad) {
    r.loc = make_float4(0.f, 0.f, 0.f, 0.f);
    r.attn = make_float2(pad, pad);
    if (on) {
        r.loc = __ldg(reinterpret_cast<const float4*>(locp + qm * 32) + k);
        r.attn = __ldg(reinterpret_cast<const float2*>(attnp + qm * 16) + k);
    }
}
template <typename HT>
__device__ __forceinline__ void load_raw4(const HT* locp, const HT* attnp, int64_t qm, int k, bool on, RawSamples4& r,
                                          float pad) {
    r.loc = make_float4(0.f, 0.f, 0.f, 0.f);
    r.attn = make_float2(pad, pad);
    if (on) {
        r.loc = ld4(locp + qm * 32 + k * 4);
        const HT* a = attnp + qm * 16 + k * 2;
        r.attn = make_float2(to_f32(a[0]), to_f32(a[1]));
    }
}

// L = 4 levels, P = 4 points, D = 32 channels (the CAPE configuration); M is a run-time value.
template <typename VT, typename AT, bool FUSED>
__global__ void __launch_bounds__(kStThreads, 1)
msda_fwd_staged_kernel(const __grid_constant__ CUtensorMap vmap, const VT* __restrict__ value,
                       const int64_t* __restrict__ shapes, const int64_t* __restrict__ starts,
                       const void* __restrict__ locp, const void* __restrict__ attnp, const float* __restrict__ refp,
                       VT* __restrict__ out, int N, int S, int M, int Lq, int cap_rows, long long per_cta) {
    constexpr int L = 4, D = 32;
    constexpr int kRowB = D * static_cast<int>(sizeof(VT));
    using LT = typename std::conditional<FUSED, float, AT>::type;
    extern __shared__ __align__(128) uint8_t staged[];
    __shared__ __align__(8) uint64_t bar_storage;
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
    const int g = lane >> 3, k = lane & 7;
    const int rowStride = M * D;

    // level table; a level that does not fit inside S contributes nothing
    int H[L], W[L], st[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        H[l] = static_cast<int>(__ldg(shapes + 2 * l));
        W[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
        const long long s0 = __ldg(starts + l);
        st[l] = static_cast<int>(s0);
        if (s0 < 0 || H[l] < 0 || W[l] < 0 || s0 + static_cast<long long>(H[l]) * W[l] > S) H[l] = W[l] = st[l] = 0;
    }
    // rows [base_row, S) are staged: the longest suffix of levels (in start order) that fits cap_rows
    int base_row = S;
    bool suffix = true;
#pragma unroll
    for (int l = L - 1; l >= 0; --l) {
        suffix = suffix && H[l] > 0 && st[l] < base_row && S - st[l] <= cap_rows;
        if (suffix) base_row = st[l];
    }
    const int rows = S - base_row;
    const int nboxes = (rows + kBoxRows - 1) / kBoxRows;
    bool in_smem[L];
    uint32_t lvl_off[L];      // staged: byte offset of the level's first row in shared memory; else element offset in the image
#pragma unroll
    for (int l = 0; l < L; ++l) {
        in_smem[l] = rows > 0 && H[l] > 0 && st[l] >= base_row;
        lvl_off[l] = in_smem[l] ? static_cast<uint32_t>(st[l] - base_row) * kRowB : static_cast<uint32_t>(st[l]) * rowStride;
    }
    const uint32_t smem0 = smem_addr_u32(staged) + k * (kRowB / 8);
    const uint32_t bar = smem_addr_u32(&bar_storage);
    if (tid == 0) {
        mbarrier_init(bar, 1);
        mbarrier_init_fence();
    }
    __syncthreads();

    float ownW = 1.f, ownH = 1.f;    // dimensions of the level whose samples this lane converts (level k >> 1)
#pragma unroll
    for (int l = 0; l < L; ++l)
        if ((k >> 1) == l) {
            ownW = static_cast<float>(W[l]);
            ownH = static_cast<float>(H[l]);
        }
    const LT* loc_t = static_cast<const LT*>(locp);
    const LT* attn_t = static_cast<const LT*>(attnp);
    const float pad = FUSED ? -INFINITY : 0.f;
    const int grp = lane & 24;
    constexpr int nwarps = kStThreads / 32;

    const long long total = static_cast<long long>(N) * M * Lq;
    long long pos = static_cast<long long>(blockIdx.x) * per_cta;
    const long long end = min(total, pos + per_cta);
    uint32_t phase = 0;
    while (pos < end) {
        const int nm = static_cast<int>(pos / Lq);
        const int q_begin = static_cast<int>(pos - static_cast<long long>(nm) * Lq);
        const int q_end = static_cast<int>(min(static_cast<long long>(Lq), q_begin + (end - pos)));
        const int n = nm / M, m = nm - n * M;
        if (rows > 0) {
            __syncthreads();                 // every warp is done with the previous image's rows
            if (warp == 0) {
                if (lane == 0) mbarrier_arrive_expect_tx(bar, static_cast<uint32_t>(nboxes) * kBoxRows * kRowB);
                __syncwarp();
                for (int b = lane; b < nboxes; b += 32)
                    tma_load_box_2d(smem_addr_u32(staged) + b * kBoxRows * kRowB, &vmap, bar, m * D,
                                    n * S + base_row + b * kBoxRows);
            }
            mbarrier_wait(bar, phase);
            phase ^= 1;
        }
        const VT* gimg = value + (static_cast<int64_t>(n) * S * M + m) * D + k * 4;
        int qw = q_begin + warp * 4;
        RawSamples4 cur, nxt;
        load_raw4(loc_t, attn_t, (static_cast<int64_t>(n) * Lq + qw + g) * M + m, k, qw + g < q_end, nxt, pad);
        for (; qw < q_end; qw += nwarps * 4) {
            const int q = qw + g;
            const bool on = q < q_end;
            const int64_t nq = static_cast<int64_t>(n) * Lq + q;
            cur = nxt;
            load_raw4(loc_t, attn_t, (nq + nwarps * 4) * M + m, k, q + nwarps * 4 < q_end, nxt, pad);
            if (FUSED) {   // softmax over the group's 16 logits; loc = ref + off / (W_l, H_l)  (deformable_transformer.py:100-105)
                float mx = fmaxf(cur.attn.x, cur.attn.y);
#pragma unroll
                for (int s = 4; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, s));
                const float e0 = on ? expf(cur.attn.x - mx) : 0.f, e1 = on ? expf(cur.attn.y - mx) : 0.f;
                float sum = e0 + e1;
#pragma unroll
                for (int s = 4; s >= 1; s >>= 1) sum += __shfl_xor_sync(kFullMask, sum, s);
                cur.attn = make_float2(e0 / sum, e1 / sum);
                if (on) {
                    const float2 r = __ldg(reinterpret_cast<const float2*>(refp + nq * (L * 2)) + (k >> 1));
                    cur.loc.x = r.x + cur.loc.x / ownW;
                    cur.loc.y = r.y + cur.loc.y / ownH;
                    cur.loc.z = r.x + cur.loc.z / ownW;
                    cur.loc.w = r.y + cur.loc.w / ownH;
                }
            }
            const float px0 = on ? pixel_coord(cur.loc.x, ownW) : -4.f, py0 = on ? pixel_coord(cur.loc.y, ownH) : -4.f;
            const float px1 = on ? pixel_coord(cur.loc.z, ownW) : -4.f, py1 = on ? pixel_coord(cur.loc.w, ownH) : -4.f;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int l = 0; l < L; ++l) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int s = l * 4 + p, src = grp | (s >> 1);
                    const float px = __shfl_sync(kFullMask, (s & 1) ? px1 : px0, src);
                    const float py = __shfl_sync(kFullMask, (s & 1) ? py1 : py0, src);
                    const float a = __shfl_sync(kFullMask, (s & 1) ? cur.attn.y : cur.attn.x, src);
                    const float xf = floorf(px), yf = floorf(py);
                    const float lx = px - xf, ly = py - yf;
                    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
                    const bool x0ok = static_cast<unsigned>(x0) < static_cast<unsigned>(W[l]);
                    const bool x1ok = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(W[l]);
                    const bool y0ok = static_cast<unsigned>(y0) < static_cast<unsigned>(H[l]);
                    const bool y1ok = static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(H[l]);
                    const int r00 = y0 * W[l] + x0;
                    float4 v00, v01, v10, v11;
                    if (in_smem[l]) {        // warp-uniform
                        const uint32_t a00 = smem0 + lvl_off[l] + static_cast<uint32_t>(r00 * kRowB);
                        const uint32_t a10 = a00 + static_cast<uint32_t>(W[l] * kRowB);
                        v00 = lds4_or_zero<VT>(a00, y0ok & x0ok);
                        v01 = lds4_or_zero<VT>(a00 + kRowB, y0ok & x1ok);
                        v10 = lds4_or_zero<VT>(a10, y1ok & x0ok);
                        v11 = lds4_or_zero<VT>(a10 + kRowB, y1ok & x1ok);
                    } else {
                        const VT* p00 = gimg + lvl_off[l] + r00 * rowStride;
                        const VT* p10 = p00 + W[l] * rowStride;
                        v00 = ld4_or_zero(p00, y0ok & x0ok);
                        v01 = ld4_or_zero(p00 + rowStride, y0ok & x1ok);
                        v10 = ld4_or_zero(p10, y1ok & x0ok);
                        v11 = ld4_or_zero(p10 + rowStride, y1ok & x1ok);
                    }
                    const float ahy = a * (1.f - ly), aly = a * ly, hx = 1.f - lx;
                    fma4(ahy * hx, v00, acc);
                    fma4(ahy * lx, v01, acc);
                    fma4(aly * hx, v10, acc);
                    fma4(aly * lx, v11, acc);
                }
            }
            if (on) st4(out + (nq * M + m) * D + k * 4, acc);
        }
        pos += q_end - q_begin;
    }
}

template <typename VT, typename AT, bool FUSED>
cudaError_t launch_staged_typed(const FwdArgs& a, const CUtensorMap& vmap, int grid, int cap_rows, size_t smem_bytes,
                                long long per_cta, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (first_use_on_device(&configured)) {
        const cudaError_t e = cudaFuncSetAttribute(msda_fwd_staged_kernel<VT, AT, FUSED>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) return e;
    }
    const cape_msda_dims& d = a.d;
    msda_fwd_staged_kernel<VT, AT, FUSED><<<grid, kStThreads, smem_bytes, stream>>>(
        vmap, static_cast<const VT*>(a.value), a.shapes, a.starts, a.loc, a.attn, a.ref_points, static_cast<VT*>(a.out), d.N,
        d.S, d.M, d.Lq, cap_rows, per_cta);
    return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_staged_value(const FwdArgs& a, const CUtensorMap& vmap, int grid, int cap_rows, size_t smem_bytes,
                                long long per_cta, cudaStream_t stream) {
    if (a.fused) return launch_staged_typed<VT, float, true>(a, vmap, grid, cap_rows, smem_bytes, per_cta, stream);
    if (a.aux_dtype == CAPE_DTYPE_F32)
        return launch_staged_typed<VT, float, false>(a, vmap, grid, cap_rows, smem_bytes, per_cta, stream);
    return launch_staged_typed<VT, VT, false>(a, vmap, grid, cap_rows, smem_bytes, per_cta, stream);
}

}  // namespace

// Returns cudaErrorNotSupported when the configuration is outside this kernel (the caller then uses the L1 kernels).
cudaError_t launch_forward_staged(const FwdArgs& a, cudaStream_t stream) {
    const cape_msda_dims& d = a.d;
    if (d.D != 32 || d.P != 4 || d.L != 4) return cudaErrorNotSupported;
    const long long total = static_cast<long long>(d.N) * d.M * d.Lq;
    if (total < tuning(kTuneFwdStagedMinQm, 148 * 2048)) return cudaErrorNotSupported;   // small problems: latency kernels
    const int esize = a.value_dtype == CAPE_DTYPE_F32 ? 4 : 2;
    const int row_bytes = 32 * esize;
    const int budget_kb = min(tuning(kTuneFwdStagedKb, 200), kMaxDynSmem / 1024);
    const int cap_rows = (budget_kb * 1024 / (kBoxRows * row_bytes)) * kBoxRows;
    if (cap_rows < kBoxRows) return cudaErrorNotSupported;
    int smem_rows = cap_rows;
    const long long rows_total = static_cast<long long>(d.S);
    if (rows_total < smem_rows) smem_rows = static_cast<int>((rows_total + kBoxRows - 1) / kBoxRows) * kBoxRows;
    const size_t smem_bytes = static_cast<size_t>(smem_rows) * row_bytes;
    CUtensorMap vmap;
    if (!make_tensor_map_2d(&vmap, a.value, a.value_dtype, static_cast<uint64_t>(d.N) * d.S,
                            static_cast<uint64_t>(d.M) * d.D, kBoxRows, 32, CU_TENSOR_MAP_SWIZZLE_NONE))
        return cudaErrorNotSupported;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long per_cta = (total + sms - 1) / sms;
    per_cta = (per_cta + 127) / 128 * 128;                 // whole sweeps of the CTA's 32 warps x 4 queries
    const int grid = static_cast<int>((total + per_cta - 1) / per_cta);
    cudaError_t e;
    switch (a.value_dtype) {
        case CAPE_DTYPE_F32: e = launch_staged_value<float>(a, vmap, grid, smem_rows, smem_bytes, per_cta, stream); break;
        case CAPE_DTYPE_BF16: e = launch_staged_value<__nv_bfloat16>(a, vmap, grid, smem_rows, smem_bytes, per_cta, stream); break;
        case CAPE_DTYPE_F16: e = launch_staged_value<__half>(a, vmap, grid, smem_rows, smem_bytes, per_cta, stream); break;
        default: return cudaErrorInvalidValue;
    }
    return e;
}

}  // namespace cape
