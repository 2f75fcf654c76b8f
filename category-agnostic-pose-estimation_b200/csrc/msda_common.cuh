// Shared device helpers for the MSDeformAttn kernels (sm_100a).
//
// Arithmetic contract (what the parity tests pin): for one sample at normalised location (locx, locy) on a level
// of size (H, W), following /root/reference/models/deformable_transformer.py:129,136-137
//     g = 2*loc - 1                           (:129)
//     x = (g + 1) * (W/2) - 0.5               (grid_sample, align_corners=False; separate mul and sub, like ATen's CPU path)
//     x0 = floor(x), lx = x - x0, weights (1-lx, lx); same for y
//     corners (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1); a corner outside the map contributes nothing (zeros padding)
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_launch.h"

namespace cape {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- element access: 4 consecutive channels <-> float4 ------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    float4 f;
    f.x = __uint_as_float(r.x << 16);
    f.y = __uint_as_float(r.x & 0xffff0000u);
    f.z = __uint_as_float(r.y << 16);
    f.w = __uint_as_float(r.y & 0xffff0000u);
    return f;
}
__device__ __forceinline__ float4 ld4(const __half* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
// Streaming loads for tensors that are touched once (sampling_locations, attention_weights, grad_out): read-only path
// without L1 allocation, so the L1 keeps the value rows that neighbouring queries gather again.
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld2_stream(const float* p) {
    float2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld1_stream(const float* p) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld1_stream(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld1_stream(const __half* p) { return __half2float(*p); }

// 4-channel load that yields zeros when `pred` is false (zeros padding: an out-of-bounds corner contributes nothing).
// The destination is zero-initialised in C++ and conditionally overwritten by a predicated load inside one asm
// statement ("+" constraints), which keeps ptxas from loading into temporaries and selecting afterwards.
__device__ __forceinline__ float4 ld4_or_zero(const float* p, bool pred) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef CAPE_EXP_EVICT_LAST   // profiling variant: keep gathered value rows in L1 with evict-last priority
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(p), "r"(static_cast<int>(pred)));
    return v;
#endif
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(p), "r"(static_cast<int>(pred)));
    return v;
}
__device__ __forceinline__ uint2 ld2u_or_zero(const void* p, bool pred) {
    uint2 r = make_uint2(0u, 0u);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
        : "+r"(r.x), "+r"(r.y)
        : "l"(p), "r"(static_cast<int>(pred)));
    return r;
}
__device__ __forceinline__ float4 ld4_or_zero(const __nv_bfloat16* p, bool pred) {
    const uint2 r = ld2u_or_zero(p, pred);
    float4 f;
    f.x = __uint_as_float(r.x << 16);
    f.y = __uint_as_float(r.x & 0xffff0000u);
    f.z = __uint_as_float(r.y << 16);
    f.w = __uint_as_float(r.y & 0xffff0000u);
    return f;
}
__device__ __forceinline__ float4 ld4_or_zero(const __half* p, bool pred) {
    const uint2 r = ld2u_or_zero(p, pred);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
// acc += w * v as two packed FFMA2 (sm_100 `fma.rn.f32x2`, scalar operand broadcast): same rounding as four FFMA, half
// the issue slots — the sampling kernels are issue-limited as much as LSU-limited.
__device__ __forceinline__ void fma4(float w, const float4& v, float4& acc) {
    asm("{\n\t.reg .b64 a, b, c;\n\t"
        "mov.b64 a, {%4, %4};\n\t"
        "mov.b64 b, {%5, %6};\n\tmov.b64 c, {%0, %1};\n\tfma.rn.f32x2 c, a, b, c;\n\tmov.b64 {%0, %1}, c;\n\t"
        "mov.b64 b, {%7, %8};\n\tmov.b64 c, {%2, %3};\n\tfma.rn.f32x2 c, a, b, c;\n\tmov.b64 {%2, %3}, c;\n\t}"
        : "+f"(acc.x), "+f"(acc.y), "+f"(acc.z), "+f"(acc.w)
        : "f"(w), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
// c * g as two packed FMUL2
__device__ __forceinline__ float4 mul4(float c, const float4& g) {
    float4 r;
    asm("{\n\t.reg .b64 a, b, d;\n\t"
        "mov.b64 a, {%4, %4};\n\t"
        "mov.b64 b, {%5, %6};\n\tmul.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t"
        "mov.b64 b, {%7, %8};\n\tmul.rn.f32x2 d, a, b;\n\tmov.b64 {%2, %3}, d;\n\t}"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "f"(c), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w));
    return r;
}
__device__ __forceinline__ void st4(float* p, const float4& v) {
#ifdef CAPE_EXP_STREAM_STORE   // profiling variant: streaming (evict-first) output stores
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    return;
#endif
    *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float4& v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ void st4(__half* p, const float4& v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// fp32 vector reduction into global memory (REDG.E.ADD.F32x4 on sm_90+; p must be 16-byte aligned), executed only when
// `pred`: the predicate rides on the REDG itself, no branch in the source.
__device__ __forceinline__ void red_add4_if(float* p, bool pred, float x, float y, float z, float w) {
#ifdef CAPE_EXP_RED_CTA_SCOPE   // profiling variant: CTA-scope reduction (does the scope change where the add happens?)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p red.relaxed.cta.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w), "r"(static_cast<int>(pred)) : "memory");
    return;
#endif
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w), "r"(static_cast<int>(pred)) : "memory");
}

// Pixel coordinates of one sample.  Returns false when no corner can be in bounds (also for NaN / inf locations),
// in which case x0/y0/lx/ly are not written.
__device__ __forceinline__ bool sample_coords(float locx, float locy, int H, int W, int& x0, int& y0, float& lx, float& ly) {
    const float gx = fmaf(2.f, locx, -1.f);            // exact product, one rounding: same as mul then sub
    const float gy = fmaf(2.f, locy, -1.f);
    const float x = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f * static_cast<float>(W)), 0.5f);
    const float y = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f * static_cast<float>(H)), 0.5f);
    if (!(x >= -1.f && x < static_cast<float>(W) && y >= -1.f && y < static_cast<float>(H))) return false;
    const float xf = floorf(x), yf = floorf(y);
    lx = x - xf;
    ly = y - yf;
    x0 = static_cast<int>(xf);
    y0 = static_cast<int>(yf);
    return true;
}

// Pixel coordinate of one normalised location component on a map dimension `dim` (= W for x, H for y), or the
// sentinel -4 when no corner can be in bounds along this axis (also for NaN / inf): with the sentinel, floor() gives
// -4 and both unsigned corner tests ((unsigned)i < dim, (unsigned)(i + 1) < dim) fail, so no separate "ok" flag is
// carried around.  Arithmetic order as in sample_coords().
__device__ __forceinline__ float pixel_coord(float loc, float dimf) {
    const float g = fmaf(2.f, loc, -1.f);
    const float x = __fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f * dimf), 0.5f);
    return (x >= -1.f && x < dimf) ? x : -4.f;
}

// Warp index that the compiler can prove uniform (so shuffles in warp-uniform loops need no reconvergence code).
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(kFullMask, static_cast<int>(threadIdx.x >> 5), 0); }

// <a, b> over 4 channels: one FMUL2 + one FFMA2 + one FADD  ((a.x b.x + a.z b.z) + (a.y b.y + a.w b.w))
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    float lo, hi;
    asm("{\n\t.reg .b64 p, q, t;\n\t"
        "mov.b64 p, {%2, %3};\n\tmov.b64 q, {%6, %7};\n\tmul.rn.f32x2 t, p, q;\n\t"
        "mov.b64 p, {%4, %5};\n\tmov.b64 q, {%8, %9};\n\tfma.rn.f32x2 t, p, q, t;\n\t"
        "mov.b64 {%0, %1}, t;\n\t}"
        : "=f"(lo), "=f"(hi)
        : "f"(a.z), "f"(a.w), "f"(a.x), "f"(a.y), "f"(b.z), "f"(b.w), "f"(b.x), "f"(b.y));
    return lo + hi;
}

}  // namespace cape
