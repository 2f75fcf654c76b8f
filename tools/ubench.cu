// Micro-benchmarks that size the MSDeformAttn kernel design on B200: how fast can an SM gather / scatter 128 B
// (fp32) or 64 B (bf16) value rows through L1, shared memory and L2 atomics?  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kIters = 256;      // gathers per warp-slot

// idx: per (block, iter, warp, slot) row index.  Each LDG.128 warp instruction covers 4 rows of 128 B (8 lanes each).
__global__ void gather_f32_v4(const float4* __restrict__ table, const int* __restrict__ idx, float4* __restrict__ out, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7;
    const int warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    const float4* base = table + (size_t)(blockIdx.x % window_stride) * rows_per_cta_window * 8;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        const float4 v = __ldg(base + (size_t)r * 8 + k);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x == 123.456f) out[threadIdx.x] = acc;
}

// same gather through the texture path (tex1Dfetch<float4> on a linear texture object): does TEX add gather throughput
// next to the LSU?
__global__ void gather_f32_tex(cudaTextureObject_t tex, const int* __restrict__ idx, float4* __restrict__ out, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7;
    const int warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    const int base = (blockIdx.x % window_stride) * rows_per_cta_window * 8;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        const float4 v = tex1Dfetch<float4>(tex, base + r * 8 + k);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x == 123.456f) out[threadIdx.x] = acc;
}

// alternate: even iterations through LDG.128, odd iterations through the texture path
__global__ void gather_f32_mixed(const float4* __restrict__ table, cudaTextureObject_t tex, const int* __restrict__ idx, float4* __restrict__ out, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7;
    const int warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    const int base = (blockIdx.x % window_stride) * rows_per_cta_window * 8;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 4
    for (int i = 0; i < kIters; i += 2) {
        const int r0 = __ldg(my + i * 4 + slot), r1 = __ldg(my + i * 4 + 4 + slot);
        const float4 v = __ldg(table + base + (size_t)r0 * 8 + k);
        const float4 u = tex1Dfetch<float4>(tex, base + r1 * 8 + k);
        acc.x += v.x + u.x; acc.y += v.y + u.y; acc.z += v.z + u.z; acc.w += v.w + u.w;
    }
    if (acc.x == 123.456f) out[threadIdx.x] = acc;
}

// lane = channel: one row (128 B) per LDG.32 warp instruction
__global__ void gather_f32_scalar(const float* __restrict__ table, const int* __restrict__ idx, float* __restrict__ out, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    const float* base = table + (size_t)(blockIdx.x % window_stride) * rows_per_cta_window * 32;
    float acc = 0;
#pragma unroll 8
    for (int i = 0; i < kIters * 4; ++i) {
        const int r = __ldg(my + i);
        acc += __ldg(base + (size_t)r * 32 + lane);
    }
    if (acc == 123.456f) out[threadIdx.x] = acc;
}

// bf16 rows (64 B): LDG.64, 8 lanes per row, 4 rows per instruction (row stride 64 B)
__global__ void gather_bf16_v2(const uint2* __restrict__ table, const int* __restrict__ idx, uint2* __restrict__ out, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    const uint2* base = table + (size_t)(blockIdx.x % window_stride) * rows_per_cta_window * 8;
    unsigned acc = 0;
#pragma unroll 8
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        const uint2 v = __ldg(base + (size_t)r * 8 + k);
        acc += v.x ^ v.y;
    }
    if (acc == 0x12345678u) out[threadIdx.x] = make_uint2(acc, acc);
}

// shared-memory gather: window of `rows` fp32 rows (128 B) staged in smem, LDS.128, 4 rows per instruction
__global__ void gather_smem_f32(const float4* __restrict__ table, const int* __restrict__ idx, float4* __restrict__ out, int rows) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) sm[i] = table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        const float4 v = sm[r * 8 + k];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x == 123.456f) out[threadIdx.x] = acc;
}

// shared-memory gather of bf16 rows (64 B): LDS.128, 4 lanes per row, 8 rows per instruction
__global__ void gather_smem_bf16(const uint4* __restrict__ table, const int* __restrict__ idx, uint4* __restrict__ out, int rows) {
    extern __shared__ uint4 smu[];
    for (int i = threadIdx.x; i < rows * 4; i += blockDim.x) smu[i] = table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, slot = lane >> 2, k = lane & 3, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    unsigned acc = 0;
#pragma unroll 8
    for (int i = 0; i < kIters / 2; ++i) {
        // pair rows (r, r+1) so that lanes 0-3 / 4-7 read 128 contiguous bytes (the x0 / x0+1 corner pair)
        const int r = __ldg(my + i * 8 + (slot >> 1) * 2) + (slot & 1);
        const uint4 v = smu[r * 4 + k];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[threadIdx.x] = make_uint4(acc, acc, acc, acc);
}

// read-modify-write of fp32 rows in shared memory (LDS.128 + 4 FADD + STS.128), 4 rows per instruction, no atomics
__global__ void rmw_smem_f32(const int* __restrict__ idx, float4* __restrict__ out, int rows) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) sm[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        float4 v = sm[r * 8 + k];
        v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 4.f;
        sm[r * 8 + k] = v;
    }
    __syncthreads();
    if (sm[threadIdx.x].x == 123.456f) out[threadIdx.x] = sm[threadIdx.x];
}

// atomicAdd(float) on shared memory (CAS loop), lane = channel, one row per warp instruction
__global__ void atomic_smem_f32(const int* __restrict__ idx, float* __restrict__ out, int rows) {
    extern __shared__ float smf[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) smf[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    for (int i = 0; i < kIters * 4; ++i) {
        const int r = __ldg(my + i);
        atomicAdd(&smf[r * 32 + lane], 1.f);
    }
    __syncthreads();
    if (smf[threadIdx.x] == 123.456f) out[threadIdx.x] = smf[threadIdx.x];
}

// global vector reductions: REDG.E.ADD.F32x4, 8 lanes per 128 B row, 4 rows per instruction
__global__ void red_global_v4(float* __restrict__ table, const int* __restrict__ idx, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    float* base = table + (size_t)(blockIdx.x % window_stride) * rows_per_cta_window * 32;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        float* p = base + (size_t)r * 32 + k * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
}

// global scalar reductions, lane = channel, one 128 B row per warp instruction
__global__ void red_global_scalar(float* __restrict__ table, const int* __restrict__ idx, int rows_per_cta_window, int window_stride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    float* base = table + (size_t)(blockIdx.x % window_stride) * rows_per_cta_window * 32;
#pragma unroll 4
    for (int i = 0; i < kIters * 4; ++i) {
        const int r = __ldg(my + i);
        atomicAdd(base + (size_t)r * 32 + lane, 1.f);
    }
}


// bulk reductions through the TMA unit: every lane group stages its 128 B row in shared memory (STS.128) and one lane per row
// issues cp.reduce.async.bulk (UBLKRED) to the row's global address; kStages commit groups in flight per warp
constexpr int kBulkStages = 8;
__global__ void red_bulk_tma(float* __restrict__ table, const int* __restrict__ idx, int rows_per_cta_window, int window_stride) {
    extern __shared__ __align__(128) float4 ring[];           // [warp][stage][4 rows][8 float4]
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    float* base = table + (size_t)(blockIdx.x % window_stride) * rows_per_cta_window * 32;
    float4* mine = ring + (size_t)warp * kBulkStages * 32;
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        float4* stage = mine + (i % kBulkStages) * 32;
        if (i >= kBulkStages) {                                  // the stage's previous bulk op must have read its rows
            if (k == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBulkStages - 1) : "memory");
            __syncwarp();
        }
        stage[slot * 8 + k] = make_float4(1.f, 2.f, 3.f, 4.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (k == 0) {
            const unsigned src = (unsigned)__cvta_generic_to_shared(stage + slot * 8);
            float* dst = base + (size_t)r * 32;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(dst), "r"(src) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (k == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1000.0);
    const int grid = sms * 16;                       // 2 waves at 8 CTAs/SM
    const size_t n_idx = (size_t)grid * (kThreads / 32) * kIters * 4;
    const double rows_total = (double)n_idx;         // every kernel touches n_idx rows in total
    struct Case { const char* name; int rows; int windows; };
    // rows per window: 128 rows = 16 KB (L1-resident: 8 CTAs per SM), 43520 rows = 5.57 MB fp32 (one image, L2), shared by all CTAs
    const Case cases[] = {{"L1-window 16KB per CTA", 128, 148}, {"L2 5.5MB shared", 43520, 1}, {"L2 111MB shared", 870400, 1}};
    float* table; CK(cudaMalloc(&table, (size_t)870400 * 128 + (size_t)148 * 512 * 128));
    CK(cudaMemset(table, 0, (size_t)870400 * 128 + (size_t)148 * 512 * 128));
    int* idx; CK(cudaMalloc(&idx, n_idx * sizeof(int)));
    float4* out; CK(cudaMalloc(&out, 1 << 20));
    std::vector<int> h(n_idx);
    cudaTextureObject_t tex = 0;
    {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = table;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = (size_t)870400 * 128 + (size_t)148 * 512 * 128;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    }
    for (const Case& c : cases) {
        srand(1);
        for (size_t i = 0; i < n_idx; ++i) h[i] = rand() % c.rows;
        CK(cudaMemcpy(idx, h.data(), n_idx * sizeof(int), cudaMemcpyHostToDevice));
        printf("== %s (%d rows of 128 B)\n", c.name, c.rows);
        auto report = [&](const char* what, float ms, double bytes_per_row) {
            const double rows_per_s = rows_total / (ms * 1e-3);
            printf("  %-34s %8.3f ms  %8.2f Grows/s  %8.2f TB/s  %6.2f clk/row/SM @%.0fMHz\n", what, ms, rows_per_s / 1e9,
                   rows_per_s * bytes_per_row / 1e12, (clk_khz * 1e3) / (rows_per_s / sms), clk_khz / 1000.0);
        };
        report("gather fp32 LDG.128 (4 rows/instr)", time_ms([&] { gather_f32_v4<<<grid, kThreads>>>((const float4*)table, idx, out, c.rows, c.windows); }), 128);
        report("gather fp32 TEX float4 (4 rows/instr)", time_ms([&] { gather_f32_tex<<<grid, kThreads>>>(tex, idx, out, c.rows, c.windows); }), 128);
        report("gather fp32 LDG.128 + TEX alternating", time_ms([&] { gather_f32_mixed<<<grid, kThreads>>>((const float4*)table, tex, idx, out, c.rows, c.windows); }), 128);
        report("gather fp32 LDG.32 (1 row/instr)", time_ms([&] { gather_f32_scalar<<<grid, kThreads>>>(table, idx, (float*)out, c.rows, c.windows); }), 128);
        report("gather bf16 LDG.64 (4 rows/instr)", time_ms([&] { gather_bf16_v2<<<grid, kThreads>>>((const uint2*)table, idx, (uint2*)out, c.rows, c.windows); }), 64);
        report("red.global.add.v4.f32 (4 rows/instr)", time_ms([&] { red_global_v4<<<grid, kThreads>>>(table, idx, c.rows, c.windows); }), 128);
        report("atomicAdd f32 scalar (1 row/instr)", time_ms([&] { red_global_scalar<<<grid, kThreads>>>(table, idx, c.rows, c.windows); }), 128);
        report("cp.reduce.async.bulk add.f32 128 B rows", time_ms([&] { red_bulk_tma<<<grid, kThreads, (kThreads / 32) * kBulkStages * 512>>>(table, idx, c.rows, c.windows); }), 128);
    }
    // shared-memory cases: 512 rows of 128 B = 64 KB per CTA (3 CTAs/SM)
    {
        const int rows = 512;
        srand(2);
        for (size_t i = 0; i < n_idx; ++i) h[i] = rand() % (rows - 1);
        CK(cudaMemcpy(idx, h.data(), n_idx * sizeof(int), cudaMemcpyHostToDevice));
        const int smem = rows * 128;
        CK(cudaFuncSetAttribute(gather_smem_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(gather_smem_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(rmw_smem_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(atomic_smem_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        printf("== shared memory, 64 KB window per CTA\n");
        auto report = [&](const char* what, float ms, double rows_done, double bytes_per_row) {
            const double rows_per_s = rows_done / (ms * 1e-3);
            printf("  %-34s %8.3f ms  %8.2f Grows/s  %8.2f TB/s  %6.2f clk/row/SM\n", what, ms, rows_per_s / 1e9,
                   rows_per_s * bytes_per_row / 1e12, (clk_khz * 1e3) / (rows_per_s / sms));
        };
        report("smem gather fp32 LDS.128", time_ms([&] { gather_smem_f32<<<grid, kThreads, smem>>>((const float4*)table, idx, out, rows); }), rows_total, 128);
        report("smem gather bf16 LDS.128 (pairs)", time_ms([&] { gather_smem_bf16<<<grid, kThreads, smem>>>((const uint4*)table, idx, (uint4*)out, rows * 2); }), rows_total, 64);
        report("smem RMW fp32 LDS+FADD+STS", time_ms([&] { rmw_smem_f32<<<grid, kThreads, smem>>>(idx, out, rows); }), rows_total, 128);
        report("smem atomicAdd f32 (CAS loop)", time_ms([&] { atomic_smem_f32<<<grid, kThreads, smem>>>(idx, (float*)out, rows); }), rows_total, 128);
    }
    printf("done\n");
    return 0;
}
