// Experiment: can the TMA engine (cp.reduce.async.bulk ... add.f32, 128 B per op, staged through shared memory) push
// fp32 row reductions into L2 faster than REDG.E.ADD.F32x4 (measured ceiling ~50 G rows/s)?  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tma_red tools/ubench_tma_red.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kIters = 256;

__global__ void red_v4(float* __restrict__ table, const int* __restrict__ idx) {
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        const int r = __ldg(my + i * 4 + slot);
        float* p = table + (size_t)r * 32 + k * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
}

// ROWS_PER_OP rows of 128 B are staged contiguously only when they are contiguous in global memory too; here every row
// goes to its own address, so one bulk op per row (128 B).
template <int STAGES>
__global__ void red_tma(float* __restrict__ table, const int* __restrict__ idx) {
    __shared__ __align__(128) float4 stage[kThreads / 32][STAGES][4][8];   // per warp: STAGES x 4 rows x 128 B
    const int lane = threadIdx.x & 31, slot = lane >> 3, k = lane & 7, warp = threadIdx.x >> 5;
    const int* my = idx + ((size_t)blockIdx.x * (kThreads / 32) + warp) * kIters * 4;
    for (int i = 0; i < kIters; ++i) {
        const int s = i % STAGES;
        // the bulk ops that read this stage STAGES iterations ago must have finished READING shared memory
        if (i >= STAGES) {
            if (STAGES == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (STAGES == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            __syncwarp();
        }
        const int r = __ldg(my + i * 4 + slot);
        stage[warp][s][slot][k] = make_float4(1.f, 2.f, 3.f, 4.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (k == 0) {
            float* dst = table + (size_t)r * 32;
            const unsigned src = (unsigned)__cvta_generic_to_shared(&stage[warp][s][slot][0]);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(dst), "r"(src) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    const int grid = sms * 16;
    const size_t n_idx = (size_t)grid * (kThreads / 32) * kIters * 4;
    const int rows = 43520;   // one 5.57 MB fp32 image, L2 resident
    float* table; CK(cudaMalloc(&table, (size_t)rows * 128));
    CK(cudaMemset(table, 0, (size_t)rows * 128));
    int* idx; CK(cudaMalloc(&idx, n_idx * sizeof(int)));
    std::vector<int> h(n_idx);
    srand(1);
    for (size_t i = 0; i < n_idx; ++i) h[i] = rand() % rows;
    CK(cudaMemcpy(idx, h.data(), n_idx * sizeof(int), cudaMemcpyHostToDevice));
    auto report = [&](const char* what, float ms) {
        printf("  %-44s %8.3f ms  %8.2f Grows/s\n", what, ms, n_idx / (ms * 1e-3) / 1e9);
    };
    report("REDG.E.ADD.F32x4 (4 rows / instr)", time_ms([&] { red_v4<<<grid, kThreads>>>(table, idx); }));
    report("TMA bulk reduce 128 B, 1 stage", time_ms([&] { red_tma<1><<<grid, kThreads>>>(table, idx); }));
    report("TMA bulk reduce 128 B, 2 stages", time_ms([&] { red_tma<2><<<grid, kThreads>>>(table, idx); }));
    report("TMA bulk reduce 128 B, 4 stages", time_ms([&] { red_tma<4><<<grid, kThreads>>>(table, idx); }));
    // check: total sum must equal n_idx * (1+2+3+4) * 8 per run ... (3 + 5 launches each kernel) just print a sample
    float hsum[4]; CK(cudaMemcpy(hsum, table, 16, cudaMemcpyDeviceToHost));
    printf("  sample row0 = %.0f %.0f %.0f %.0f (ratios must be 1:2:3:4)\n", hsum[0], hsum[1], hsum[2], hsum[3]);
    return 0;
}
