// Developer probe: what a per-step CUDA-event pair reads around (almost) empty kernels after a 512 MiB L2 flush, as a
// function of the kernel's parameter-block size and grid size.  Answers whether the 1.3 KB __grid_constant__ parameter
// block of pnr_step_kernel (or its 888-CTA grid) is part of the fixed ~6 us the bench's event pairs carry.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o launch_floor tools/csrc/launch_floor.cu && ./launch_floor
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

struct Small { float x[4]; };
struct Big { float x[330]; };          // 1,320 bytes, the size of PnrParams

__global__ void k_small(const __grid_constant__ Small p, float* out) { if (threadIdx.x == 0 && blockIdx.x == 0 && p.x[0] == 123.f) out[0] = p.x[1]; }
__global__ void k_big(const __grid_constant__ Big p, float* out) { if (threadIdx.x == 0 && blockIdx.x == 0 && p.x[0] == 123.f) out[0] = p.x[329]; }
// touches its whole parameter block (like the step kernel) and a little shared memory per CTA
__global__ void k_big_touch(const __grid_constant__ Big p, float* out) {
    extern __shared__ float sm[];
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 330; ++i) acc += p.x[i];              // compile-time indices: constant-bank operands, no local copy
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (sm[(threadIdx.x + 1) % blockDim.x] == 123.f) out[blockIdx.x] = acc;
}

__global__ void k_small_smem(const __grid_constant__ Small p, float* out) {
    extern __shared__ float sm[];
    sm[threadIdx.x] = p.x[0];
    __syncthreads();
    if (sm[(threadIdx.x + 1) % blockDim.x] == 123.f) out[blockIdx.x] = 1.f;
}

template <typename F>
static float median_us(F launch, void* flush, size_t flush_bytes, int reps = 200) {
    std::vector<float> t;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < reps + 10; ++i) {
        cudaMemsetAsync(flush, 0, flush_bytes, 0);
        cudaEventRecord(a, 0);
        launch();
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (i >= 10) t.push_back(ms * 1e3f);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main() {
    void* flush; float* out;
    const size_t fb = 512ull << 20;
    cudaMalloc(&flush, fb); cudaMalloc(&out, 1 << 20);
    Small s = {}; Big b = {};
    cudaFuncSetAttribute(k_big_touch, cudaFuncAttributeMaxDynamicSharedMemorySize, 36 * 1024);
    cudaFuncSetAttribute(k_small_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 36 * 1024);
    printf("small params, 1 CTA           %6.2f us\n", median_us([&] { k_small<<<1, 32>>>(s, out); }, flush, fb));
    printf("1320 B params, 1 CTA          %6.2f us\n", median_us([&] { k_big<<<1, 32>>>(b, out); }, flush, fb));
    printf("small params, 888 CTAs x 128  %6.2f us\n", median_us([&] { k_small<<<888, 128>>>(s, out); }, flush, fb));
    printf("1320 B params, 888 CTAs x 128 %6.2f us\n", median_us([&] { k_big<<<888, 128>>>(b, out); }, flush, fb));
    printf("1320 B params read by every CTA, 888 x 128, 35 KB smem %6.2f us\n",
           median_us([&] { k_big_touch<<<888, 128, 35 * 1024>>>(b, out); }, flush, fb));
    printf("small params, 888 x 128, 35 KB smem, params untouched %6.2f us\n",
           median_us([&] { k_small_smem<<<888, 128, 35 * 1024>>>(s, out); }, flush, fb));
    printf("no flush: small 1 CTA         %6.2f us\n", median_us([&] { k_small<<<1, 32>>>(s, out); }, flush, 4));
    printf("no flush: 1320 B, 888 x 128 touch %6.2f us\n", median_us([&] { k_big_touch<<<888, 128, 35 * 1024>>>(b, out); }, flush, 4));
    return 0;
}
