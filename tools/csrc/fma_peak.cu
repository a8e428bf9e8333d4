// FP32 FMA peak microbenchmark (MEASURED_PEAKS.json has no FP32 figure): 8 independent FMA chains per thread,
// full occupancy.  extern "C" double fma_peak_tflops(int iters) -> achieved TFLOP/s (2 flops per FMA).
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
          x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" double fma_peak_tflops(int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    float* out = nullptr;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    fma_kernel<<<blocks, threads>>>(out, iters / 10, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(s);
        fma_kernel<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s, e);
        const double flops = 2.0 * 128.0 * (double)iters * blocks * threads;      // 8 chains x 16 per iteration
        const double t = flops / (ms * 1e-3) / 1e12;
        if (t > best) best = t;
    }
    cudaFree(out);
    return best;
}
