// FP32 throughput by instruction FORM on this GPU (developer microbenchmark; python tools/fp32_forms.py):
//   0  FFMA  x = fma(x, a, b), a / b launch-uniform (uniform-register / constant operands: ONE vector register read)
//   1  FFMA  x = fma(x, y, z), y / z per-thread registers (three vector register reads) -- what real code looks like
//   2  FFMA2 the packed form (fma.rn.f32x2 on register pairs), three register-pair reads, 2 FMAs per lane per instruction
//   3  FMUL  x = x * y          4  FADD  x = x + y          (two vector register reads)
// 8 independent chains per thread (4 for the packed form: same flops), 2,048 threads per SM.  Prints TFLOP/s with an FMA = 2
// flops, a MUL / ADD = 1, and warp instructions per clock per SM sub-partition at the sampled SM clock.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; "
        "mov.b64 {%0,%1}, rd; }" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) form_kernel(float* out, const float* in, int iters, float a, float b) {
    float x[8], y[8], z[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] = threadIdx.x * 1e-3f + k;
        y[k] = in[(threadIdx.x + 32 * k) & 1023];            // per-thread, per-chain: no operand reuse
        z[k] = in[(threadIdx.x + 32 * k + 7) & 1023];
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    const float2 v = fma2(make_float2(x[k], x[k + 1]), make_float2(y[k], y[k + 1]), make_float2(z[k], z[k + 1]));
                    x[k] = v.x; x[k + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (MODE == 0) x[k] = fmaf(x[k], a, b);
                    if (MODE == 1) x[k] = fmaf(x[k], y[k], z[k]);
                    if (MODE == 3) x[k] = x[k] * y[k];
                    if (MODE == 4) x[k] = x[k] + y[k];
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static double run(int iters, int blocks, float* out, const float* in) {
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    form_kernel<MODE><<<blocks, 256>>>(out, in, iters / 10, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(s);
        form_kernel<MODE><<<blocks, 256>>>(out, in, iters, 0.999f, 1e-3f);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

// out5: seconds per launch for the five forms; returns the number of scalar operations per launch (same for all forms)
extern "C" double fp32_forms(int iters, double* out5) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8;
    float *out = nullptr, *in = nullptr;
    cudaMalloc(&out, sizeof(float) * blocks * 256);
    cudaMalloc(&in, sizeof(float) * 1024);
    float host[1024];
    for (int i = 0; i < 1024; ++i) host[i] = 0.999f + 1e-6f * i;
    cudaMemcpy(in, host, sizeof(host), cudaMemcpyHostToDevice);
    out5[0] = run<0>(iters, blocks, out, in);
    out5[1] = run<1>(iters, blocks, out, in);
    out5[2] = run<2>(iters, blocks, out, in);
    out5[3] = run<3>(iters, blocks, out, in);
    out5[4] = run<4>(iters, blocks, out, in);
    cudaFree(out); cudaFree(in);
    return 128.0 * (double)iters * blocks * 256;             // 8 chains x 16 per iteration, per thread
}
