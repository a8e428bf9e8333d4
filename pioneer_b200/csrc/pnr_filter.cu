// N2: fused observation normaliser (the ConcurrentMeanStdFilter -- RLlib's MeanStdFilter behind a lock -- the reference launcher configures for its rollout workers,
// pioneer/launch/pioneer_knm_train.py:66 'observation_filter': 'ConcurrentMeanStdFilter').  ONE streaming pass over the
// observation batch does what the host-side filter does in three (push statistics, demean, scale): a CTA pulls a
// tile of 32 rows into shared memory with a TMA bulk load, every thread owns one of the 137 columns -- accumulates
// sum and sum of squares of (x - applied_mean) over the tile's rows, rewrites the column as
// clip((x - mean) * inv_std) -- and the tile leaves with a TMA bulk store.  Bound: HBM, 548 B read + 548 B written
// per row.  Statistics of the rows seen since the last synchronisation accumulate in float64 on the device
// (one atomicAdd per column per CTA); pnr_filter_sync merges them into the running mean / M2 (Chan et al.
// parallel update) and refreshes the applied mean / inverse std, so between synchronisations every rank
// normalises with the SAME statistics (what RLlib's synchronised filters converge to once per iteration).
#include <cuda_runtime.h>
#include <cstdint>
#include "pnr_device.cuh"
#include "pnr_launch.h"

#define PNR_FILTER_THREADS 160                      // >= PNR_OBS_DIM columns, 5 warps
#define PNR_FILTER_ROWS 32
#define PNR_FILTER_TILE_BYTES (PNR_FILTER_ROWS * PNR_OBS_DIM * 4)

__device__ __forceinline__ void pnr_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pnr_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pnr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pnr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pnr_bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pnr_smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(pnr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pnr_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP;\n"
        "}\n" ::"r"(pnr_smem_u32(bar)), "r"(parity) : "memory");
}

// applied: [mean[137], inv_std[137]] float; delta: [count, sum_d[137], sumsq_d[137]] double
__global__ void __launch_bounds__(PNR_FILTER_THREADS)
pnr_filter_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n_rows,
                  const float* __restrict__ applied, double* __restrict__ delta, float clip, int update,
                  int normalize) {
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const bool col_ok = tid < PNR_OBS_DIM;
    const float mean = col_ok ? applied[tid] : 0.f;
    const float inv_std = col_ok ? applied[PNR_OBS_DIM + tid] : 0.f;
    if (tid == 0) pnr_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int64_t n_tiles = (n_rows + PNR_FILTER_ROWS - 1) / PNR_FILTER_ROWS;
    double acc_s = 0.0, acc_q = 0.0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t row0 = t * PNR_FILTER_ROWS;
        const int rows = (int)((n_rows - row0) < PNR_FILTER_ROWS ? (n_rows - row0) : PNR_FILTER_ROWS);
        const float* src = in + row0 * PNR_OBS_DIM;
        float* dst = out + row0 * PNR_OBS_DIM;
        const bool bulk = (rows & 3) == 0;                      // byte count a multiple of 16
        if (bulk) {
            if (tid == 0) {
                pnr_mbar_expect_tx(&bar, (uint32_t)(rows * PNR_OBS_DIM * 4));
                pnr_bulk_load(tile, src, (uint32_t)(rows * PNR_OBS_DIM * 4), &bar);
            }
            pnr_mbar_wait(&bar, phase);
            phase ^= 1;
        } else {
            for (int i = tid; i < rows * PNR_OBS_DIM; i += PNR_FILTER_THREADS) tile[i] = src[i];
            __syncthreads();
        }
        if (col_ok) {
            // float64 accumulation (B200 has the FP64 rate for 2 DFMA per element of an HBM-bound stream): before the
            // first synchronisation the applied mean is 0, and sum(x^2) - sum(x)^2 / n of a constant column must
            // cancel to ~0, which float32 partial sums do not deliver
            double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};   // four independent FP64 chains
            int r = 0;
            for (; r + 3 < rows; r += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = tile[(r + k) * PNR_OBS_DIM + tid] - mean;
                    const double dd = (double)d;
                    s[k] += dd;
                    q[k] = fma(dd, dd, q[k]);
                    if (normalize) tile[(r + k) * PNR_OBS_DIM + tid] = fminf(fmaxf(d * inv_std, -clip), clip);
                }
            }
            for (; r < rows; ++r) {
                const float d = tile[r * PNR_OBS_DIM + tid] - mean;
                const double dd = (double)d;
                s[0] += dd;
                q[0] = fma(dd, dd, q[0]);
                if (normalize) tile[r * PNR_OBS_DIM + tid] = fminf(fmaxf(d * inv_std, -clip), clip);
            }
            acc_s += (s[0] + s[1]) + (s[2] + s[3]);
            acc_q += (q[0] + q[1]) + (q[2] + q[3]);
        }
        if (normalize || out != in) {
            pnr_fence_async_smem();
            __syncthreads();
            if (bulk) {
                if (tid == 0) {
                    pnr_bulk_store(dst, tile, (uint32_t)(rows * PNR_OBS_DIM * 4));
                    pnr_bulk_commit();
                    pnr_bulk_wait_read<0>();
                }
            } else {
                for (int i = tid; i < rows * PNR_OBS_DIM; i += PNR_FILTER_THREADS) dst[i] = tile[i];
            }
        }
        __syncthreads();                                        // the tile may be overwritten
    }
    double* slot = delta + (size_t)(blockIdx.x & (PNR_FILTER_SLOTS - 1)) * PNR_FILTER_DELTA_LEN;
    if (update && col_ok) {
        atomicAdd(&slot[1 + tid], acc_s);
        atomicAdd(&slot[1 + PNR_OBS_DIM + tid], acc_q);
    }
    if (update && blockIdx.x == 0 && tid == 0) atomicAdd(&delta[0], (double)n_rows);
}

// copy 0 += copies 1 .. SLOTS-1, which are cleared (see pnr_launch.h)
__global__ void pnr_filter_fold_kernel(double* __restrict__ slots) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= PNR_FILTER_DELTA_LEN) return;
    double acc = slots[c];
    for (int s = 1; s < PNR_FILTER_SLOTS; ++s) {
        acc += slots[(size_t)s * PNR_FILTER_DELTA_LEN + c];
        slots[(size_t)s * PNR_FILTER_DELTA_LEN + c] = 0.0;
    }
    slots[c] = acc;
}

// applied statistics from the running ones: mean (float) and 1 / (std + 1e-8); RunningStat.var = S / (n - 1) if n > 1 else
// mean^2 (RLlib, restated in oracle/filter_oracle.py)
__device__ __forceinline__ void pnr_filter_refresh_column(double count, double mean, double m2, float* applied, int c,
                                                          int demean, int destd) {
    if (!(count > 0.0)) {           // nothing pushed yet: identity (an RLlib filter that has seen no sample never normalises)
        applied[c] = 0.f;
        applied[PNR_OBS_DIM + c] = 1.f;
        return;
    }
    const double var = count > 1.0 ? m2 / (count - 1.0) : mean * mean;
    applied[c] = demean ? (float)mean : 0.f;
    applied[PNR_OBS_DIM + c] = destd ? (float)(1.0 / (sqrt(var) + 1e-8)) : 1.f;
}

__global__ void pnr_filter_refresh_kernel(const double* state, float* applied, int demean, int destd) {
    const int c = threadIdx.x;
    if (c < PNR_OBS_DIM) pnr_filter_refresh_column(state[0], state[1 + c], state[1 + PNR_OBS_DIM + c], applied, c, demean, destd);
}

// Once per iteration, entirely on the device (one CTA, thread = column): fold the accumulator copies (or take `merged`,
// the all-reduced delta of all ranks), merge into the running count / mean / M2 (Chan et al.; the batch sums are relative
// to the APPLIED mean, i.e. what the kernels subtracted), refresh the applied statistics, clear the accumulator.
__global__ void __launch_bounds__(PNR_FILTER_THREADS)
pnr_filter_sync_kernel(double* slots, const double* merged, double* state, float* applied, int demean, int destd) {
    // no __restrict__ here: the count is written by one thread and every thread carries its own copy of the new value
    const int c = threadIdx.x;
    const bool col_ok = c < PNR_OBS_DIM;
    double nb, sum_d = 0.0, sum_q = 0.0;
    if (merged) {
        nb = merged[0];
        if (col_ok) { sum_d = merged[1 + c]; sum_q = merged[1 + PNR_OBS_DIM + c]; }
    } else {
        nb = 0.0;
        for (int k = 0; k < PNR_FILTER_SLOTS; ++k) {
            const double* sl = slots + (size_t)k * PNR_FILTER_DELTA_LEN;
            nb += sl[0];
            if (col_ok) { sum_d += sl[1 + c]; sum_q += sl[1 + PNR_OBS_DIM + c]; }
        }
    }
    const double na = state[0];
    const double n = nb > 0.0 ? na + nb : na;
    __syncthreads();                                            // everybody has read the old count and the accumulator
    if (col_ok) {
        double mean = state[1 + c], m2 = state[1 + PNR_OBS_DIM + c];
        if (nb > 0.0) {
            const double applied_mean = demean ? (double)applied[c] : 0.0;
            const double mean_b = applied_mean + sum_d / nb;
            const double m2_b = sum_q - sum_d * sum_d / nb;
            const double delta = mean_b - mean;
            m2 += (m2_b > 0.0 ? m2_b : 0.0) + delta * delta * na * nb / n;
            mean += delta * nb / n;
            state[1 + c] = mean;
            state[1 + PNR_OBS_DIM + c] = m2;
        }
        pnr_filter_refresh_column(n, mean, m2, applied, c, demean, destd);
    }
    if (c == 0) state[0] = n;
    for (int k = 0; k < PNR_FILTER_SLOTS; ++k) {                // clear the accumulator
        double* sl = slots + (size_t)k * PNR_FILTER_DELTA_LEN;
        if (c == 0) sl[0] = 0.0;
        if (col_ok) { sl[1 + c] = 0.0; sl[1 + PNR_OBS_DIM + c] = 0.0; }
    }
}

cudaError_t pnr_launch_filter_refresh(const double* state, float* applied, int demean, int destd, cudaStream_t stream) {
    pnr_filter_refresh_kernel<<<1, PNR_FILTER_THREADS, 0, stream>>>(state, applied, demean, destd);
    return cudaGetLastError();
}

cudaError_t pnr_launch_filter_sync(double* slots, const double* merged, double* state, float* applied, int demean, int destd,
                                   cudaStream_t stream) {
    pnr_filter_sync_kernel<<<1, PNR_FILTER_THREADS, 0, stream>>>(slots, merged, state, applied, demean, destd);
    return cudaGetLastError();
}

cudaError_t pnr_launch_filter_fold(double* delta_slots, cudaStream_t stream) {
    pnr_filter_fold_kernel<<<(PNR_FILTER_DELTA_LEN + 127) / 128, 128, 0, stream>>>(delta_slots);
    return cudaGetLastError();
}

cudaError_t pnr_launch_filter(int device, const float* in, float* out, int64_t n_rows, const float* applied,
                              double* delta, float clip, int update, int normalize, cudaStream_t stream) {
    static int resident[PNR_MAX_DEVICES] = {};
    int& res = resident[device % PNR_MAX_DEVICES];
    if (res == 0) {
        cudaError_t e = cudaFuncSetAttribute((const void*)pnr_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             PNR_FILTER_TILE_BYTES);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)pnr_filter_kernel, PNR_FILTER_THREADS,
                                                      PNR_FILTER_TILE_BYTES);
        res = sms * (per_sm < 1 ? 1 : per_sm);
    }
    if (n_rows <= 0) return cudaSuccess;
    int64_t grid = (n_rows + PNR_FILTER_ROWS - 1) / PNR_FILTER_ROWS;
    if (grid > res) grid = res;
    pnr_filter_kernel<<<(unsigned)grid, PNR_FILTER_THREADS, PNR_FILTER_TILE_BYTES, stream>>>(in, out, n_rows, applied, delta,
                                                                                            clip, update, normalize);
    return cudaGetLastError();
}
