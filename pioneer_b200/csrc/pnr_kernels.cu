// Kernels of the kinematic reach env (Tier A): fused step, reset, observe, state get/set, stats.
// sm_100a; launched from pnr_api.cu.
#include "pnr_kernels.cuh"
#include "pnr_launch.h"

// ---------------------------------------------------------------------------------------------
// K1: the fused env step.  Per env: load 6 state planes + the action, integrate the 6 joints
// (act(), pioneer_knm_env.py:111-146), forward kinematics, reward / done / TimeLimit
// (:151-165, gym TimeLimit), episode statistics, in-kernel auto-reset (reset_world, :76-105),
// observation (observe(), :184-211) staged per warp in shared memory and streamed out.
// Grid-stride over warp tiles; grid = resident CTAs of the whole GPU (multiple of the SM count).
// ---------------------------------------------------------------------------------------------
template <int ARITH, int OBS_MODE>
__global__ void __launch_bounds__(PNR_STEP_THREADS)
pnr_step_kernel(const __grid_constant__ PnrParams p, float4* __restrict__ state, const float* __restrict__ actions,
                float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                PnrStats* __restrict__ stats, uint32_t tick) {
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = smem + warp * PNR_TILE_FLOATS;
    float* row = tile + lane * PNR_OBS_DIM;
    const int64_t N = p.n_envs;
    const int64_t n_tiles = (N + PNR_TILE_ENVS - 1) / PNR_TILE_ENVS;
    const int64_t stride = (int64_t)gridDim.x * PNR_STEP_WARPS;
    int64_t t_idx = (int64_t)blockIdx.x * PNR_STEP_WARPS + warp;
    if (t_idx >= n_tiles) return;

    pnr_pack_obs_const(p, row);                               // obs[18:54] never change: once per warp

    // software pipeline: the loads of tile i+1 are in flight while tile i is computed
    PnrRaw raw;
    {
        const int64_t e0 = t_idx * PNR_TILE_ENVS + lane;
        pnr_load_raw(state, actions, N, e0 < N ? e0 : N - 1, raw);
    }
    bool tile_busy = false;                                   // a bulk store of this warp's tile is in flight
    for (; t_idx < n_tiles; t_idx += stride) {
        const int64_t env_raw = t_idx * PNR_TILE_ENVS + lane;
        const bool active = env_raw < N;
        const int64_t env = active ? env_raw : N - 1;       // tail lanes shadow the last env, stores masked

        PnrEnv s;
        pnr_unpack_raw(raw, s);
        const float2 act01 = raw.a01, act23 = raw.a23, act45 = raw.a45;
        if (t_idx + stride < n_tiles) {
            const int64_t en = (t_idx + stride) * PNR_TILE_ENVS + lane;
            pnr_load_raw(state, actions, N, en < N ? en : N - 1, raw);
        }

        // --- act(): integrate with the PREVIOUS action (one-step actuation delay), then latch the new one
#pragma unroll
        for (int i = 0; i < PNR_DOF; ++i) {
            float v1, r1;
            pnr_integrate_joint<ARITH>(p, i, s.a[i], s.v[i], s.r[i], v1, r1);
            s.v[i] = v1; s.r[i] = r1;
        }
        s.a[0] = act01.x; s.a[1] = act01.y; s.a[2] = act23.x; s.a[3] = act23.y; s.a[4] = act45.x; s.a[5] = act45.y;

        // --- pose, distance, reward, done (r is inside the joint limits after the integrator: fast sincos)
        PnrPose o;
        pnr_pose<true>(p, s, o);
        bool reached = o.dist < p.done_distance;
        if (fabsf(o.dist - p.done_distance) < p.done_band)   // decide in float64 where float32 could flip it
            pnr_fk_tip_f64(p, s.r, s.tgt, o.ptr, o.dist, reached);
        const float pot_new = pnr_potential(p, o.dist);
        // (potential - old_potential) + (-penalty_step) + (award_done | 0), pioneer_knm_env.py:162-165
        const float rew = __fadd_rn(__fadd_rn(__fsub_rn(pot_new, s.pot), -p.penalty_step), reached ? p.award_done : 0.f);
        s.pot = pot_new;
        s.t += 1;
        s.ep_ret = __fadd_rn(s.ep_ret, rew);
        const bool timeout = p.max_episode_steps > 0 && s.t >= p.max_episode_steps;
        const bool is_done = reached || timeout;
        const uint8_t flags = (is_done ? PNR_DONE : 0) | ((timeout && !reached) ? PNR_TRUNCATED : 0);
        if (active) {
            reward[env] = rew;
            done[env] = flags;
        }

        // --- episode statistics: one set of atomics per warp that saw an episode end
        const unsigned done_mask = __ballot_sync(PNR_FULL_MASK, is_done && active);
        if (done_mask) {
            const bool mine = (done_mask >> lane) & 1u;
            float ret = mine ? s.ep_ret : 0.f, ret2 = ret * ret, len = mine ? (float)s.t : 0.f;
            float mx = mine ? s.ep_ret : -INFINITY, mn = mine ? s.ep_ret : INFINITY;
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) {
                ret += __shfl_xor_sync(PNR_FULL_MASK, ret, ofs);
                ret2 += __shfl_xor_sync(PNR_FULL_MASK, ret2, ofs);
                len += __shfl_xor_sync(PNR_FULL_MASK, len, ofs);
                mx = fmaxf(mx, __shfl_xor_sync(PNR_FULL_MASK, mx, ofs));
                mn = fminf(mn, __shfl_xor_sync(PNR_FULL_MASK, mn, ofs));
            }
            const unsigned reach_mask = __ballot_sync(PNR_FULL_MASK, reached && active);
            if (lane == 0) {
                atomicAdd(&stats->episodes, (double)__popc(done_mask));
                atomicAdd(&stats->sum_return, (double)ret);
                atomicAdd(&stats->sum_length, (double)len);
                atomicAdd(&stats->sum_return_sq, (double)ret2);
                atomicAdd(&stats->reached, (double)__popc(reach_mask));
                atomicMax(&stats->max_return_ord, pnr_float_to_ordered(mx));
                atomicMin(&stats->min_return_ord, pnr_float_to_ordered(mn));
            }
        }

        // --- observation + auto-reset.  The previous tile's bulk store must have finished reading smem.
        if (tile_busy) pnr_tile_wait(lane);
        const bool do_reset = is_done && (p.auto_reset != 0);
        if (OBS_MODE == PNR_OBS_TERMINAL) {
            pnr_pack_obs_dyn<true>(p, row, s, o, s.pot);       // what BulletEnv.step returns
            if (do_reset) {
                float q[PNR_DOF], tg[3];
                pnr_reset_draws(p, p.env_id_base + env, tick, q, tg);
                pnr_reset_env(s, q, tg);
            }
        } else {
            if (do_reset) {
                float q[PNR_DOF], tg[3];
                pnr_reset_draws(p, p.env_id_base + env, tick, q, tg);
                pnr_reset_env(s, q, tg);
            }
            if (__any_sync(PNR_FULL_MASK, do_reset)) {         // warp-uniform; rare
                PnrPose o2;
                pnr_pose<true>(p, s, o2);                      // fresh joint angles lie inside the limits
                if (do_reset) o = o2;
            }
            pnr_pack_obs_dyn<true>(p, row, s, o, s.pot);       // first observation of the next episode
        }
        if (active) pnr_store_env(state, N, env, s);

        const int64_t rows_left = N - t_idx * PNR_TILE_ENVS;
        pnr_emit_tile(tile, obs + t_idx * (int64_t)PNR_TILE_FLOATS,
                      rows_left < PNR_TILE_ENVS ? (int)rows_left : PNR_TILE_ENVS, lane);
        tile_busy = true;
    }
    if (lane == 0) pnr_bulk_wait_read();                      // smem must outlive the copy engine's reads
}

// ---------------------------------------------------------------------------------------------
// K0: reset / observe on a list of envs (idx == nullptr: env k = k).  MODE 0 = reset, 1 = observe.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(PNR_STEP_THREADS)
pnr_reset_observe_kernel(const __grid_constant__ PnrParams p, float4* __restrict__ state,
                         const int64_t* __restrict__ idx, int64_t n, const float* __restrict__ q0,
                         const float* __restrict__ target, float* __restrict__ obs_out, uint32_t tick) {
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = smem + warp * PNR_TILE_FLOATS;
    const int64_t N = p.n_envs;
    const int64_t n_tiles = (n + PNR_TILE_ENVS - 1) / PNR_TILE_ENVS;
    bool tile_busy = false;
    for (int64_t t_idx = (int64_t)blockIdx.x * PNR_STEP_WARPS + warp; t_idx < n_tiles;
         t_idx += (int64_t)gridDim.x * PNR_STEP_WARPS) {
        const int64_t k_raw = t_idx * PNR_TILE_ENVS + lane;
        const bool active = k_raw < n;
        const int64_t k = active ? k_raw : n - 1;
        const int64_t env = idx ? idx[k] : k;
        PnrEnv s;
        if (MODE == 0) {
            float q[PNR_DOF], tg[3];
            pnr_reset_draws(p, p.env_id_base + env, tick, q, tg);
            if (q0) {
#pragma unroll
                for (int i = 0; i < PNR_DOF; ++i) q[i] = q0[k * PNR_DOF + i];
            }
            if (target) { tg[0] = target[k * 3 + 0]; tg[1] = target[k * 3 + 1]; tg[2] = target[k * 3 + 2]; }
            pnr_reset_env(s, q, tg);
            if (active) pnr_store_env(state, N, env, s);
        } else {
            pnr_load_env(state, N, env, s);
        }
        if (obs_out) {
            // injected joint angles may lie anywhere: library sincos for every column here (not the hot path)
            PnrPose o;
            pnr_pose<false>(p, s, o);
            bool within;
            if (fabsf(o.dist - p.done_distance) < p.done_band) pnr_fk_tip_f64(p, s.r, s.tgt, o.ptr, o.dist, within);
            if (tile_busy) pnr_tile_wait(lane);
            pnr_pack_obs_const(p, tile + lane * PNR_OBS_DIM);
            pnr_pack_obs_dyn<false>(p, tile + lane * PNR_OBS_DIM, s, o, s.pot);
            const int64_t rows_left = n - t_idx * PNR_TILE_ENVS;
            pnr_emit_tile(tile, obs_out + t_idx * (int64_t)PNR_TILE_FLOATS,
                          rows_left < PNR_TILE_ENVS ? (int)rows_left : PNR_TILE_ENVS, lane);
            tile_busy = true;
        }
    }
    if (lane == 0) pnr_bulk_wait_read();
}

// ---------------------------------------------------------------------------------------------
// state planes <-> row-major arrays (Joint.position()/velocity(), env.a/v/r/potential); utility path
// ---------------------------------------------------------------------------------------------
template <bool SET>
__global__ void pnr_state_io_kernel(float4* __restrict__ state, int64_t N, float* r, float* v, float* a,
                                    float* potential, float* target, int32_t* t, float* ep_return) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    PnrEnv s;
    pnr_load_env(state, N, e, s);
#pragma unroll
    for (int i = 0; i < PNR_DOF; ++i) {
        if (r) { if (SET) s.r[i] = r[e * PNR_DOF + i]; else r[e * PNR_DOF + i] = s.r[i]; }
        if (v) { if (SET) s.v[i] = v[e * PNR_DOF + i]; else v[e * PNR_DOF + i] = s.v[i]; }
        if (a) { if (SET) s.a[i] = a[e * PNR_DOF + i]; else a[e * PNR_DOF + i] = s.a[i]; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (target) { if (SET) s.tgt[i] = target[e * 3 + i]; else target[e * 3 + i] = s.tgt[i]; }
    if (potential) { if (SET) s.pot = potential[e]; else potential[e] = s.pot; }
    if (t) { if (SET) s.t = t[e]; else t[e] = s.t; }
    if (ep_return) { if (SET) s.ep_ret = ep_return[e]; else ep_return[e] = s.ep_ret; }
    if (SET) pnr_store_env(state, N, e, s);
}

__global__ void pnr_stats_snapshot_kernel(PnrStats* stats, double env_steps, double* out, int clear) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = stats->episodes; out[1] = stats->sum_return; out[2] = stats->sum_length;
        out[3] = stats->sum_return_sq;
        out[4] = (double)pnr_ordered_to_float(stats->max_return_ord);
        out[5] = (double)pnr_ordered_to_float(stats->min_return_ord);
        out[6] = env_steps; out[7] = stats->reached;
        if (clear) {
            stats->episodes = stats->sum_return = stats->sum_length = stats->sum_return_sq = stats->reached = 0.0;
            stats->max_return_ord = pnr_float_to_ordered(-INFINITY);
            stats->min_return_ord = pnr_float_to_ordered(INFINITY);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-callable launchers (declared in pnr_launch.h)
// ---------------------------------------------------------------------------------------------
static int pnr_resident_grid(const void* fn, size_t smem) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PNR_STEP_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

template <typename K>
static cudaError_t pnr_prepare(K kernel, int* grid_out) {
    cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PNR_STEP_SMEM);
    if (e != cudaSuccess) return e;
    *grid_out = pnr_resident_grid((const void*)kernel, PNR_STEP_SMEM);
    return cudaSuccess;
}

static int64_t pnr_grid_for(int64_t n_units, int resident) {
    const int64_t blocks = (n_units + PNR_TILE_ENVS * PNR_STEP_WARPS - 1) / (PNR_TILE_ENVS * PNR_STEP_WARPS);
    return blocks < resident ? (blocks > 0 ? blocks : 1) : resident;
}

cudaError_t pnr_launch_step(const PnrParams& p, int device, int arith, int obs_mode, float4* state, const float* actions, float* obs,
                            float* reward, uint8_t* done, PnrStats* stats, uint32_t tick, cudaStream_t stream) {
    typedef void (*Kern)(const PnrParams, float4*, const float*, float*, float*, uint8_t*, PnrStats*, uint32_t);
    static Kern kernels[2][2] = {
        {pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_TERMINAL>, pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_AUTORESET>},
        {pnr_step_kernel<PNR_ARITH_LEGACY64, PNR_OBS_TERMINAL>, pnr_step_kernel<PNR_ARITH_LEGACY64, PNR_OBS_AUTORESET>}};
    static int grids[PNR_MAX_DEVICES][2][2] = {};     // cudaFuncSetAttribute is per device
    Kern k = kernels[arith][obs_mode];
    int& resident = grids[device % PNR_MAX_DEVICES][arith][obs_mode];
    if (resident == 0) {
        cudaError_t e = pnr_prepare(k, &resident);
        if (e != cudaSuccess) return e;
    }
    const int64_t grid = pnr_grid_for(p.n_envs, resident);
    k<<<(unsigned)grid, PNR_STEP_THREADS, PNR_STEP_SMEM, stream>>>(p, state, actions, obs, reward, done, stats, tick);
    return cudaGetLastError();
}

cudaError_t pnr_launch_reset_observe(const PnrParams& p, int device, int mode, float4* state, const int64_t* idx, int64_t n,
                                     const float* q0, const float* target, float* obs_out, uint32_t tick,
                                     cudaStream_t stream) {
    typedef void (*Kern)(const PnrParams, float4*, const int64_t*, int64_t, const float*, const float*, float*, uint32_t);
    static Kern kernels[2] = {pnr_reset_observe_kernel<0>, pnr_reset_observe_kernel<1>};
    static int grids[PNR_MAX_DEVICES][2] = {};
    int& resident = grids[device % PNR_MAX_DEVICES][mode];
    if (resident == 0) {
        cudaError_t e = pnr_prepare(kernels[mode], &resident);
        if (e != cudaSuccess) return e;
    }
    if (n <= 0) return cudaSuccess;
    const int64_t grid = pnr_grid_for(n, resident);
    kernels[mode]<<<(unsigned)grid, PNR_STEP_THREADS, PNR_STEP_SMEM, stream>>>(p, state, idx, n, q0, target, obs_out, tick);
    return cudaGetLastError();
}

cudaError_t pnr_launch_state_io(bool set, float4* state, int64_t N, float* r, float* v, float* a, float* potential,
                                float* target, int32_t* t, float* ep_return, cudaStream_t stream) {
    const int threads = 256;
    const unsigned grid = (unsigned)((N + threads - 1) / threads);
    if (set) pnr_state_io_kernel<true><<<grid, threads, 0, stream>>>(state, N, r, v, a, potential, target, t, ep_return);
    else pnr_state_io_kernel<false><<<grid, threads, 0, stream>>>(state, N, r, v, a, potential, target, t, ep_return);
    return cudaGetLastError();
}

cudaError_t pnr_launch_stats_snapshot(PnrStats* stats, double env_steps, double* out, int clear, cudaStream_t stream) {
    pnr_stats_snapshot_kernel<<<1, 32, 0, stream>>>(stats, env_steps, out, clear);
    return cudaGetLastError();
}
