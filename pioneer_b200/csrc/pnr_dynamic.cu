// K2: the fused env step of the dynamic (Tier-B) mode.  One thread per env: frame_skip substeps of
// [PD / torque control -> Featherstone ABA -> semi-implicit Euler -> limit stops] with the joint state in
// registers, then the same task code as the kinematic env (pioneer_knm_env.py:151-211): forward kinematics of
// the pointer, reward, done, TimeLimit, statistics, auto-reset, 137-float observation staged per warp in shared
// memory and stored with one TMA bulk copy.  Bound: FP32 pipe (10 x ~1.45 kflop per env-step against 761 B).
#include <cstdlib>
#include "pnr_kernels.cuh"
#include "pnr_dynamics.cuh"
#include "pnr_launch.h"

#ifndef PNR_DYN_MIN_CTAS
// Resident CTAs per SM, measured on B200 at 1,048,576 envs.  First ABA (223 registers): 2 CTAs 72 us per 65,536-env step
// against 93 us with 3 (spills).  Sparsity-aware ABA (183 registers at 2 CTAs, 152 at 3, no spills): 486 us with 2,
// 461 us with 3.
#define PNR_DYN_MIN_CTAS 3
#endif
template <int OBS_MODE, bool OBSTACLES, int CHAIN>
__global__ void __launch_bounds__(PNR_STEP_THREADS, PNR_DYN_MIN_CTAS)
pnr_step_dynamic_kernel(const __grid_constant__ PnrParams p, float4* __restrict__ state, const float* __restrict__ actions,
                        float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                        PnrStats* __restrict__ stats, uint32_t tick, uint32_t domain, const float* __restrict__ f_applied,
                        double* __restrict__ f_delta, float f_clip) {
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = smem + warp * PNR_TILE_FLOATS;
    float* row = tile + lane * PNR_OBS_DIM;
    const int64_t N = p.n_envs;
    const int64_t n_tiles = (N + PNR_TILE_ENVS - 1) / PNR_TILE_ENVS;
    bool tile_busy = false;
    pnr_pdl_trigger();                                         // see pnr_step_kernel: the next step's set-up runs under this tail
    pnr_pack_obs_const(p, row);
    // fused observation normaliser (pnr_filter_fuse; warp-uniform run-time switch): this warp owns the whole tile, so
    // after the rows are packed it runs the column pass of pnr_filter_kernel on the 101 changing columns (lane = column,
    // four passes) and pushes the tile's float64 column sums to accumulator copy blockIdx % SLOTS.  The 36 constant
    // columns are normalised here, once, and their statistics added analytically by one thread of the grid.
    const bool filt = f_applied != nullptr;
    if (filt) {
#pragma unroll
        for (int c = 18; c < 54; ++c) row[c] = pnr_normalise(row[c], f_applied[c], f_applied[PNR_OBS_DIM + c], f_clip);
    }
    pnr_pdl_wait();                                            // the previous step's state planes are complete and visible
    for (int64_t t_idx = (int64_t)blockIdx.x * PNR_STEP_WARPS + warp; t_idx < n_tiles;
         t_idx += (int64_t)gridDim.x * PNR_STEP_WARPS) {
        const int64_t env_raw = t_idx * PNR_TILE_ENVS + lane;
        const bool active = env_raw < N;
        const int64_t env = active ? env_raw : N - 1;

        PnrEnv s;
        pnr_load_env(state, N, env, s);
        const float2* a2 = reinterpret_cast<const float2*>(actions + env * PNR_DOF);
        const float2 act01 = pnr_ld_stream(a2), act23 = pnr_ld_stream(a2 + 1), act45 = pnr_ld_stream(a2 + 2);
        // the action drives THIS step's substeps (a motor target, not the kinematic env's delayed acceleration)
        s.a[0] = act01.x; s.a[1] = act01.y; s.a[2] = act23.x; s.a[3] = act23.y; s.a[4] = act45.x; s.a[5] = act45.y;
        pnr_dynamic_substeps<CHAIN>(p, s.r, s.v, s.a);

        PnrPose o;
        pnr_pose<true>(p, s, o);                               // q is inside the joint limits
        bool reached = o.dist < p.done_distance;
        if (fabsf(o.dist - p.done_distance) < p.done_band)
            pnr_fk_tip_f64(p, s.r, s.tgt, o.ptr, o.dist, reached);
        const float pot_new = pnr_potential(p, o.dist);
        float rew = __fadd_rn(__fadd_rn(__fsub_rn(pot_new, s.pot), -p.penalty_step), reached ? p.award_done : 0.f);
        if (OBSTACLES) {
            PnrSinCos sc;
#pragma unroll
            for (int i = 0; i < PNR_DOF; ++i) { sc.sn[i] = o.sn[i]; sc.cs[i] = o.cs[i]; }
            rew = __fsub_rn(rew, __fmul_rn(p.contact_penalty, pnr_contact_depth(p, sc, pnr_load_box(p, env))));
        }
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
        pnr_episode_stats(stats, is_done && active, reached && active, s.ep_ret, s.t, lane);

        if (tile_busy) pnr_tile_wait(lane);
        float vmax_abs = 0.f;                                  // joint rates are not bounded by construction here
#pragma unroll
        for (int i = 0; i < PNR_DOF; ++i) vmax_abs = fmaxf(vmax_abs, fabsf(s.v[i]));
        const bool slow = !(vmax_abs <= PNR_TRIG_FAST_LIMIT);
        const bool do_reset = is_done && (p.auto_reset != 0);
        if (OBS_MODE == PNR_OBS_TERMINAL) pnr_pack_obs_dyn<true>(p, row, s, o, s.pot, slow);
        if (do_reset) {
            float q[PNR_DOF], tg[3], box[5];
            pnr_reset_draws(p, p.env_id_base + env, pnr_reset_key(p, pnr_tickdom(tick, domain)), q, tg, box);
            pnr_reset_env(s, q, tg);
            if (active) pnr_store_box(p, env, box);
        }
        if (OBS_MODE == PNR_OBS_AUTORESET) {
            if (__any_sync(PNR_FULL_MASK, do_reset)) {
                PnrPose o2;
                pnr_pose<true>(p, s, o2);
                if (do_reset) o = o2;
            }
            pnr_pack_obs_dyn<true>(p, row, s, o, s.pot, slow && !do_reset);
        }
        if (active) pnr_store_env(state, N, env, s);
        const int64_t rows_left = N - t_idx * PNR_TILE_ENVS;
        const int rows_valid = rows_left < PNR_TILE_ENVS ? (int)rows_left : PNR_TILE_ENVS;
        if (filt) {
            __syncwarp();
            double* f_slot = f_delta ? f_delta + (size_t)(blockIdx.x & (PNR_FILTER_SLOTS - 1)) * PNR_FILTER_DELTA_LEN : nullptr;
#pragma unroll 1
            for (int pass = 0; pass < 4; ++pass) {
                const int cc = pass * 32 + lane;                   // index among the 101 changing columns
                if (cc < PNR_OBS_DIM - 36) {
                    const int c = cc < 18 ? cc : cc + 36;
                    const float mean = f_applied[c], inv = f_applied[PNR_OBS_DIM + c];
                    float* colp = tile + c;
                    double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0;
                    int r = 0;
#pragma unroll 4
                    for (; r + 1 < rows_valid; r += 2) {
                        const float d0 = colp[r * PNR_OBS_DIM] - mean, d1 = colp[(r + 1) * PNR_OBS_DIM] - mean;
                        const double e0 = (double)d0, e1 = (double)d1;
                        s0 += e0; q0 = fma(e0, e0, q0);
                        s1 += e1; q1 = fma(e1, e1, q1);
                        colp[r * PNR_OBS_DIM] = fminf(fmaxf(d0 * inv, -f_clip), f_clip);
                        colp[(r + 1) * PNR_OBS_DIM] = fminf(fmaxf(d1 * inv, -f_clip), f_clip);
                    }
                    if (r < rows_valid) {
                        const float d0 = colp[r * PNR_OBS_DIM] - mean;
                        const double e0 = (double)d0;
                        s0 += e0; q0 = fma(e0, e0, q0);
                        colp[r * PNR_OBS_DIM] = fminf(fmaxf(d0 * inv, -f_clip), f_clip);
                    }
                    if (f_slot) {
                        atomicAdd(&f_slot[1 + c], s0 + s1);
                        atomicAdd(&f_slot[1 + PNR_OBS_DIM + c], q0 + q1);
                    }
                }
            }
        }
        pnr_emit_tile(tile, obs + t_idx * (int64_t)PNR_TILE_FLOATS, rows_valid, lane);
        tile_busy = true;
    }
    if (filt && f_delta && blockIdx.x == 0 && warp == 0) {         // rows pushed + the constant columns of all N rows
        if (lane == 0) atomicAdd(&f_delta[0], (double)N);
        for (int c = 18 + lane; c < 54; c += 32) {
            const int g = (c - 18) / PNR_DOF, j = (c - 18) % PNR_DOF;
            const float x = g == 0 ? p.r_lo[j] : g == 1 ? p.cos_r_lo[j] : g == 2 ? p.sin_r_lo[j]
                          : g == 3 ? p.r_hi[j] : g == 4 ? p.cos_r_hi[j] : p.sin_r_hi[j];
            const double d = (double)(x - f_applied[c]);
            atomicAdd(&f_delta[1 + c], (double)N * d);
            atomicAdd(&f_delta[1 + PNR_OBS_DIM + c], (double)N * d * d);
        }
    }
    if (blockIdx.x == 0 && warp == 0 && lane == 0) stats->env_steps += (double)N;   // one writer per launch
    if (lane == 0) pnr_bulk_wait_read<0>();
}

static int pnr_dyn_resident(const void* fn) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PNR_STEP_THREADS, PNR_RO_SMEM);
    return sms * (per_sm < 1 ? 1 : per_sm);
}

cudaError_t pnr_launch_step_dynamic(const PnrParams& p, int device, int obs_mode, float4* state, const float* actions,
                                    float* obs, float* reward, uint8_t* done, PnrStats* stats, uint32_t tick, uint32_t domain,
                                    const float* f_applied, double* f_delta, float f_clip, cudaStream_t stream) {
    typedef void (*Kern)(const PnrParams, float4*, const float*, float*, float*, uint8_t*, PnrStats*, uint32_t, uint32_t,
                         const float*, double*, float);
#define PNR_DYN_ROW(CH) \
    {{pnr_step_dynamic_kernel<PNR_OBS_TERMINAL, false, CH>, pnr_step_dynamic_kernel<PNR_OBS_TERMINAL, true, CH>}, \
     {pnr_step_dynamic_kernel<PNR_OBS_AUTORESET, false, CH>, pnr_step_dynamic_kernel<PNR_OBS_AUTORESET, true, CH>}}
    static Kern kernels[3][2][2] = {PNR_DYN_ROW(PNR_CHAIN_GENERIC), PNR_DYN_ROW(PNR_CHAIN_PIONEER),
                                    PNR_DYN_ROW(PNR_CHAIN_PIONEER_ISO)};
    static int grids[PNR_MAX_DEVICES][3][2][2] = {};
    const int obst = p.n_obstacles > 0 ? 1 : 0;
    int chain = p.chain_kind == 1 ? (p.dyn_iso_links ? PNR_CHAIN_PIONEER_ISO : PNR_CHAIN_PIONEER) : PNR_CHAIN_GENERIC;
    if (const char* e = getenv("PNR_DYN_CHAIN")) chain = atoi(e) < chain ? atoi(e) : chain;   // developer / test knob: force a less specialised kernel
    Kern kern = kernels[chain][obs_mode][obst];
    int& resident = grids[device % PNR_MAX_DEVICES][chain][obs_mode][obst];
    if (resident == 0) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PNR_RO_SMEM);
        if (e != cudaSuccess) return e;
        resident = pnr_dyn_resident((const void*)kern);
    }
    const int64_t per_cta = PNR_TILE_ENVS * PNR_STEP_WARPS;
    int64_t grid = (p.n_envs + per_cta - 1) / per_cta;
    if (grid > resident) grid = resident;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(PNR_STEP_THREADS); cfg.dynamicSmemBytes = PNR_RO_SMEM; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pnr_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p, state, actions, obs, reward, done, stats, tick, domain, f_applied, f_delta, f_clip);
}
