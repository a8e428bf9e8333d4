// Kernels of the kinematic reach env (Tier A): fused step, reset, observe, state get/set, stats.
// sm_100a; launched from pnr_api.cu.
#include "pnr_kernels.cuh"
#include "pnr_launch.h"
#include <cstdlib>

// ---------------------------------------------------------------------------------------------
// K1: the fused env step.  One CTA of 4 warps owns a tile of 32 envs (lane = env); the work of ONE env is
// split across the 4 warps so that its dependent instruction chain is ~4x shorter and 4x more warps are
// resident per byte of shared memory than with one thread per env:
//   warps 0..2 ("joint warps") : joints 2k, 2k+1 -- act() integrator (pioneer_knm_env.py:111-146), their
//                                21 + 21 observation columns incl. 10 (sin, cos) pairs, state planes RV_k, A_k
//   warp 3     ("task warp")   : forward kinematics of the pointer from the joint warps' sin/cos (read back
//                                from the tile), reward / done / TimeLimit (:151-165), episode statistics,
//                                auto-reset (reset_world, :76-105), observation tail, planes X0, X1,
//                                reward / done outputs, and the TMA bulk store of the finished tile
// The two roles run as a producer / consumer pipeline over TWO tile buffers, coupled only by named barriers:
//   HEAD[b]  joint warps arrive once r, cos r, sin r of their joints are in buffer b; the task warp waits on it
//   DONE[b]  joint warps arrive once all their columns and planes are written; the task warp waits, then stores
//   FREE[b]  the task warp arrives once the bulk store of buffer b has been read; joint warps wait before reuse
//   JOINT    (OBSTACLES only) the three joint warps among themselves, once their heads are in the tile: each then sums the
//            penetration depths of its share of the capsule table into shared memory; the task warp adds the three partials
//            behind DONE and only then writes reward, return and episode statistics
// so the joint warps of a CTA work on tile i+1 while its task warp finishes tile i (before: 49 % of stall samples
// were joint warps parked at a CTA-wide barrier).  Each warp prefetches its own planes of the CTA's next tile.
// Grid = min(tiles, resident CTAs of the whole GPU); grid-stride over tiles.
// ---------------------------------------------------------------------------------------------
// In-kernel auto-reset of ONE env by the task warp after the tile is complete: new joint angles and target from
// the Philox stream, a = v = 0, potential = 0, elapsed = 0 written over all eight planes of the env; in
// PNR_OBS_AUTORESET mode the env's observation row is replaced by the first observation of the new episode.
// Rare (once per episode), so deliberately not inlined: keeps registers and code out of the hot loop.
// In-kernel auto-reset of ONE env by the task warp after the tile is complete: new joint angles and target from
// the Philox stream, a = v = 0, potential = 0, elapsed = 0 written over all eight planes of the env; in
// PNR_OBS_AUTORESET mode the env's observation row is replaced by the first observation of the new episode.
// Rare (once per episode), so deliberately not inlined: keeps registers and code out of the hot loop.
// (Tried and dropped: drawing the next episode as soon as the TimeLimit hit is known, before the task warp waits for the
// joint warps -- the divergent Philox lane delays the whole task warp, which IS the critical path there: 11.8 -> 12.9 us
// per 65,536-env step.)
template <int OBS_MODE>
__device__ __noinline__ void pnr_auto_reset(const PnrParams& p, float4* __restrict__ state, int64_t env, uint64_t tickdom,
                                            float* __restrict__ row) {
    const int64_t N = p.n_envs;
    PnrEnv s;
    float q[PNR_DOF], tg[3], box[5];
    pnr_reset_draws(p, p.env_id_base + env, pnr_reset_key(p, tickdom), q, tg, box);
    pnr_reset_env(s, q, tg);
    pnr_store_env(state, N, env, s);
    pnr_store_box(p, env, box);
    if (OBS_MODE == PNR_OBS_AUTORESET) {
        PnrPose o;
        pnr_pose<true>(p, s, o);                              // fresh joint angles lie inside the limits
        pnr_pack_obs_dyn<true>(p, row, s, o, s.pot);
    }
}

#ifdef PNR_TRACE
// developer instrumentation (python -m pioneer_b200.build --trace): per-CTA phase timestamps of the first tile
#define PNR_TRACE_SLOTS 12
#define PNR_TRACE_CTAS 2048
#define PNR_TRACE_ITERS 4
__device__ unsigned long long pnr_trace_buf[PNR_TRACE_CTAS][2][PNR_TRACE_ITERS][PNR_TRACE_SLOTS];
__device__ __forceinline__ void pnr_trace_mark(int part, int lane, int slot, int iter) {
    if (lane == 0 && (part == 0 || part == 3) && blockIdx.x < PNR_TRACE_CTAS && iter < PNR_TRACE_ITERS) {
        unsigned long long t;
        if (slot == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); else t = clock64();
        pnr_trace_buf[blockIdx.x][part == 3][iter][slot] = t;
    }
}
extern "C" int pnr_debug_trace(unsigned long long* out_host) {
    return (int)cudaMemcpyFromSymbol(out_host, pnr_trace_buf, sizeof(pnr_trace_buf));
}
#define PNR_MARK(slot) pnr_trace_mark(part, lane, slot, trace_iter)
#define PNR_TRACE_NEXT() (++trace_iter)
#else
#define PNR_MARK(slot)
#define PNR_TRACE_NEXT()
#endif

// FILTER = true fuses the observation normaliser (pnr_filter.cu, 'ConcurrentMeanStdFilter') into the step: what leaves the
// kernel is clip((x - mean) * inv_std), and the column statistics of the raw values go to the same float64 accumulator
// pnr_filter_apply feeds, so the 548 B/env observation is never re-read.  Every warp normalises the columns it owns
// while the tile is still in shared memory: a joint warp runs one column pass (lane = one of its 30 changing
// columns, 32 rows, float64 sum / sum of squares of x - mean), the task warp normalises its 11 tail values in
// registers and keeps per-lane float partial sums; the 36 constant columns are normalised once per CTA and their
// statistics are added analytically by block 0.  The raw r / cos r / sin r the task warp needs for the forward
// kinematics travel through a small scratch area next to the tile buffers (the tile copy gets normalised).
template <int ARITH, int OBS_MODE, bool OBSTACLES, bool FILTER>
__global__ void __launch_bounds__(PNR_STEP_THREADS, FILTER ? PNR_STEP_MIN_CTAS_FILTER : PNR_STEP_MIN_CTAS)
pnr_step_kernel(const __grid_constant__ PnrParams p, float4* __restrict__ state, const float* __restrict__ actions,
                float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                PnrStats* __restrict__ stats, uint32_t tick, uint32_t domain, const float* __restrict__ f_applied,
                double* __restrict__ f_delta, float f_clip, const PnrMulti multi) {
    extern __shared__ __align__(128) float tiles[];           // PNR_STEP_BUFS tiles: the bulk store of one drains
    const int lane = threadIdx.x & 31;                        // while the next is being filled
    const int part = threadIdx.x >> 5;                        // warp-uniform role
    float* tile = tiles;
    float* row = tile + lane * PNR_OBS_DIM;
    float* scratch = tiles + PNR_STEP_BUFS * PNR_TILE_FLOATS; // FILTER only: raw r, cos r, sin r per env, per buffer
    float* srow = scratch + lane * PNR_FSCRATCH_STRIDE;
    // OBSTACLES only: partial contact depths of the three joint warps, per buffer [joint warp][env]
    float* const pen = scratch + (FILTER ? PNR_STEP_BUFS * PNR_FSCRATCH_FLOATS : 0);
    const int64_t N = p.n_envs;
    const int64_t n_tiles = (N + PNR_TILE_ENVS - 1) / PNR_TILE_ENVS;
    const int64_t stride = gridDim.x;
    int64_t t_idx = blockIdx.x;
    if (t_idx >= n_tiles) return;                             // CTA-uniform
    // back-to-back steps (pnr_step_many, CUDA graphs): the next step's CTAs may take the SM slots this grid's CTAs free one
    // by one and run their set-up (constant columns, barriers) under this grid's tail; they stop at pnr_pdl_wait()
    pnr_pdl_trigger();

#ifdef PNR_TRACE
    int trace_iter = 0;
#endif
    PNR_MARK(0); PNR_MARK(1);
    const int j0 = part < 3 ? 2 * part : 0;                   // first joint of a joint warp
    float* rowj = row + j0;
    // loop-invariant limits of this warp's two joints, read once from the constant bank
    const float vmax0 = p.v_max[j0], vmax1 = p.v_max[j0 + 1];
    const float rlo0 = p.r_lo[j0], rlo1 = p.r_lo[j0 + 1], rhi0 = p.r_hi[j0], rhi1 = p.r_hi[j0 + 1];
    float4* const rv_plane = pnr_plane_rv(state, N, part < 3 ? part : 0);
    float2* const a_plane = pnr_plane_a(state, N, part < 3 ? part : 0);
    float4* const x0_plane = pnr_plane_x0(state, N);
    float2* const x1_plane = pnr_plane_x1(state, N);

    // obs[18:54] never change: written once per CTA (into every tile buffer)
    if (part < 3) {
#pragma unroll
        for (int b = 0; b < PNR_STEP_BUFS; ++b) {
            pnr_pack_joint_const(p, row + b * PNR_TILE_FLOATS, j0);
            pnr_pack_joint_const(p, row + b * PNR_TILE_FLOATS, j0 + 1);
        }
    }
    // fused normaliser: the column this lane owns in the column pass, its applied statistics and accumulators
    int f_col = 0;
    float f_mean = 0.f, f_inv = 0.f;
    double f_sum = 0.0, f_sq = 0.0;
    float t_sum[11], t_sq[11];
    if (FILTER) {
        if (part < 3) {
            // lanes 0..5 -> column blocks 0..2 (r, cos r, sin r), lanes 6..29 -> blocks 9..20; blocks 3..8 are constant
            const int kk = lane < 6 ? (lane >> 1) : (lane >> 1) + 6;
            f_col = 6 * kk + j0 + (lane & 1);
            if (lane < 30) { f_mean = f_applied[f_col]; f_inv = f_applied[PNR_OBS_DIM + f_col]; }
#pragma unroll
            for (int g = 0; g < 6; ++g) {                      // normalise the constant columns in place
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int c = 18 + 6 * g + j0 + jj;
                    const float y = pnr_normalise(row[c], f_applied[c], f_applied[PNR_OBS_DIM + c], f_clip);
#pragma unroll
                    for (int b = 0; b < PNR_STEP_BUFS; ++b) row[b * PNR_TILE_FLOATS + c] = y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 11; ++i) { t_sum[i] = 0.f; t_sq[i] = 0.f; }
        }
    }

    // software pipeline: every warp loads its own planes of the NEXT tile while it works on the current one
    float4 ld4; float2 ld2, ld_act = make_float2(0.f, 0.f);
    // (state planes through L2 only: with programmatic dependent launch this CTA may have started on an SM before the previous
    // step finished elsewhere; nothing is re-read, so L1 would buy nothing)
    auto issue_loads = [&](int64_t tile_idx, const float* __restrict__ act) {
        const int64_t e_raw = tile_idx * PNR_TILE_ENVS + lane;
        const int64_t e = e_raw < N ? e_raw : N - 1;
        if (part < 3) {
            ld4 = __ldcg(rv_plane + e);
            ld2 = __ldcg(a_plane + e);
            ld_act = pnr_ld_stream(reinterpret_cast<const float2*>(act + e * PNR_DOF) + part);
        } else {
            ld4 = __ldcg(x0_plane + e);
            ld2 = __ldcg(x1_plane + e);
        }
    };
    pnr_pdl_wait();                                           // the previous step's state planes are complete and visible
    issue_loads(t_idx, actions);
    PNR_MARK(2);
    int buf = 0;
    int64_t iter = 0;                                         // tiles this CTA has started

    // n_steps consecutive steps on this CTA's own tiles (PnrMulti; 1 for an ordinary step).  Between two steps the warps of
    // the CTA meet at a barrier: the task warp may have reset an env of the tile, which rewrites planes the joint warps
    // are about to load again.
    for (int32_t step = 0; step < multi.n_steps; ++step) {
    const bool more_steps = step + 1 < multi.n_steps;
    // this step's buffers (launch-uniform: the bases stay in the constant bank)
    const float* __restrict__ const act_s = actions + (int64_t)step * multi.act_stride;
    float* __restrict__ const obs_s = obs + (int64_t)step * multi.obs_stride;
    float* __restrict__ const reward_s = reward + (int64_t)step * N;
    uint8_t* __restrict__ const done_s = done + (int64_t)step * N;
    const uint32_t tick_s = tick + (uint32_t)step;
    if (step > 0) {
        __syncthreads();
        t_idx = blockIdx.x;
        issue_loads(t_idx, act_s);
    }
    for (; t_idx < n_tiles; t_idx += stride) {
        const int64_t env_raw = t_idx * PNR_TILE_ENVS + lane;
        const bool active = env_raw < N;
        const int64_t env = active ? env_raw : N - 1;       // tail lanes shadow the last env, stores masked
        const int64_t rows_left = N - t_idx * PNR_TILE_ENVS;
        const int rows_valid = rows_left < PNR_TILE_ENVS ? (int)rows_left : PNR_TILE_ENVS;
        const float4 c4 = ld4;
        const float2 c2 = ld2, c_act = ld_act;
        if (t_idx + stride < n_tiles) issue_loads(t_idx + stride, act_s);

        float r1[2], v1[2], sn[2], cs[2];
        if (part < 3) {
            // --- act(): integrate with the PREVIOUS action (one-step actuation delay); registers only
            pnr_integrate_joint<ARITH>(p, vmax0, rlo0, rhi0, c2.x, c4.z, c4.x, v1[0], r1[0]);
            pnr_integrate_joint<ARITH>(p, vmax1, rlo1, rhi1, c2.y, c4.w, c4.y, v1[1], r1[1]);
            pnr_sincos_fast(r1[0], sn[0], cs[0]);             // r is inside the joint limits: fast path
            pnr_sincos_fast(r1[1], sn[1], cs[1]);
            PNR_MARK(3);
            if (iter >= PNR_STEP_BUFS) pnr_bar_sync2<PNR_BAR_FREE, PNR_STEP_THREADS>(buf);   // buffer reusable
            PNR_MARK(4);
            pnr_pack_joint_head(rowj, r1[0], sn[0], cs[0]);
            pnr_pack_joint_head(rowj + 1, r1[1], sn[1], cs[1]);
            if (FILTER) {                                      // the raw values for the task warp's forward kinematics
                pnr_pack_joint_head(srow + j0, r1[0], sn[0], cs[0]);
                pnr_pack_joint_head(srow + j0 + 1, r1[1], sn[1], cs[1]);
            }
            pnr_bar_arrive2<PNR_BAR_HEAD, PNR_STEP_THREADS>(buf);
        } else {
            PNR_MARK(3);
            pnr_bar_sync2<PNR_BAR_HEAD, PNR_STEP_THREADS>(buf);   // r, cos r, sin r of all six joints are in the tile
            PNR_MARK(4);
        }
        PNR_MARK(5);

        bool do_reset = false;
        // OBSTACLES: what the task warp carries across the DONE barrier (the penalty is summed from the joint warps' partials)
        float d_rew = 0.f, d_ep = 0.f, d_pot = 0.f;
        int32_t d_t = 0;
        bool d_done = false, d_reached = false;
        if (part < 3) {
            // --- the other 15 dynamic columns of each joint; latch the new action (stored unclipped, :144)
            const bool fast = !p.trig_slow && fmaxf(fabsf(c_act.x), fabsf(c_act.y)) <= PNR_TRIG_FAST_LIMIT;
            pnr_pack_joint_rest(rowj, rlo0, rhi0, r1[0], v1[0], c_act.x, fast);
            pnr_pack_joint_rest(rowj + 1, rlo1, rhi1, r1[1], v1[1], c_act.y, fast);
            if (active) {
                rv_plane[env] = make_float4(r1[0], r1[1], v1[0], v1[1]);
                a_plane[env] = c_act;
            }
            if (OBSTACLES) {
                // capsule penetration depths, the capsule table split over the three joint warps (on the task warp alone
                // they were the CTA's critical path: 59 % of the stall samples were joint warps parked at DONE).  Every
                // joint's sin/cos is needed: the joint warps meet once their heads are in the tile.
                pnr_bar_sync<PNR_BAR_JOINT, 3 * 32>();
                const float* head = FILTER ? srow : row;
                PnrSinCos sc;
#pragma unroll
                for (int i = 0; i < PNR_DOF; ++i) { sc.cs[i] = head[6 + i]; sc.sn[i] = head[12 + i]; }
                pen[(buf * 3 + part) * 32 + lane] = pnr_contact_depth(p, sc, pnr_load_box(p, env), part, 3);
            }
            if (FILTER) {                                      // column pass over this warp's 30 changing columns
                __syncwarp();
                if (lane < 30) {
                    float* colp = tile + f_col;
                    double s1 = 0.0, q1 = 0.0;                 // second accumulator pair: two independent FP64 chains
                    int r = 0;
#pragma unroll 4
                    for (; r + 1 < rows_valid; r += 2) {
                        const float d0 = colp[r * PNR_OBS_DIM] - f_mean, d1 = colp[(r + 1) * PNR_OBS_DIM] - f_mean;
                        const double e0 = (double)d0, e1 = (double)d1;
                        f_sum += e0; f_sq = fma(e0, e0, f_sq);
                        s1 += e1; q1 = fma(e1, e1, q1);
                        colp[r * PNR_OBS_DIM] = fminf(fmaxf(d0 * f_inv, -f_clip), f_clip);
                        colp[(r + 1) * PNR_OBS_DIM] = fminf(fmaxf(d1 * f_inv, -f_clip), f_clip);
                    }
                    if (r < rows_valid) {
                        const float d0 = colp[r * PNR_OBS_DIM] - f_mean;
                        const double e0 = (double)d0;
                        f_sum += e0; f_sq = fma(e0, e0, f_sq);
                        colp[r * PNR_OBS_DIM] = fminf(fmaxf(d0 * f_inv, -f_clip), f_clip);
                    }
                    f_sum += s1; f_sq += q1;
                }
            }
        } else {
            // --- pose from the joint warps' sin/cos, distance, reward, done
            PnrPose o;
            const float* head = FILTER ? srow : row;           // raw r | cos r | sin r of this env
#pragma unroll
            for (int i = 0; i < PNR_DOF; ++i) { o.cs[i] = head[6 + i]; o.sn[i] = head[12 + i]; }
            const float tgt[3] = {c4.x, c4.y, c4.z};
            int32_t t = __float_as_int(c4.w);
            const float pot_old = c2.x;
            float ep_ret = c2.y;
            pnr_fk_tip(p, o.sn, o.cs, o.ptr);
            const float dx = tgt[0] - o.ptr[0], dy = tgt[1] - o.ptr[1], dz = tgt[2] - o.ptr[2];
            o.dist = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            bool reached = o.dist < p.done_distance;
            if (fabsf(o.dist - p.done_distance) < p.done_band) {   // decide in float64 where float32 could flip it
                PnrBandIn in;
#pragma unroll
                for (int i = 0; i < PNR_DOF; ++i) in.r[i] = head[i];
                in.tgt[0] = tgt[0]; in.tgt[1] = tgt[1]; in.tgt[2] = tgt[2];
                const PnrBandOut bo = pnr_fk_band_f64(p, in);
                o.ptr[0] = bo.ptr[0]; o.ptr[1] = bo.ptr[1]; o.ptr[2] = bo.ptr[2];
                o.dist = bo.dist;
                reached = bo.within != 0;
            }
            const float pot_new = pnr_potential(p, o.dist);
            // (potential - old_potential) + (-penalty_step) + (award_done | 0), pioneer_knm_env.py:162-165
            float rew = __fadd_rn(__fadd_rn(__fsub_rn(pot_new, pot_old), -p.penalty_step),
                                  reached ? p.award_done : 0.f);
            t += 1;
            const bool timeout = p.max_episode_steps > 0 && t >= p.max_episode_steps;
            const bool is_done = reached || timeout;
            const uint8_t flags = (is_done ? PNR_DONE : 0) | ((timeout && !reached) ? PNR_TRUNCATED : 0);
            if (active) done_s[env] = flags;
            if (OBSTACLES) {                                   // the penalty arrives with DONE: reward, return and statistics wait
                d_rew = rew; d_ep = ep_ret; d_pot = pot_new; d_t = t; d_done = is_done; d_reached = reached;
            } else {
                ep_ret = __fadd_rn(ep_ret, rew);
                if (active) reward_s[env] = rew;
                pnr_episode_stats(stats, is_done && active, reached && active, ep_ret, t, lane);
            }

            // observation tail: pointer, target, difference, distance, potential (terminal values)
            float tail[11] = {o.ptr[0], o.ptr[1], o.ptr[2], tgt[0], tgt[1], tgt[2],
                              tgt[0] - o.ptr[0], tgt[1] - o.ptr[1], tgt[2] - o.ptr[2], o.dist, pot_new};
            if (FILTER) {                                      // normalise in registers, per-lane partial statistics
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const float d = tail[i] - f_applied[126 + i];
                    if (active) { t_sum[i] += d; t_sq[i] = fmaf(d, d, t_sq[i]); }
                    tail[i] = fminf(fmaxf(d * f_applied[PNR_OBS_DIM + 126 + i], -f_clip), f_clip);
                }
            }
#pragma unroll
            for (int i = 0; i < 11; ++i) row[126 + i] = tail[i];

            if (active) {
                x0_plane[env] = make_float4(tgt[0], tgt[1], tgt[2], __int_as_float(t));
                if (!OBSTACLES) x1_plane[env] = make_float2(pot_new, ep_ret);
            }
            do_reset = is_done && active && (p.auto_reset != 0);   // handled after B2 (rare)
        }
        PNR_MARK(6);
        pnr_fence_async_smem();                               // generic-proxy tile writes -> async proxy
        const bool bulk = pnr_tile_is_bulk(rows_valid);       // CTA-uniform
        if (!bulk) __syncthreads();                           // ragged last tile of the grid: everybody streams it out
        else if (part < 3) pnr_bar_arrive2<PNR_BAR_DONE, PNR_STEP_THREADS>(buf);
        else pnr_bar_sync2<PNR_BAR_DONE, PNR_STEP_THREADS>(buf);   // tile complete, joint planes stored
        PNR_MARK(7);

        if (OBSTACLES && part == 3) {                         // obstacle variant: capsule penetration penalty, then the deferred writes
            const float* pb = pen + buf * PNR_PEN_FLOATS + lane;
            const float depth = (pb[0] + pb[32]) + pb[64];
            const float rew = __fsub_rn(d_rew, __fmul_rn(p.contact_penalty, depth));
            const float ep_ret = __fadd_rn(d_ep, rew);
            if (active) {
                reward_s[env] = rew;
                x1_plane[env] = make_float2(d_pot, ep_ret);
            }
            pnr_episode_stats(stats, d_done && active, d_reached && active, ep_ret, d_t, lane);
        }
        if (part == 3) {
            // rare: auto-reset (reset_world, :76-105).  Only the auto-reset observation mode rewrites the row, so only there
            // does the reset have to precede the tile's store; terminal mode resets AFTER handing the tile to the copy engine
            const bool any_reset = __any_sync(PNR_FULL_MASK, do_reset);
            if (OBS_MODE == PNR_OBS_AUTORESET && any_reset) {
                if (do_reset) pnr_auto_reset<OBS_MODE>(p, state, env, pnr_tickdom(tick_s, domain), row);
                pnr_fence_async_smem();
                __syncwarp();
            }
            if (bulk) {
                if (lane == 0) {
                    pnr_bulk_store(obs_s + t_idx * (int64_t)PNR_TILE_FLOATS, tile,
                                   (uint32_t)(rows_valid * PNR_OBS_DIM * sizeof(float)));
                    pnr_bulk_commit();
                    // every store but the one just committed has finished READING its buffer: the previous tile's
                    // buffer is free again (only signalled if a tile that will reuse it exists)
                    if (iter >= 1) pnr_bulk_wait_read<1>();
                }
                __syncwarp();
                if (iter >= 1 && (t_idx + stride < n_tiles || more_steps)) pnr_bar_arrive2<PNR_BAR_FREE, PNR_STEP_THREADS>(buf ^ 1);
            }
            if (OBS_MODE != PNR_OBS_AUTORESET && any_reset && do_reset)
                pnr_auto_reset<OBS_MODE>(p, state, env, pnr_tickdom(tick_s, domain), row);
        }
        PNR_MARK(8);
        PNR_TRACE_NEXT();
        if (!bulk) {                                          // ragged last tile: all 128 threads stream it out
            if (part == 3 && iter >= 1 && more_steps) {       // (another step follows: keep the buffer protocol going)
                if (lane == 0) pnr_bulk_wait_read<0>();
                __syncwarp();
                pnr_bar_arrive2<PNR_BAR_FREE, PNR_STEP_THREADS>(buf ^ 1);
            }
            __syncthreads();
            pnr_emit_tile_manual(tile, obs_s + t_idx * (int64_t)PNR_TILE_FLOATS, rows_valid, threadIdx.x, PNR_STEP_THREADS);
        }
        buf ^= 1;                                             // next tile goes into the other buffer
        ++iter;
        tile = tiles + buf * PNR_TILE_FLOATS;
        row = tile + lane * PNR_OBS_DIM;
        rowj = row + j0;
        if (FILTER) srow = scratch + buf * PNR_FSCRATCH_FLOATS + lane * PNR_FSCRATCH_STRIDE;
    }
    }                                                         // steps
    if (FILTER && f_delta != nullptr) {                       // column statistics of this CTA's rows, into copy blockIdx % SLOTS
        double* f_slot = f_delta + (size_t)(blockIdx.x & (PNR_FILTER_SLOTS - 1)) * PNR_FILTER_DELTA_LEN;
        if (part < 3) {
            if (lane < 30) {
                atomicAdd(&f_slot[1 + f_col], f_sum);
                atomicAdd(&f_slot[1 + PNR_OBS_DIM + f_col], f_sq);
            }
            if (blockIdx.x == 0 && lane < 12) {                // the constant columns of all N rows, analytically
                const int g = lane >> 1, j = j0 + (lane & 1), c = 18 + 6 * g + j;
                const float x = g == 0 ? p.r_lo[j] : g == 1 ? p.cos_r_lo[j] : g == 2 ? p.sin_r_lo[j]
                              : g == 3 ? p.r_hi[j] : g == 4 ? p.cos_r_hi[j] : p.sin_r_hi[j];
                const double d = (double)(x - f_applied[c]);
                atomicAdd(&f_delta[1 + c], (double)N * multi.n_steps * d);
                atomicAdd(&f_delta[1 + PNR_OBS_DIM + c], (double)N * multi.n_steps * d * d);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 11; ++i) {
                float a = t_sum[i], b = t_sq[i];
#pragma unroll
                for (int ofs = 16; ofs > 0; ofs >>= 1) {
                    a += __shfl_xor_sync(PNR_FULL_MASK, a, ofs);
                    b += __shfl_xor_sync(PNR_FULL_MASK, b, ofs);
                }
                if (lane == 0) {
                    atomicAdd(&f_slot[1 + 126 + i], (double)a);
                    atomicAdd(&f_slot[1 + PNR_OBS_DIM + 126 + i], (double)b);
                }
            }
            if (blockIdx.x == 0 && lane == 0) atomicAdd(&f_delta[0], (double)N * multi.n_steps);
        }
    }
    if (part == 3 && lane == 0 && blockIdx.x == 0) atomicAdd(&stats->env_steps, (double)N * multi.n_steps);   // one writer per launch
    if (part == 3 && lane == 0) pnr_bulk_wait_read<0>();      // smem must outlive the copy engine's reads
#ifdef PNR_TRACE
    trace_iter = 0;
#endif
    PNR_MARK(9);
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
            float q[PNR_DOF], tg[3], box[5];
            pnr_reset_draws(p, p.env_id_base + env, pnr_reset_key(p, pnr_tickdom(tick, 0u)), q, tg, box);
            if (q0) {
#pragma unroll
                for (int i = 0; i < PNR_DOF; ++i) q[i] = q0[k * PNR_DOF + i];
            }
            if (target) { tg[0] = target[k * 3 + 0]; tg[1] = target[k * 3 + 1]; tg[2] = target[k * 3 + 2]; }
            pnr_reset_env(s, q, tg);
            if (active) { pnr_store_env(state, N, env, s); pnr_store_box(p, env, box); }
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
    if (lane == 0) pnr_bulk_wait_read<0>();
}

// ---------------------------------------------------------------------------------------------
// RLlib's reset_at(i) after a done (the sampler feeds the policy the RESET observation, bullet_env.py:187-190): for every
// env whose done flag is set -- pnr_step has already reset it -- optionally save the terminal row, then overwrite the row with
// the first observation of the new episode, normalised / pushed into the statistics like the step kernel's output when the
// filter is fused.  Rare rows, thread per env, scalar global stores; fixed launch shape, so it can live in a CUDA graph.
// ---------------------------------------------------------------------------------------------
__global__ void pnr_observe_done_kernel(const __grid_constant__ PnrParams p, float4* __restrict__ state,
                                        const uint8_t* __restrict__ done, float* __restrict__ obs,
                                        float* __restrict__ terminal_out, const float* __restrict__ f_applied,
                                        double* __restrict__ f_delta, float f_clip) {
    const int64_t N = p.n_envs;
    for (int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; env < N; env += (int64_t)gridDim.x * blockDim.x) {
        if (!(done[env] & PNR_DONE)) continue;
        float* row = obs + env * PNR_OBS_DIM;
        if (terminal_out) {
            float* dst = terminal_out + env * PNR_OBS_DIM;
            for (int c = 0; c < PNR_OBS_DIM; ++c) dst[c] = row[c];
        }
        PnrEnv s;
        pnr_load_env(state, N, env, s);
        PnrPose o;
        pnr_pose<false>(p, s, o);
        bool within;
        if (fabsf(o.dist - p.done_distance) < p.done_band) pnr_fk_tip_f64(p, s.r, s.tgt, o.ptr, o.dist, within);
        pnr_pack_obs_const(p, row);
        pnr_pack_obs_dyn<false>(p, row, s, o, s.pot);
        if (f_applied) {
            double* slot = f_delta ? f_delta + (size_t)(env & (PNR_FILTER_SLOTS - 1)) * PNR_FILTER_DELTA_LEN : nullptr;
            for (int c = 0; c < PNR_OBS_DIM; ++c) {
                const float d = row[c] - f_applied[c];
                if (slot) {
                    atomicAdd(&slot[1 + c], (double)d);
                    atomicAdd(&slot[1 + PNR_OBS_DIM + c], (double)d * (double)d);
                }
                row[c] = fminf(fmaxf(d * f_applied[PNR_OBS_DIM + c], -f_clip), f_clip);
            }
            if (slot) atomicAdd(&slot[0], 1.0);
        }
    }
}

cudaError_t pnr_launch_observe_done(const PnrParams& p, float4* state, const uint8_t* done, float* obs, float* terminal_out,
                                    const float* f_applied, double* f_delta, float f_clip, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (p.n_envs + 127) / 128;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    pnr_observe_done_kernel<<<(unsigned)grid, 128, 0, stream>>>(p, state, done, obs, terminal_out, f_applied, f_delta, f_clip);
    return cudaGetLastError();
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

// per-env random box <-> row-major float[N,6] (centre xyz, half extents xyz); utility path
template <bool SET>
__global__ void pnr_box_io_kernel(float4* __restrict__ box_a, float* __restrict__ box_z, int64_t N, float* __restrict__ box) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    if (SET) {
        box_a[e] = make_float4(box[e * 6 + 0], box[e * 6 + 1], box[e * 6 + 3], box[e * 6 + 4]);
        box_z[e] = box[e * 6 + 5];                              // the box stands on z = 0: centre z = half height
    } else {
        const float4 a = box_a[e];
        const float z = box_z[e];
        box[e * 6 + 0] = a.x; box[e * 6 + 1] = a.y; box[e * 6 + 2] = z;
        box[e * 6 + 3] = a.z; box[e * 6 + 4] = a.w; box[e * 6 + 5] = z;
    }
}

// PNR_HOST_COMPACT: gather the 101 changing columns (0:18 and 54:137) of every row for the D2H copy; thread = output float
__global__ void pnr_compact_obs_kernel(const float* __restrict__ full, float* __restrict__ compact, int64_t n_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / PNR_OBS_COMPACT_DIM;
        const int c = (int)(i - row * PNR_OBS_COMPACT_DIM);
        const int col = c < PNR_OBS_CONST_BEGIN ? c : c + (PNR_OBS_CONST_END - PNR_OBS_CONST_BEGIN);
        pnr_st_stream(compact + i, full[row * PNR_OBS_DIM + col]);
    }
}

__global__ void pnr_stats_snapshot_kernel(PnrStats* stats, double* out, int clear) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = stats->episodes; out[1] = stats->sum_return; out[2] = stats->sum_length;
        out[3] = stats->sum_return_sq;
        out[4] = (double)pnr_ordered_to_float(stats->max_return_ord);
        out[5] = (double)pnr_ordered_to_float(stats->min_return_ord);
        out[6] = stats->env_steps; out[7] = stats->reached;
        if (clear) {
            stats->episodes = stats->sum_return = stats->sum_length = stats->sum_return_sq = stats->reached = 0.0;
            stats->env_steps = 0.0;
            stats->max_return_ord = pnr_float_to_ordered(-INFINITY);
            stats->min_return_ord = pnr_float_to_ordered(INFINITY);
        }
    }
}

// The reduction half of the path's one collective: `gathered` = the packed statistics vectors of all ranks (what one
// all-gather delivered), out = SUM over {0,1,2,3,6,7}, MAX over 4, MIN over 5; `extra` doubles per rank behind the eight
// (the observation filter's delta) are summed.  One small kernel instead of half a dozen indexing / reduction launches.
__global__ void pnr_stats_merge_kernel(const double* __restrict__ gathered, int world, int len, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= len) return;
    double acc = gathered[c];
    for (int r = 1; r < world; ++r) {
        const double x = gathered[(size_t)r * len + c];
        acc = c == 4 ? fmax(acc, x) : (c == 5 ? fmin(acc, x) : acc + x);
    }
    out[c] = acc;
}

cudaError_t pnr_launch_stats_merge(const double* gathered, int world, int len, double* out, cudaStream_t stream) {
    pnr_stats_merge_kernel<<<(len + 127) / 128, 128, 0, stream>>>(gathered, world, len, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host-callable launchers (declared in pnr_launch.h)
// ---------------------------------------------------------------------------------------------
// developer knob: PNR_NO_PDL=1 launches the step kernels without programmatic dependent launch (A/B measurements)
bool pnr_pdl_enabled() {
    static const bool on = getenv("PNR_NO_PDL") == nullptr;
    return on;
}

static int pnr_resident_grid(const void* fn, size_t smem) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PNR_STEP_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

template <typename K>
static cudaError_t pnr_prepare(K kernel, size_t smem, int* grid_out) {
    cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    *grid_out = pnr_resident_grid((const void*)kernel, smem);
    return cudaSuccess;
}

// grid = min(CTAs needed, CTAs resident on the whole GPU): a multiple of the SM count once the batch is large
static int64_t pnr_grid_for(int64_t n_envs, int envs_per_cta, int resident) {
    const int64_t blocks = (n_envs + envs_per_cta - 1) / envs_per_cta;
    return blocks < resident ? (blocks > 0 ? blocks : 1) : resident;
}

cudaError_t pnr_launch_step(const PnrParams& p, int device, int arith, int obs_mode, float4* state, const float* actions, float* obs,
                            float* reward, uint8_t* done, PnrStats* stats, uint32_t tick, uint32_t domain, const float* f_applied,
                            double* f_delta, float f_clip, PnrMulti multi, cudaStream_t stream) {
    typedef void (*Kern)(const PnrParams, float4*, const float*, float*, float*, uint8_t*, PnrStats*, uint32_t, uint32_t,
                         const float*, double*, float, const PnrMulti);
    // the obstacle variant and the fused normaliser are separate instantiations: the plain kernel carries no trace of
    // them (a call site alone cost 40 % at 1M envs through caller-saved register spills)
    static Kern kernels[2][2][2] = {
        {{pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_TERMINAL, false, false>, pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_TERMINAL, true, false>},
         {pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_AUTORESET, false, false>, pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_AUTORESET, true, false>}},
        {{pnr_step_kernel<PNR_ARITH_LEGACY64, PNR_OBS_TERMINAL, false, false>, pnr_step_kernel<PNR_ARITH_LEGACY64, PNR_OBS_TERMINAL, true, false>},
         {pnr_step_kernel<PNR_ARITH_LEGACY64, PNR_OBS_AUTORESET, false, false>, pnr_step_kernel<PNR_ARITH_LEGACY64, PNR_OBS_AUTORESET, true, false>}}};
    static Kern fused[2] = {pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_TERMINAL, false, true>,
                            pnr_step_kernel<PNR_ARITH_F32, PNR_OBS_TERMINAL, true, true>};
    static int grids[PNR_MAX_DEVICES][2][2][2][2] = {};  // cudaFuncSetAttribute is per device
    const int obst = p.n_obstacles > 0 ? 1 : 0;
    const int filt = f_applied != nullptr ? 1 : 0;
    if (filt && (arith != PNR_ARITH_F32 || obs_mode != PNR_OBS_TERMINAL)) return cudaErrorInvalidValue;
    Kern k = filt ? fused[obst] : kernels[arith][obs_mode][obst];
    const size_t smem = (filt ? PNR_STEP_SMEM_FILTER : PNR_STEP_SMEM) + (obst ? PNR_STEP_SMEM_OBST : 0);
    int& resident = grids[device % PNR_MAX_DEVICES][arith][obs_mode][obst][filt];
    if (resident == 0) {
        cudaError_t e = pnr_prepare(k, smem, &resident);
        if (e != cudaSuccess) return e;
    }
    // Resident CTAs per SM, measured on B200 (profiles/r01_v4_cta_sweep.md): long grid-stride runs stream best with
    // 4 CTAs per SM (87 % of the HBM copy peak at 2M envs against 80 % with 6), short ones need all 6 to fill the GPU.
    static int sms_of[PNR_MAX_DEVICES] = {};
    int& dev_sms = sms_of[device % PNR_MAX_DEVICES];
    if (dev_sms == 0) cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t n_tiles = (p.n_envs + PNR_TILE_ENVS - 1) / PNR_TILE_ENVS;
    int per_sm = n_tiles >= 16384 ? 4 : (n_tiles >= 8192 ? 5 : 6);
    if (const char* e = getenv("PNR_CTAS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;   // developer knob
    int cap = per_sm * dev_sms < resident ? per_sm * dev_sms : resident;
    const int64_t grid = pnr_grid_for(p.n_envs, PNR_TILE_ENVS, cap);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(PNR_STEP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // programmatic dependent launch pays where the launch tail is a visible share of the step (65,536 envs: 13.2 -> 11.8 us);
    // on long grids the early-resident successor only takes SM slots from this grid (1,048,576 envs: 146 -> 151 us)
    attr[0].val.programmaticStreamSerializationAllowed = (pnr_pdl_enabled() && n_tiles <= 8192) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p, state, actions, obs, reward, done, stats, tick, domain, f_applied, f_delta, f_clip, multi);
}

cudaError_t pnr_launch_reset_observe(const PnrParams& p, int device, int mode, float4* state, const int64_t* idx, int64_t n,
                                     const float* q0, const float* target, float* obs_out, uint32_t tick,
                                     cudaStream_t stream) {
    typedef void (*Kern)(const PnrParams, float4*, const int64_t*, int64_t, const float*, const float*, float*, uint32_t);
    static Kern kernels[2] = {pnr_reset_observe_kernel<0>, pnr_reset_observe_kernel<1>};
    static int grids[PNR_MAX_DEVICES][2] = {};
    int& resident = grids[device % PNR_MAX_DEVICES][mode];
    if (resident == 0) {
        cudaError_t e = pnr_prepare(kernels[mode], PNR_RO_SMEM, &resident);
        if (e != cudaSuccess) return e;
    }
    if (n <= 0) return cudaSuccess;
    const int64_t grid = pnr_grid_for(n, PNR_TILE_ENVS * PNR_STEP_WARPS, resident);
    kernels[mode]<<<(unsigned)grid, PNR_STEP_THREADS, PNR_RO_SMEM, stream>>>(p, state, idx, n, q0, target, obs_out, tick);
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

cudaError_t pnr_launch_box_io(bool set, float4* box_a, float* box_z, int64_t N, float* box, cudaStream_t stream) {
    const int threads = 256;
    const unsigned grid = (unsigned)((N + threads - 1) / threads);
    if (set) pnr_box_io_kernel<true><<<grid, threads, 0, stream>>>(box_a, box_z, N, box);
    else pnr_box_io_kernel<false><<<grid, threads, 0, stream>>>(box_a, box_z, N, box);
    return cudaGetLastError();
}

cudaError_t pnr_launch_compact_obs(const float* full, float* compact, int64_t n_rows, cudaStream_t stream) {
    const int64_t n_out = n_rows * PNR_OBS_COMPACT_DIM;
    if (n_out <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (n_out + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;          // 8 CTAs of 256 threads per SM, grid-stride
    pnr_compact_obs_kernel<<<(unsigned)grid, 256, 0, stream>>>(full, compact, n_out);
    return cudaGetLastError();
}

__global__ void pnr_tick_advance_kernel(PnrStats* stats, uint32_t n, int absolute) {
    if (threadIdx.x == 0 && blockIdx.x == 0) stats->tick_offset = absolute ? n : stats->tick_offset + n;
}
__global__ void pnr_env_steps_set_kernel(PnrStats* stats, double env_steps) {
    if (threadIdx.x == 0 && blockIdx.x == 0) stats->env_steps = env_steps;
}
cudaError_t pnr_launch_env_steps_set(PnrStats* stats, double env_steps, cudaStream_t stream) {
    pnr_env_steps_set_kernel<<<1, 32, 0, stream>>>(stats, env_steps);
    return cudaGetLastError();
}

cudaError_t pnr_launch_tick_advance(PnrStats* stats, uint32_t n, int absolute, cudaStream_t stream) {
    pnr_tick_advance_kernel<<<1, 32, 0, stream>>>(stats, n, absolute);
    return cudaGetLastError();
}

cudaError_t pnr_launch_stats_snapshot(PnrStats* stats, double* out, int clear, cudaStream_t stream) {
    pnr_stats_snapshot_kernel<<<1, 32, 0, stream>>>(stats, out, clear);
    return cudaGetLastError();
}
