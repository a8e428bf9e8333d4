/*
 * pioneer_b200 — C-ABI of the batched Pioneer 6-DoF "reach" environment for NVIDIA B200 (sm_100a).
 *
 * The reference (xdralex/pioneer) has no FFI of its own: its only seams are the gym.Env
 * contract (pioneer/launch/pioneer_knm_train.py:20-29) and the pybullet method set
 * (pioneer/envs/bullet/bullet_scene.py).  This header is the boundary a maintainer binds
 * instead of pybullet; each entry point names the reference interface it replaces.
 * INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C, no torch/CUDA types in signatures: streams travel as void* (a cudaStream_t),
 *     device buffers as raw pointers (e.g. torch.Tensor.data_ptr()).
 *   - every data pointer is CALLER-OWNED, row-major, contiguous; float buffers 16-byte aligned.
 *     The library owns only its internal struct-of-arrays env state.
 *   - every call returns 0 or a negative pnr_status and never throws; pnr_last_error() gives
 *     the text for the calling thread.  Device work is asynchronous on the given stream.
 *   - one handle per GPU; a handle is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device pnr_create fails with PNR_ERR_CUDA.
 */
#ifndef PIONEER_B200_H_
#define PIONEER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNR_ABI_VERSION 4
#define PNR_DOF 6                 /* revolute joints of the Pioneer arm (pioneer_knm_env.py:213-215) */
#define PNR_OBS_DIM 137           /* 21*dof + 11 (pioneer_knm_env.py:194-211)                          */
#define PNR_MAX_CAPSULES 8
#define PNR_MAX_OBSTACLES 4
#define PNR_STATE_WORDS 24        /* 32-bit words of per-env state (6 float4 planes)                   */
#define PNR_STATS_LEN 8

typedef enum pnr_status {
    PNR_OK = 0,
    PNR_ERR_INVALID = -1,         /* bad argument (null pointer, shape, dof != PNR_DOF, ...)           */
    PNR_ERR_CUDA = -2,            /* CUDA runtime error, text in pnr_last_error()                      */
    PNR_ERR_ALLOC = -3,
    PNR_ERR_UNSUPPORTED = -4
} pnr_status;

/* arithmetic of the joint integrator, see DESIGN.md "arithmetic modes" */
#define PNR_ARITH_F32 0           /* float32 throughout = the reference source under NumPy >= 2        */
#define PNR_ARITH_LEGACY64 1      /* NumPy 1.x scalar promotion: float64 intermediates, float32 stores */

/* what `obs` holds for an env whose episode ended in this step */
#define PNR_OBS_TERMINAL 0        /* terminal observation (what BulletEnv.step returns, bullet_env.py:192-197) */
#define PNR_OBS_AUTORESET 1       /* first observation of the next episode (vector-env style)         */

#define PNR_MODE_KINEMATIC 0      /* the reference env: act() integrator + teleport (pioneer_knm_env.py:111-148) */
#define PNR_MODE_DYNAMIC 1        /* ABA forward dynamics + PD torque + semi-implicit Euler substeps   */

/* substep of the dynamic mode */
#define PNR_STEPPING_EXPLICIT 0   /* tau = clamp(kp (u - q) - kd qd) - damping qd; qd += qdd dt; q += qd dt              */
#define PNR_STEPPING_BULLET 1     /* opt-in Bullet-like substep: per-link damping, +-max_velocity clamp, POSITION_CONTROL as
                                   * a velocity-level motor constraint with impulse clamp force * dt (Joint.control_position,
                                   * bullet_scene.py:123-142; btMultiBody semantics restated from memory, unpinned)       */

/* row layout of the HOST observation buffer of pnr_step_host_begin */
#define PNR_HOST_FULL 0           /* float[N,137], the reference's observe() row (pioneer_knm_env.py:184-211)            */
#define PNR_HOST_COMPACT 1        /* float[N,101]: columns 0:18 and 54:137; the 36 columns 18:54 (r_lo, r_hi and their
                                   * cos / sin) never change and are delivered once by pnr_get_obs_constants             */
#define PNR_OBS_COMPACT_DIM 101
#define PNR_OBS_CONST_BEGIN 18
#define PNR_OBS_CONST_END 54

/* bits of the per-env `done` byte */
#define PNR_DONE 1                /* episode over (distance < done_distance, or time limit)            */
#define PNR_TRUNCATED 2           /* gym TimeLimit's info['TimeLimit.truncated']                        */

/* Flattened robot: serial chain with fixed joints folded into their parents.
 * Replaces loadURDF/getJointInfo (bullet_env.py:105-138) — built once by pioneer_b200/urdf.py. */
typedef struct pnr_model {
    int32_t dof;                          /* must equal PNR_DOF                                         */
    int32_t n_capsules;
    double axis[PNR_DOF][3];              /* unit joint axis, joint frame                               */
    double origin_xyz[PNR_DOF][3];        /* previous moving frame -> joint frame                       */
    double origin_rot[PNR_DOF][9];        /* row-major 3x3                                              */
    double tip_xyz[3];                    /* tracked point ('robot:pointer') in the last moving frame   */
    double lower[PNR_DOF], upper[PNR_DOF];/* joint limits as written in the URDF (double)               */
    double effort[PNR_DOF];               /* <limit effort>: torque clamp in dynamic mode               */
    double damping[PNR_DOF];              /* <dynamics damping>                                         */
    double body_mass[PNR_DOF];            /* composite rigid body of each moving frame                  */
    double body_com[PNR_DOF][3];
    double body_inertia[PNR_DOF][9];      /* about the COM, moving-frame axes                           */
    int32_t capsule_body[PNR_MAX_CAPSULES];
    double capsule_radius[PNR_MAX_CAPSULES];
    double capsule_p0[PNR_MAX_CAPSULES][3];
    double capsule_p1[PNR_MAX_CAPSULES][3];
} pnr_model;

#define PNR_OBSTACLE_NONE 0
#define PNR_OBSTACLE_PLANE 1              /* p = point on plane, e = unit normal                        */
#define PNR_OBSTACLE_BOX 2                /* p = centre, e = half extents (axis aligned)                */
#define PNR_OBSTACLE_SPHERE 3             /* p = centre, e[0] = radius                                  */

/* PioneerKinematicConfig (pioneer_knm_env.py:19-34) + SimulationConfig (bullet_env.py:36-44)
 * + gym TimeLimit (pioneer_knm_train.py:27) + the knobs of this implementation. */
typedef struct pnr_config {
    double max_v_to_r, max_a_to_v;
    double done_distance;
    double award_max, award_done, award_potential_slope, penalty_step;
    double target_lo[3], target_hi[3];
    double timestep;                      /* 1/240                                                      */
    int32_t frame_skip;                   /* 10                                                         */
    double gravity;                       /* 0; acts along -z in dynamic mode                           */
    int32_t max_episode_steps;            /* TimeLimit; 0 = no limit                                    */
    int32_t arith;                        /* PNR_ARITH_*                                                */
    int32_t obs_mode;                     /* PNR_OBS_*                                                  */
    int32_t auto_reset;                   /* reset finished envs inside the step kernel                 */
    int32_t mode;                         /* PNR_MODE_*                                                 */
    /* dynamic mode: explicit PD position control  tau = kp*(q_des-q) - kd*qd, clamped to +-effort*torque_scale */
    double kp, kd, torque_scale;
    /* obstacle variant (Joint/Scene.create_body_box/plane, pioneer_knm_env.py:249-261) */
    int32_t n_obstacles;
    int32_t obstacle_type[PNR_MAX_OBSTACLES];
    double obstacle_p[PNR_MAX_OBSTACLES][3];
    double obstacle_e[PNR_MAX_OBSTACLES][3];
    double contact_penalty;               /* reward -= contact_penalty * sum(penetration depth)         */
    /* per-env random box (the legacy randomizer, pioneer/temp/pioneer_env.py:169-192): with random_box != 0 the FIRST box
     * among the obstacles is redrawn at every reset of an env: half extents ~ U(box_size_lo, box_size_hi), centre =
     * (U(box_pos_lo, box_pos_hi), half height) -- the box stands on z = 0 */
    int32_t random_box;
    double box_pos_lo[2], box_pos_hi[2], box_size_lo[3], box_size_hi[3];
    /* dynamic mode, PNR_STEPPING_BULLET only */
    int32_t stepping;                     /* PNR_STEPPING_*                                             */
    double link_damping;                  /* Bullet's linear = angular link damping, 0.04               */
    double max_velocity;                  /* joint velocity clamp, 100 rad/s                            */
    double motor_kp, motor_kd;            /* setJointMotorControl2 positionGain 0.1 / velocityGain 1.0  */
    double motor_max_force;               /* `force`: impulse clamp force * dt per substep; 0 = motors off */
} pnr_config;

typedef struct pnr_handle pnr_handle;

int pnr_abi_version(void);
const char* pnr_last_error(void);

/* Fill `cfg` with the reference defaults (pioneer_knm_env.py:19-34, bullet_env.py:36-44, TimeLimit 500). */
void pnr_default_config(pnr_config* cfg);

/* Replaces PioneerKinematicEnv.__init__ + BulletEnv.reset_simulator (pioneer_knm_env.py:39-74,
 * bullet_env.py:90-99) for n_envs independent envs on CUDA device `device`.  `env_id_base` is the
 * global id of local env 0 (multi-GPU sharding; reset randomness is keyed on the global id so results
 * do not depend on the number of GPUs).  All envs start reset (as reset_world()).
 * cfg->mode = PNR_MODE_DYNAMIC selects the ABA + PD/torque kernel (World.step / Joint.control_position,
 * bullet_scene.py:123-155, 273-275); cfg->n_obstacles > 0 with a non-zero contact_penalty selects the obstacle
 * variant (Scene.create_body_box / create_body_plane, bullet_scene.py:206-259).  Both are defined in DESIGN.md
 * sections 8-9 and are unpinned by the reference. */
int pnr_create(const pnr_model* model, const pnr_config* cfg, int64_t n_envs, int64_t env_id_base,
               int device, uint64_t seed, pnr_handle** out);
void pnr_destroy(pnr_handle* h);

int64_t pnr_num_envs(const pnr_handle* h);
/* derived bounds (pioneer_knm_env.py:56-58): each out pointer is HOST float[PNR_DOF] or NULL */
int pnr_get_bounds(const pnr_handle* h, float* r_lo, float* r_hi, float* v_max, float* a_max);
/* obs[18:54] of every observation row (r_lo, cos r_lo, sin r_lo, r_hi, cos r_hi, sin r_hi; pioneer_knm_env.py:196-197)
 * exactly as the kernels write them: HOST float[36].  With these a PNR_HOST_COMPACT row re-expands to the 137-column row
 * bit for bit (pnr_expand_obs_host). */
int pnr_get_obs_constants(const pnr_handle* h, float* out36);
/* re-seed the reset generator (PioneerKinematicEnv.seed, pioneer_knm_env.py:107-109).  The key lives in device memory, so
 * steps already captured into a CUDA graph see the new seed too.  Synchronises the device. */
int pnr_seed(pnr_handle* h, uint64_t seed);

/* Replaces BulletEnv.reset / reset_world(joint_positions, target_position) (bullet_env.py:187-190,
 * pioneer_knm_env.py:76-105).  idx: DEVICE int64[n] local env indices, or NULL = all envs (then n must
 * be n_envs).  q0: DEVICE float[n,6] or NULL = uniform in [r_lo, r_hi].  target: DEVICE float[n,3] or
 * NULL = uniform in [target_lo, target_hi].  obs_out: DEVICE float[n,137] or NULL. */
int pnr_reset(pnr_handle* h, const int64_t* idx, int64_t n, const float* q0, const float* target,
              float* obs_out, void* stream);

/* Replaces BulletEnv.step -> act() + observe() through gym TimeLimit (bullet_env.py:192-197,
 * pioneer_knm_env.py:111-211) for every env in one fused kernel launch.
 * actions DEVICE float[N,6] (stored unclipped, as pioneer_knm_env.py:144); obs DEVICE float[N,137];
 * reward DEVICE float[N]; done DEVICE uint8[N] (PNR_DONE | PNR_TRUNCATED bits). */
int pnr_step(pnr_handle* h, const float* actions, float* obs, float* reward, uint8_t* done, void* stream);

/* A rollout fragment with pre-computed actions: n_steps consecutive steps in ONE kernel launch.  Step t reads
 * actions + t * action_stride and writes obs + t * obs_stride (strides in floats; obs_stride * 4 must be a multiple of 16
 * bytes, action_stride * 4 of 8), reward + t * N, done + t * N; results are those of n_steps pnr_step calls, bit for bit.
 * Envs never depend on each other, so inside the launch every CTA (kinematic mode) or warp (dynamic mode, env state kept in
 * registers) runs the n_steps steps on its own 32-env tiles: no launch gap and no grid-wide wait between the steps. */
int pnr_step_many(pnr_handle* h, int32_t n_steps, const float* actions, int64_t action_stride, float* obs,
                  int64_t obs_stride, float* reward, uint8_t* done, void* stream);

/* Same step with HOST buffers (pinned or pageable): H2D of actions, the kernel, D2H of obs / reward / done on the
 * library's own streams (the observation copy split over two copy engines); returns after the results are on the host.
 * The work is ordered after whatever was queued on the default stream; callers using other non-blocking streams
 * synchronise them first. */
int pnr_step_host(pnr_handle* h, const float* actions, float* obs, float* reward, uint8_t* done);

/* The same step, asynchronous and double-buffered: _begin queues H2D of the actions, the kernel and the D2H of the results
 * into the caller's HOST buffers (pinned memory for real overlap) and returns at once; _end blocks until the results of the
 * OLDEST step begun have landed.  At most two steps may be in flight (begin, begin, end, begin, end, ...): the D2H of step
 * k then overlaps H2D + kernel of step k + 1 -- for callers whose next action does not depend on the observation still in
 * flight.  layout = PNR_HOST_FULL: obs is float[N,137]; PNR_HOST_COMPACT: obs is float[N,101] (26 % fewer bytes over PCIe;
 * a small gather kernel runs after the step).  pnr_step_host is _begin(FULL) + _end. */
int pnr_step_host_begin(pnr_handle* h, const float* actions, float* obs, float* reward, uint8_t* done, int layout);
int pnr_step_host_end(pnr_handle* h);
/* HOST helper (plain memory formatting, no device work): compact float[n_rows,101] -> full float[n_rows,137]. */
int pnr_expand_obs_host(const pnr_handle* h, const float* compact, float* full, int64_t n_rows);

/* Replaces PioneerKinematicEnv.observe (pioneer_knm_env.py:184-211) on the current state.
 * idx DEVICE int64[n] or NULL = all. */
int pnr_observe(pnr_handle* h, const int64_t* idx, int64_t n, float* obs_out, void* stream);

/* What RLlib's sampler does after a done: reset_at(i) and feed the policy the RESET observation (bullet_env.py:187-190;
 * VectorEnv.reset_at).  With auto_reset the envs flagged in `done` (the DEVICE uint8[N] pnr_step wrote) are already in
 * their new episode; this call overwrites their rows of obs DEVICE float[N,137] (PNR_OBS_TERMINAL content) with the first
 * observation of the new episode -- normalised and pushed into the statistics like the step's output when the filter is
 * fused -- after copying the terminal rows to terminal_obs DEVICE float[N,137] (or NULL; only done rows are written).
 * Fixed launch shape: capturable in a CUDA graph. */
int pnr_observe_done(pnr_handle* h, const uint8_t* done, float* obs, float* terminal_obs, void* stream);

/* Replaces Joint.position()/velocity() and the env attributes a, v, r, potential (bullet_scene.py:115-121,
 * pioneer_knm_env.py:63-66).  All DEVICE, any may be NULL: r,v,a float[N,6]; potential float[N];
 * target float[N,3]; t int32[N] (steps since reset); ep_return float[N]. */
int pnr_get_state(pnr_handle* h, float* r, float* v, float* a, float* potential, float* target,
                  int32_t* t, float* ep_return, void* stream);
int pnr_set_state(pnr_handle* h, const float* r, const float* v, const float* a, const float* potential,
                  const float* target, const int32_t* t, const float* ep_return, void* stream);
/* The per-env random box of the obstacle variant (cfg->random_box): DEVICE float[N,6] = centre xyz, half extents xyz.
 * PNR_ERR_UNSUPPORTED without random_box. */
int pnr_get_boxes(pnr_handle* h, float* box, void* stream);
int pnr_set_boxes(pnr_handle* h, const float* box, void* stream);

/* Counters of the handle, for checkpoint / resume together with pnr_get_state / pnr_set_state: `tick` keys the
 * reset generator (one per reset / step call, plus what pnr_tick_advance added on the device), `env_steps` feeds
 * pnr_stats (counted on the device by the step kernels, so CUDA-graph replays count).  Both calls synchronise the DEVICE (every
 * stream), so the values are current whatever stream the steps ran on.  (The reference env has no state
 * save / restore of its own -- it is re-created from constructor arguments, pioneer_knm_env.py:38,51.)
 * Reset keys never repeat: eager launches draw with (tick = host call counter, domain 0); launches captured into the g-th
 * CUDA graph of this handle draw with (tick = captured counter + device-side advance, domain g), so eager steps and graph
 * replays may be interleaved freely. */
int pnr_get_counters(const pnr_handle* h, uint32_t* tick, double* env_steps, uint64_t* seed);
/* Advance the reset generator's counter by `n` ON THE DEVICE (a one-thread kernel on `stream`).  The host counter that
 * pnr_step passes to its kernel is frozen into a captured CUDA graph; capture this call as the first node of a graph of
 * `n` steps and every replay draws fresh reset states (BatchedPioneerEnv.capture_rollout does).  pnr_get_counters reports
 * host counter + device advance; pnr_set_counters clears the device part. */
int pnr_tick_advance(pnr_handle* h, uint32_t n, void* stream);
int pnr_set_counters(pnr_handle* h, uint32_t tick, double env_steps);

/* Episode statistics accumulated on the device since the last clear (the columns cli.py:32-38 prints):
 * out HOST double[8] = {episodes, sum_return, sum_length, sum_return^2, max_return, min_return,
 * env_steps, reached_target}.  Synchronises `stream`. */
int pnr_stats(pnr_handle* h, double* out, int clear, void* stream);
/* Same 8 numbers written to a caller-owned DEVICE double[8] without synchronising, for an NCCL
 * all-reduce on the same stream (SUM over {0,1,2,3,6,7}, MAX over {4, -5}). */
int pnr_stats_device(pnr_handle* h, double* out_device, int clear, void* stream);
/* The reduction half of the path's one collective (the reference moves the same numbers from its Ray rollout workers to
 * the trainer, pioneer/launch/pioneer_knm_train.py:49): gathered DEVICE double[world, len] = what ONE all-gather of every
 * rank's packed vector delivered (len >= 8: the statistics, optionally followed by additive extras such as the filter
 * delta); out DEVICE double[len] = SUM over slots {0,1,2,3,6,7} and every extra, MAX over slot 4, MIN over slot 5.
 * One kernel on `stream` of the current device; no handle needed. */
int pnr_stats_merge_device(const double* gathered, int world, int len, double* out, void* stream);
/* The whole once-per-iteration exchange as ONE kernel over NVLink peer memory: snapshot of this rank's statistics window
 * (cleared if `clear`) [+ with_filter: this rank's filter delta], stores into every rank's window, a flag per peer, the
 * merge in rank order (every rank ends with bit-identical results) [+ the Chan merge of the summed delta into the running
 * filter statistics].  out_device DEVICE double[8 (+ PNR_FILTER_DELTA_LEN)] = what pnr_stats_merge_device would have
 * produced from an all-gather.  No NCCL call, so the launch can sit in a CUDA graph at any world size.
 *   pnr_sync_window_create   allocates this handle's window (PNR_SYNC_WINDOW_BYTES, zeroed) and returns its CUDA IPC handle
 *                            (PNR_SYNC_IPC_BYTES bytes) for the host to exchange over whatever transport it has
 *                            (torch.distributed, MPI, a socket); pnr_sync_window_ptr returns the raw device pointer for
 *                            ranks that share a process or map the memory themselves.
 *   pnr_sync_window_connect  opens the peers' windows: ipc_handles HOST bytes [world][PNR_SYNC_IPC_BYTES] in rank order
 *                            (entry `rank` is ignored).  pnr_sync_window_connect_ptrs takes mapped device pointers instead.
 *                            All ranks must have connected (a host barrier) before the first pnr_iteration_sync.
 *   pnr_iteration_sync       world = 1 (never connected): the same kernel without the exchange.  timeout_ms > 0 bounds
 *                            the wait for a peer: on expiry the output is NaN and pnr_sync_status reports 1 (0 = wait
 *                            for ever).  Every rank must make the same sequence of calls with the same with_filter.
 * The reference moves these numbers from its Ray rollout workers to the trainer process
 * (pioneer/launch/pioneer_knm_train.py:49, :66). */
#define PNR_SYNC_MAX_PEERS 16
#define PNR_SYNC_WINDOW_BYTES 73984
#define PNR_SYNC_IPC_BYTES 64
int pnr_sync_window_create(pnr_handle* h, unsigned char* ipc_handle_out);
int pnr_sync_window_ptr(pnr_handle* h, void** window_out);
int pnr_sync_window_connect(pnr_handle* h, const unsigned char* ipc_handles, int world, int rank);
int pnr_sync_window_connect_ptrs(pnr_handle* h, void* const* windows, int world, int rank);
int pnr_iteration_sync(pnr_handle* h, int with_filter, int clear, double* out_device, int timeout_ms, void* stream);
int pnr_sync_status(pnr_handle* h, int* timed_out);
/* Restore the statistics window from HOST double[8] (what pnr_stats returned): checkpoint / resume.  Synchronises. */
int pnr_set_stats(pnr_handle* h, const double* in8);

/* ---- observation normaliser: the 'ConcurrentMeanStdFilter' observation_filter of the reference launcher
 * (pioneer/launch/pioneer_knm_train.py:66; the filter is RLlib's MeanStdFilter behind a lock, third party) -----------
 * pnr_filter_apply: ONE pass over obs_in DEVICE float[n_rows,137]: if `update`, the rows enter the statistics
 * accumulated since the last pnr_filter_sync; if `normalize`, obs_out = clip((obs_in - mean) / (std + 1e-8), +-clip)
 * with the mean / std of the last synchronisation (obs_out may alias obs_in; with normalize = 0 and obs_out == obs_in
 * nothing is written).  Defaults: clip = 10, demean and destd on (RLlib's MeanStdFilter).  Before the first synchronisation
 * that has seen a row the filter is the identity (mean 0, scale 1). */
#define PNR_FILTER_DELTA_LEN (1 + 2 * PNR_OBS_DIM)   /* rows, sum(x - mean)[137], sum((x - mean)^2)[137] */
int pnr_filter_configure(pnr_handle* h, double clip, int demean, int destd);
int pnr_filter_apply(pnr_handle* h, const float* obs_in, float* obs_out, int64_t n_rows, int update, int normalize,
                     void* stream);
/* Fuse the normaliser into pnr_step: with `on`, the observations pnr_step writes are already normalised and (with
 * `update`) their raw values have entered the statistics -- no second pass over the 548 B/env observation.  Kinematic
 * mode: PNR_ARITH_F32 and PNR_OBS_TERMINAL only (PNR_ERR_UNSUPPORTED otherwise); dynamic mode: both observation modes. */
int pnr_filter_fuse(pnr_handle* h, int on, int update);
/* Copy the statistics accumulated since the last sync to a caller-owned DEVICE double[PNR_FILTER_DELTA_LEN]
 * (not synchronised): sum it over the ranks with one NCCL all-reduce, then hand it to pnr_filter_sync on every rank. */
int pnr_filter_delta_device(pnr_handle* h, double* out_device, void* stream);
/* Merge a delta (HOST double[PNR_FILTER_DELTA_LEN], or NULL = this handle's own) into the running statistics, refresh the
 * mean / std used by pnr_filter_apply, clear the accumulator.  Synchronises `stream`. */
int pnr_filter_sync(pnr_handle* h, const double* merged_delta_host, void* stream);
/* The same merge with the (all-reduced) delta still on the DEVICE, or NULL = this handle's own accumulator: one small
 * kernel on `stream`, no host round trip -- the running count / mean / M2 live on the device, so a rollout loop can
 * synchronise its filter once per iteration without stalling the stream.  pnr_filter_sync is this call after an H2D copy. */
int pnr_filter_sync_device(pnr_handle* h, const double* merged_device, void* stream);
/* Running statistics: HOST count[1], mean[137], var[137] (var = M2 / (count - 1), or mean^2 while count < 2, as
 * RLlib's RunningStat reports it). */
int pnr_filter_get(pnr_handle* h, double* count, double* mean, double* var);
int pnr_filter_set(pnr_handle* h, double count, const double* mean, const double* var, void* stream);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t pnr_launch_count(const pnr_handle* h);

#ifdef __cplusplus
}
#endif
#endif  /* PIONEER_B200_H_ */
