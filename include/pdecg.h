/*
 * pdecg.h -- C ABI of libpdecg.so: B200-native (sm_100a) batched PDE-control environment dynamics.
 *
 * This is the drop-in boundary for the hot path of lukebhan/PDEControlGym (per-step PDE
 * integration + boundary actuation + observation sensing + reward / termination evaluation of
 * N independent, lock-stepped environment instances).  The reference has no FFI layer of its
 * own -- the path sits behind the gymnasium `Env.reset/step` contract and the `BaseReward.reward`
 * protocol (reference: pde_control_gym/src/environments1d/base_env_1d.py:20-68,
 * environments2d/base_env_2d.py:24-63, rewards/base_reward.py:14) -- so each entry point below
 * cites the reference method(s) whose work it performs for a whole batch of environments.
 *
 * Conventions
 *   - plain C: pointers, sizes, PODs.  No C++ / torch types cross this boundary.
 *   - every `void*` buffer is DEVICE memory owned by the caller (e.g. a torch tensor's
 *     data_ptr()); the library allocates nothing that outlives a call.  The `*_host` entry points
 *     are the exception: they take HOST buffers and do the H2D/D2H copies themselves.
 *   - all work is enqueued on the caller's stream (`pdecg_stream_t` == cudaStream_t) and returns
 *     asynchronously, except the `*_host` variants which synchronise that stream before returning.
 *   - return value: 0 on success, a negative pdecg_status otherwise; `pdecg_last_error()` gives a
 *     thread-local message.  No exceptions, no aborts, no CPU fallback: without a CUDA device every
 *     compute entry point returns PDECG_ERR_CUDA.
 *   - layouts are env-major and dense: state `[n_envs, npts]`, one row per environment.
 */
#ifndef PDECG_H
#define PDECG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDECG_ABI_VERSION 2

typedef void* pdecg_stream_t; /* cudaStream_t */

typedef enum {
    PDECG_OK = 0,
    PDECG_ERR_INVALID = -1,     /* bad argument / inconsistent configuration            */
    PDECG_ERR_UNSUPPORTED = -2, /* valid in the reference but not implemented by a kernel */
    PDECG_ERR_CUDA = -3         /* CUDA runtime error (message has the cudaError string) */
} pdecg_status;

/* Precision modes.
 *   REF : what the reference stores/computes -- float32 state for Transport / ReactionDiffusion with
 *         float32 or float64 intermediates depending on beta's dtype (hyperbolic.py:146-155,214-216),
 *         float64 for Traffic-ARZ / BrainTumor / NavierStokes.  No FMA contraction, IEEE division:
 *         states are bit-identical to numpy.
 *   FP64: float64 state and arithmetic everywhere, same operation order, no FMA.
 *   FP32: float32 state, FMA contraction allowed (fast, tolerance documented in DESIGN.md).      */
typedef enum { PDECG_PREC_REF = 0, PDECG_PREC_FP64 = 1, PDECG_PREC_FP32 = 2 } pdecg_precision;
typedef enum { PDECG_F32 = 0, PDECG_F64 = 1 } pdecg_dtype;

/* ------------------------------------------------------------------------------------------
 * Family 1: boundary-controlled 1-D envs -- TransportPDE1D and ReactionDiffusionPDE1D
 * ------------------------------------------------------------------------------------------ */
typedef enum { PDECG_ENV_TRANSPORT = 0, PDECG_ENV_REACTION_DIFFUSION = 1 } pdecg_bc1d_env;
typedef enum { PDECG_CONTROL_DIRICHLET = 0, PDECG_CONTROL_NEUMANN = 1 } pdecg_control_type;
typedef enum { PDECG_SENSE_FULL = 0, PDECG_SENSE_COLLOCATED = 1, PDECG_SENSE_OPPOSITE = 2 } pdecg_sensing_loc;
typedef enum { PDECG_SENSETYPE_DIRICHLET = 0, PDECG_SENSETYPE_NEUMANN = 1 } pdecg_sensing_type;
typedef enum { PDECG_REWARD_TUNED1D = 0, PDECG_REWARD_NORM = 1 } pdecg_bc1d_reward;
typedef enum { PDECG_NORM_L1 = 1, PDECG_NORM_L2 = 2, PDECG_NORM_LINF = 3 } pdecg_norm_ord;
typedef enum { PDECG_HORIZON_TEMPORAL = 0, PDECG_HORIZON_DIFFERENTIAL = 1, PDECG_HORIZON_T = 2 } pdecg_horizon;

/* Everything the reference keeps as constructor kwargs (hyperbolic.py:25-35, parabolic.py:25-35,
 * base_env_1d.py:20) plus the reward descriptor (tuned_reward_1d.py:17, norm_reward.py:19). */
typedef struct {
    int32_t abi_version;          /* PDECG_ABI_VERSION                                            */
    int32_t env;                  /* pdecg_bc1d_env                                               */
    int32_t precision;            /* pdecg_precision                                              */
    int32_t beta_dtype;           /* pdecg_dtype of `beta`; REF mode only (selects the arithmetic) */
    int32_t action_dtype;         /* pdecg_dtype of `action` (actuation is evaluated in this type) */
    int32_t nx;                   /* int(round(X/dx))            base_env_1d.py:24                 */
    int32_t nt;                   /* int(round(T/dt)+1)          base_env_1d.py:23                 */
    int32_t substeps;             /* int(round(control_sample_rate/dt))  hyperbolic.py:137          */
    double dt, dx;
    int32_t control_type;         /* pdecg_control_type  ("Dirchilet"/"Neumann")                   */
    int32_t sensing_loc;          /* pdecg_sensing_loc                                            */
    int32_t sensing_type;         /* pdecg_sensing_type (only for OPPOSITE)                        */
    int32_t normalize;            /* base_env_1d.py:36-39                                          */
    double max_control_value;
    int32_t limit_pde_state_size; /* hyperbolic.py:182-194                                         */
    int32_t reward_kind;          /* pdecg_bc1d_reward                                            */
    double max_state_value;
    int32_t reward_nt;            /* the reward object's own `nt` ctor argument                    */
    int32_t norm_ord;             /* pdecg_norm_ord      (NormReward)                              */
    int32_t horizon;              /* pdecg_horizon       (NormReward)                              */
    int32_t t_horizon_length;     /* NormReward                                                   */
    double truncate_penalty, terminate_reward, norm_coeff;
    int32_t truncate_nonfinite;   /* non-zero: an env whose state row has blown up (inf / NaN) comes out `truncated` -- NOT what
                                     the reference does (there inf / NaN just propagate, parabolic.py:148-150 quirk); default 0 */
} pdecg_bc1d_cfg;

/* Device buffers of one batch.  npts = nx (Transport) or nx+1 (ReactionDiffusion, ghost point).
 * State element type S: float for PREC_REF / PREC_FP32, double for PREC_FP64. */
typedef struct {
    int64_t n_envs;
    void* u;                 /* [n_envs, npts] S   -- state; with FULL sensing this IS the observation    */
    const void* beta;        /* [n_envs or 1, npts] -- beta(x) / lambda(x); dtype cfg.beta_dtype (REF) or S */
    int64_t beta_env_stride; /* elements between envs' beta rows; 0 = one row shared by all envs          */
    const void* action;      /* [n_envs] cfg.action_dtype                                               */
    void* obs;               /* [n_envs] S -- scalar observation (COLLOCATED / OPPOSITE); unused for FULL */
    double* reward;          /* [n_envs]                                                                */
    uint8_t* terminated;     /* [n_envs]                                                                */
    uint8_t* truncated;      /* [n_envs]                                                                */
    int32_t* time_index;     /* [n_envs] -- substeps done this episode (the reference's self.time_index) */
    double* carry;           /* [n_envs, pdecg_bc1d_carry_len(cfg)] -- O(1) replacement of the u[nt,nx]
                                history the reference rewards index into                    */
    void* obs_copy;          /* optional [n_envs, npts] S: the step / reset kernels also store the new state here,
                                so a caller gets a private full-state observation without a separate device copy */
    int32_t obs_copy_f32;    /* non-zero: obs_copy (and pdecg_bc1d_autoreset::final_obs) hold float32 whatever S is --
                                the dtype of the reference's observation_space (base_env_1d.py:33-35); each element
                                is the state value rounded once to float32 (numpy `row.astype(np.float32)`).  The
                                resident state keeps its precision.                                             */
    uint8_t* nonfinite;      /* optional [n_envs]: 1 where the state row holds an inf / NaN (or squares overflow float64)
                                after this step -- a per-env failure flag for the learner                            */
} pdecg_bc1d_buffers;

/* Number of doubles of reward carry per env for this configuration. */
int32_t pdecg_bc1d_carry_len(const pdecg_bc1d_cfg* cfg);
/* Number of state points per env (nx or nx+1). */
int32_t pdecg_bc1d_npts(const pdecg_bc1d_cfg* cfg);

/* Replaces TransportPDE1D.reset / ReactionDiffusionPDE1D.reset (hyperbolic.py:198-227,
 * parabolic.py:192-221) for every env whose mask byte is non-zero (mask == NULL: all envs).
 * `init` is [n_envs or 1, npts] of dtype `init_dtype` (cast to S on store, like `u[0] = init`). */
int pdecg_bc1d_reset(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, const void* init,
                     int32_t init_dtype, int64_t init_env_stride, const uint8_t* mask,
                     pdecg_stream_t stream);

/* Randomised reset for RL rollouts: u[env] <- c_env * profile, c_env ~ U(lo, hi) -- the shape every example of
 * the reference uses for its initial conditions (np.ones(nx) * np.random.uniform(1, 10),
 * transport1Dppo.py:36-37, reactionDiffusion1Dppo.py:41-42).  c_env comes from Philox4x32-10 keyed by `seed`
 * with counter (global env id = env_id_offset + env, episode[env]); episode[env] is incremented, so results
 * do not depend on how the batch is sharded over GPUs.  `profile` is [1, npts] (dtype profile_dtype). */
int pdecg_bc1d_reset_random(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, const void* profile,
                            int32_t profile_dtype, double lo, double hi, uint64_t seed, int64_t env_id_offset,
                            int32_t* episode, const uint8_t* mask, pdecg_stream_t stream);

/* Replaces TransportPDE1D.step / ReactionDiffusionPDE1D.step incl. terminate(), truncate(), the
 * sensing closure and reward_class.reward(...) (hyperbolic.py:126-194, parabolic.py:126-189,
 * tuned_reward_1d.py:25-40, norm_reward.py:29-73) for all n_envs in one launch. */
int pdecg_bc1d_step(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, pdecg_stream_t stream);

/* Same-step auto-reset fused into the step launch (the vector-env convention of gymnasium's AsyncVectorEnv
 * NextStep/SameStep wrappers around the reference envs): envs that come out of this step terminated or
 * truncated are re-initialised exactly as pdecg_bc1d_reset (random == 0) / pdecg_bc1d_reset_random
 * (random != 0) with mask = terminated | truncated would do right after it -- state row, obs / obs_copy,
 * carry, time_index, episode counter -- while reward / terminated / truncated keep the step's values. */
typedef struct pdecg_bc1d_autoreset {
    const void* init;        /* [n_envs or 1, npts] initial rows (random == 0) or the [1, npts] profile (random != 0) */
    int32_t init_dtype;      /* PDECG_F32 / PDECG_F64 */
    int32_t random;
    int64_t init_env_stride; /* elements between env rows of `init` (0 = shared row); ignored when random != 0   */
    double lo, hi;           /* random != 0: c_env ~ U(lo, hi)                                                   */
    uint64_t seed;
    int64_t env_id_offset;
    int32_t* episode;        /* random != 0: [n_envs] episode counters                                           */
    void* final_obs;         /* optional: the observation of the step that ENDED the episode, written for restarted envs
                                only (rows of the other envs are left alone) before their row is re-initialised --
                                what SB3 calls infos[i]["terminal_observation"] and gymnasium "final_observation"
                                (examples/transportPDE/transport1Dppo.py:88-90 bootstraps from it on truncation).
                                [n_envs, npts] for FULL sensing, [n_envs] otherwise; S, or float32 if obs_copy_f32 */
} pdecg_bc1d_autoreset;
int pdecg_bc1d_step_autoreset(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf,
                              const pdecg_bc1d_autoreset* autoreset, pdecg_stream_t stream);

/* Same step, but `action_host` ([n_envs], cfg.action_dtype) and the outputs are HOST buffers:
 * H2D(actions) -> kernel -> D2H(obs, reward, flags), then synchronised.  All actions go up in one transfer on
 * `stream`; the batch is then cut into chunks (a short first one, so that the copy-out starts early) that rotate
 * over four library-owned streams ordered after it, so the observation copy of one chunk overlaps the kernels of
 * the next ones; reward and flags leave in one transfer each on a fifth stream once the last kernels have run.
 * obs_host is [n_envs, npts] for FULL sensing, [n_envs] S otherwise; any output
 * pointer may be NULL to skip that copy.  For FULL sensing the observation is copied out of buf->obs_copy when the
 * caller provides that device buffer (dtype S, or float32 with obs_copy_f32 -- half the PCIe bytes of a float64
 * state), else straight out of the state.  Device buffers in `buf` are still the caller's (buf->action is the
 * device staging buffer for the actions).  `autoreset` (may be NULL) fuses the same-step auto-reset as in
 * pdecg_bc1d_step_autoreset; the observation copied out for a restarted env is then its new initial row
 * (autoreset->final_obs, a DEVICE buffer, still receives the terminal one). */
int pdecg_bc1d_step_host(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf,
                         const void* action_host, void* obs_host, double* reward_host,
                         uint8_t* terminated_host, uint8_t* truncated_host,
                         const pdecg_bc1d_autoreset* autoreset, pdecg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Family 2: BrainTumor1D (+ TherapyWrapper semantics)           -- all float64
 * ------------------------------------------------------------------------------------------ */
typedef enum { PDECG_BT_GROWTH = 0, PDECG_BT_THERAPY = 1, PDECG_BT_POST = 2 } pdecg_bt_stage;

/* What one launch does per env (brain_tumor_env.py):
 *   RAW_STEP      BrainTumor1D.step(action)                      (:123-218)  -- one day
 *   WRAP_RESET    TherapyWrapper.reset() after the env reset: run the whole Growth stage  (:408-427)
 *   WRAP_STEP     TherapyWrapper.step(action): one therapy day (+ forced weekend), or the whole
 *                 Post-Therapy stage until termination / truncation            (:429-478)
 *   RUN_TO_END    TherapyWrapper.benchmark(): zero-dose days until the episode ends, then
 *                 t_benchmark <- simulationDays                               (:480-505)          */
typedef enum { PDECG_BT_RAW_STEP = 0, PDECG_BT_WRAP_RESET = 1, PDECG_BT_WRAP_STEP = 2, PDECG_BT_RUN_TO_END = 3 } pdecg_bt_mode;

/* Per-env integer bookkeeping, int32 [n_envs, PDECG_BT_NINT] (the reference's attributes). */
enum {
    PDECG_BT_I_T = 0, PDECG_BT_I_STAGE = 1, PDECG_BT_I_GROWTH_DAYS = 2, PDECG_BT_I_THERAPY_DAYS = 3,
    PDECG_BT_I_POST_DAYS = 4, PDECG_BT_I_SIM_DAYS = 5, PDECG_BT_I_FIRST_THERAPY = 6 /* -1 = None */,
    PDECG_BT_I_FIRST_POST = 7 /* -1 = None */, PDECG_BT_I_DEATH_DAY = 8 /* -1 = None */,
    PDECG_BT_I_T_BENCHMARK = 9 /* -1 = None */, PDECG_BT_I_CONSECUTIVE = 10, PDECG_BT_I_TREATMENT_CALLS = 11,
    PDECG_BT_I_VIOLATIONS = 12, PDECG_BT_NINT = 16
};
/* Per-env float64 values, [n_envs, PDECG_BT_NF64]. */
enum { PDECG_BT_F_REMAINING = 0, PDECG_BT_F_D = 1, PDECG_BT_F_RHO = 2, PDECG_BT_F_LAST_DOSE = 3, PDECG_BT_NF64 = 4 };

/* Constructor kwargs of BrainTumor1D (brain_tumor_env.py:33-101). */
typedef struct {
    int32_t abi_version;
    int32_t nx;                  /* int(round(X/dx)+1)     :52 */
    int32_t nt;                  /* int(round(T/dt)+1)     base_env_1d.py:23 */
    int32_t weekends;            /* TherapyWrapper(weekends=...) */
    double dt, dx;
    double k;                    /* carrying capacity */
    double t1_threshold_k;       /* t1_detection_threshold * k   (:114) */
    double t2_threshold_k;       /* t2_detection_threshold * k */
    double dosage_termination_threshold;
    double alpha, alpha_beta_ratio;
    double t1_detection_radius, t1_death_radius;
    double total_dosage;
    int32_t precision;           /* pdecg_precision: REF / FP64 = float64 like the reference; FP32 = opt-in fast mode: u / obs /
                                    final_obs hold float32, the finite-difference update runs in float32 (same operation order);
                                    doses, D, rho, the radiation factor, reward and every counter stay float64 / int32   */
} pdecg_bt_cfg;

typedef struct {
    int64_t n_envs;
    double* u;              /* [n_envs, nx] state = cell density c(x)                                     */
    double* obs;            /* [n_envs, nx] observation returned by the wrapper modes (may be NULL for RAW_STEP,
                               where the observation is the state itself)                              */
    const double* x_scale;  /* [nx] np.linspace(0, X, nx)   (:57)                                        */
    const double* action;   /* [n_envs] proportion of total_dosage (unused by WRAP_RESET / RUN_TO_END)   */
    int32_t* ienv;          /* [n_envs, PDECG_BT_NINT], 16-byte aligned                                 */
    double* fenv;           /* [n_envs, PDECG_BT_NF64], 16-byte aligned                                 */
    double* reward;         /* [n_envs]                                                                 */
    uint8_t* terminated;    /* [n_envs]                                                                 */
    uint8_t* truncated;     /* [n_envs]                                                                 */
    uint8_t* valid;         /* [n_envs] 0 where the reference's step() would have returned None (:135)  */
    double* final_obs;      /* optional, pdecg_bt_step_autoreset only: [n_envs, nx] terminal observation (the state row of
                               the day that ended the episode) of the envs that restart; other rows are left alone  */
    const uint8_t* mask;    /* optional, PDECG_BT_WRAP_RESET only: [n_envs], run the wrapper reset (Growth stage, consecutive
                               <- 0) only for envs whose byte is non-zero -- TherapyWrapper.reset of a subset of a batch */
    int32_t* stage_out;     /* optional [n_envs]: each env's stage after the launch (after the fused restart, if any) -- a
                               private copy for the caller's info dict, so that no copy kernel follows the step         */
} pdecg_bt_buffers;

/* BrainTumor1D.reset (:347-384): u <- init ([n_envs or 1, nx] float64), stage <- Growth, counters <- 0;
 * D, rho and t_benchmark are kept.  mask == NULL resets every env. */
int pdecg_bt_reset(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, const double* init,
                   int64_t init_env_stride, const uint8_t* mask, pdecg_stream_t stream);
/* One launch of `mode` (pdecg_bt_mode) for all envs. */
int pdecg_bt_run(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, int32_t mode, pdecg_stream_t stream);
/* PDECG_BT_RAW_STEP with the same-step auto-reset fused into the launch (vector-env convention, as
 * pdecg_bc1d_step_autoreset): envs that come out of the day terminated or truncated (or were already past
 * nt-1: valid == 0) are re-initialised exactly as pdecg_bt_reset(init, init_env_stride, mask = terminated |
 * truncated) would do right after it, while reward / terminated / truncated / valid keep the step's values. */
int pdecg_bt_step_autoreset(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, const double* init,
                            int64_t init_env_stride, pdecg_stream_t stream);

/* PDECG_BT_RAW_STEP over HOST buffers (the end-to-end form of BrainTumor1D.step for a caller whose learner lives on
 * the host): H2D(action_host [n_envs]) -> kernel -> D2H(state row = observation [n_envs, nx], reward, flags), all on
 * `stream`, then stream-synchronised.  Any output pointer may be NULL to skip that copy; buf->action is the device
 * staging buffer for the actions.  autoreset_init != NULL fuses the same-step auto-reset as pdecg_bt_step_autoreset. */
int pdecg_bt_step_host(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, const double* action_host,
                       double* obs_host, double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                       uint8_t* valid_host, const double* autoreset_init, int64_t autoreset_init_env_stride,
                       pdecg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Family 3: TrafficPDE1D (Aw-Rascle-Zhang) + TrafficARZReward      -- all float64
 * ------------------------------------------------------------------------------------------ */
typedef enum { PDECG_ARZ_INLET = 0, PDECG_ARZ_OUTLET = 1, PDECG_ARZ_BOTH = 2, PDECG_ARZ_OUTLET_TRAIN = 3 } pdecg_arz_type;
/* Per-env float64 values, [n_envs, PDECG_ARZ_NF64]: the float time (traffic_arz_env.py:148 -- `time_index`
 * advances by dt per STEP and is compared with T/dt, reference quirk Q9) and the steady state. */
enum { PDECG_ARZ_F_TIME = 0, PDECG_ARZ_F_RS = 1, PDECG_ARZ_F_VS = 2, PDECG_ARZ_F_QS = 3,
       /* action_space.low / .high: built ONCE in __init__ from the constructor's qs (:97-100) and never updated by
        * reset(), although 'outlet-train' re-draws rs / qs there (:243-247) -- step() keeps clipping with these */
       PDECG_ARZ_F_ACT_LO = 4, PDECG_ARZ_F_ACT_HI = 5, PDECG_ARZ_NF64 = 6 };

/* Constructor kwargs of TrafficPDE1D (traffic_arz_env.py:23-100); the dt/dx/tau combinations are
 * passed as the Python floats the reference forms (:202-222) so the kernel multiplies by the same bits. */
typedef struct {
    int32_t abi_version;
    int32_t simulation_type;      /* pdecg_arz_type */
    int32_t M;                    /* len(np.arange(0, X+dx, dx))      :74-76 */
    int32_t control_freq;         /* substeps per step                :43    */
    int32_t limit_pde_state_size;
    int32_t action_dim;           /* 1, or 2 for 'both'               :97-100 */
    double T, dt;
    double t_over_dt;             /* self.T / self.dt                 :109   */
    double vm, rm;
    double c1;                    /* dt / (2*dx)                      :202   */
    double c2;                    /* 0.25 * dt / tau                  :208   */
    double c3;                    /* dt / dx                          :218   */
    double c4;                    /* 0.5 * dt / tau                   :221   */
    double vm_over_rm;            /* self.vm / self.rm                :84    */
    int32_t precision;            /* pdecg_precision: REF / FP64 = float64 like the reference (bit-identical states);
                                     FP32 = opt-in fast mode: ry / obs / final_obs hold float32, arithmetic in float32 (same
                                     operation order); actions, fenv and reward stay float64                            */
} pdecg_arz_cfg;

typedef struct {
    int64_t n_envs;
    double* ry;              /* [n_envs, 2, M] state: density rho then y = rho*(v - Veq(rho))  (float32 data in FP32 mode) */
    double* obs;             /* [n_envs, 2*M]  concat(rho, v)  (normalised deviations for OUTLET_TRAIN)  (float32 in FP32 mode) */
    const double* action;    /* [n_envs, action_dim] flow-rate control(s)                               */
    double* fenv;            /* [n_envs, PDECG_ARZ_NF64]                                                */
    double* reward;          /* [n_envs]                                                                */
    uint8_t* terminated;     /* [n_envs]                                                                */
    uint8_t* truncated;      /* [n_envs]                                                                */
    double* final_obs;       /* optional, pdecg_arz_step_autoreset only: [n_envs, 2*M] terminal observation of the envs
                                that restart; other rows are left alone                                          */
} pdecg_arz_buffers;

/* TrafficPDE1D.reset (:237-267): rho <- rs*profile, y, v from the fixed sinusoidal initial condition.
 * `profile` is [M] = sin(3x/L*pi)*0.1 + 1 evaluated by the caller; fenv rows must already hold rs/vs/qs.
 * Like the reference it does NOT reset the time (terminate() does).  mask == NULL: all envs. */
int pdecg_arz_reset(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* profile,
                    const uint8_t* mask, pdecg_stream_t stream);
/* TrafficPDE1D.step (:131-233) + terminate/truncate (:103-127) + TrafficARZReward.reward. */
int pdecg_arz_step(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, pdecg_stream_t stream);
/* The same step with the same-step auto-reset fused into the launch: envs that come out terminated or truncated
 * are re-initialised exactly as pdecg_arz_reset(profile, mask = terminated | truncated) would do right after it
 * (state rows and the observation rows of buf->obs; the float time is left alone, like reset()), while reward /
 * terminated / truncated keep the step's values.  fenv rows keep their rs / vs / qs ('outlet-train' callers that
 * re-draw rs per episode use pdecg_arz_step + pdecg_arz_reset instead). */
int pdecg_arz_step_autoreset(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* profile,
                             pdecg_stream_t stream);

/* TrafficPDE1D.step over HOST buffers: H2D(action_host [n_envs, action_dim]) -> kernel -> D2H(obs [n_envs, 2*M],
 * reward, flags) on `stream`, then stream-synchronised.  Output pointers may be NULL; buf->action is the device
 * staging buffer.  autoreset_profile (DEVICE pointer, may be NULL) fuses the restart as pdecg_arz_step_autoreset. */
int pdecg_arz_step_host(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* action_host,
                        double* obs_host, double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                        const double* autoreset_profile, pdecg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Family 4: NavierStokes2D (Chorin projection, Jacobi pressure solve) + NSReward
 *           float64 (bit-identical fields), or opt-in float32 fields / arithmetic (precision = PDECG_PREC_FP32)
 * ------------------------------------------------------------------------------------------ */
typedef enum { PDECG_NS_DIRICHLET = 0, PDECG_NS_NEUMANN = 1, PDECG_NS_CONTROLLABLE = 2 } pdecg_ns_bc;
/* side order = the order apply_boundary visits them (navier_stokes2D.py:76): lower (row 0),
 * upper (row -1), left (col 0), right (col -1); per side [condition of u, condition of v]. */
enum { PDECG_NS_LOWER = 0, PDECG_NS_UPPER = 1, PDECG_NS_LEFT = 2, PDECG_NS_RIGHT = 3 };

typedef struct {
    int32_t abi_version;
    int32_t nx, ny;               /* int(round(X/dx + 1)), int(round(Y/dy + 1))   base_env_2d.py:28-29 */
    int32_t nt;                   /* int(round(T/dt))                             base_env_2d.py:27    */
    int32_t n_pressure_iters;     /* maximum_pressure_iteration                   navier_stokes2D.py:52 */
    int32_t bc[4][2];             /* pdecg_ns_bc per side per component           :61-91                */
    int32_t n_action_ref;         /* length of action_ref                                               */
    int32_t n_uref;               /* number of time rows of U_ref                                       */
    double dt, dx, dy;
    double viscosity, density;
    double gamma;                 /* NSReward(gamma)                              ns_reward.py:15       */
    int32_t precision;            /* PDECG_PREC_REF / PDECG_PREC_FP64: float64 everywhere, as the reference (bit-identical
                                     fields).  PDECG_PREC_FP32: opt-in fast mode -- u, v, p, obs, final_obs are FLOAT32
                                     buffers and the step runs in float32 (same operation order); actions, u_ref, action_ref,
                                     the reward and the initial fields handed to reset stay float64.  On-chip grids only
                                     (nx * ny <= 4,096).                                                              */
    int32_t reserved;
} pdecg_ns_cfg;

typedef struct {
    int64_t n_envs;
    /* u, v, p, obs, final_obs: float64, or float32 storage behind the same pointers when cfg.precision == PDECG_PREC_FP32 */
    double* u;               /* [n_envs, ny, nx]  x-velocity (row = y index, col = x index, as the reference) */
    double* v;               /* [n_envs, ny, nx]                                                            */
    double* p;               /* [n_envs, ny, nx]  pressure, warm start of the next solve (:115,129)          */
    double* obs;             /* [n_envs, ny, nx, 2] = U[t]                                                  */
    const double* action;    /* [n_envs]                                                                    */
    const double* u_ref;     /* [n_uref, ny, nx, 2] shared reference trajectory                             */
    const double* action_ref;/* [n_action_ref]                                                              */
    int32_t* time_index;     /* [n_envs]                                                                    */
    double* reward;          /* [n_envs]                                                                    */
    uint8_t* terminated;     /* [n_envs]                                                                    */
    uint8_t* truncated;      /* [n_envs] always 0 (:155)                                                    */
    uint8_t* valid;          /* [n_envs] 0 where the reference would raise IndexError (step past nt-1)      */
    double* final_obs;       /* optional, pdecg_ns_step_autoreset only: [n_envs, ny, nx, 2] terminal observation U[nt-1]
                                of the envs that restart; other rows are left alone                                 */
    void* work;              /* scratch of pdecg_ns_workspace_bytes(cfg, n_envs) bytes, needed only by grids beyond 4,096
                                points per env (the HBM-resident path: u*, v*, (dx*dy)*rhs and the pressure ping-pong) */
} pdecg_ns_buffers;

/* Bytes of pdecg_ns_buffers::work this configuration needs for n_envs envs (0: the whole step lives on chip). */
size_t pdecg_ns_workspace_bytes(const pdecg_ns_cfg* cfg, int64_t n_envs);

/* NavierStokes2D.reset (:171-194): u, v, p <- init_* ([n_envs or 1, ny, nx]); time_index <- 0; obs <- U[0]. */
int pdecg_ns_reset(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* init_u,
                   const double* init_v, const double* init_p, int64_t init_env_stride,
                   const uint8_t* mask, pdecg_stream_t stream);
/* NavierStokes2D.step (:118-157) incl. apply_boundary, solve_pressure, terminate and NSReward.reward. */
int pdecg_ns_step(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, pdecg_stream_t stream);
/* The same step with the same-step auto-reset fused into the launch: envs that come out terminated (or were already
 * past the last step: valid == 0) are re-initialised exactly as pdecg_ns_reset(init_*, mask = terminated) would do
 * right after it (u, v, p, obs <- U[0], time_index <- 0), while reward / terminated / truncated / valid keep the
 * step's values. */
int pdecg_ns_step_autoreset(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* init_u,
                            const double* init_v, const double* init_p, int64_t init_env_stride,
                            pdecg_stream_t stream);
/* NavierStokes2D.step over HOST buffers: H2D(action_host [n_envs]) -> kernel -> D2H(obs = U[t] [n_envs, ny, nx, 2],
 * reward, flags) on `stream`, then stream-synchronised.  Output pointers may be NULL; buf->action is the device staging
 * buffer.  autoreset_u/v/p (DEVICE pointers, all three or none) fuse the restart as pdecg_ns_step_autoreset. */
int pdecg_ns_step_host(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* action_host,
                       double* obs_host, double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                       uint8_t* valid_host, const double* autoreset_u, const double* autoreset_v,
                       const double* autoreset_p, int64_t autoreset_env_stride, pdecg_stream_t stream);
/* NavierStokes2D.solve_pressure (:94-116) alone, for callers such as examples/NavierStokes/NS2Doptimization.py:
 * p <- n_pressure_iters Jacobi sweeps on rhs = density/dt * (d u/dx + d v/dy), warm-started from p.  `work`: scratch of
 * pdecg_ns_workspace_bytes(cfg, n_envs) bytes (may be NULL when that is 0). */
int pdecg_ns_solve_pressure(const pdecg_ns_cfg* cfg, int64_t n_envs, const double* u, const double* v,
                            double* p, void* work, pdecg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Family 5: NeuronPDE1D (neuron growth control; closed loop: step() takes no action)   -- float64
 * ------------------------------------------------------------------------------------------ */
/* Constants the reference derives in its constructor (neuron_env.py:60-259), passed as the Python floats it forms. */
typedef struct {
    int32_t abi_version;
    int32_t M;                   /* len(np.arange(0, length + dx, dx))                 :204 */
    int32_t nt;                  /* int(round(T/dt)+1)                                 base_env_1d.py:23 */
    int32_t reserved;
    double dt, dx;
    double D, a, g, RSubG;
    double aTilde1;              /* ((a - RSubG*cInfty)/LSubC) - g - TildeRSubG        :79 */
    double beta;                 /* D / LSubC                                          :83 */
    double LSubS;
    double inv_scale;            /* float(1 / SpatialToRealScale)                      :365 */
    double H0, H1;               /* H = [1, -((a - g*LSubC)*cInfty)/D]                 :76-78 */
    double coef;                 /* (H.T @ B)/D + gamma                                :298 */
} pdecg_neuron_cfg;

typedef struct {
    int64_t n_envs;
    double* u;               /* [n_envs, M] error system u(x, t); points beyond L keep their last values          */
    double* Z;               /* [n_envs, 2] ODE state (cone concentration error, axon length error)               */
    int32_t* L;              /* [n_envs] grid index of the moving boundary l(t)                                   */
    double* time;            /* [n_envs] the float `time_index` (advances by dt per step, zeroed by terminate()) */
    double* control;         /* [n_envs] the backstepping input U(t) of the last step (ControlInput, :305)        */
    const double* pB;        /* [M]    p(x_i) @ B, evaluated by the caller exactly as the reference does (:294)    */
    const double* p;         /* [M, 2] gain kernel p(x) (:241-243; scipy expm on the host)                        */
    uint8_t* terminated;     /* [n_envs] terminate() fired during this call                                       */
} pdecg_neuron_buffers;

/* NeuronPDE1D.step (:263-339) applied n_steps times to every env in one launch (state stays in registers). */
int pdecg_neuron_step(const pdecg_neuron_cfg* cfg, const pdecg_neuron_buffers* buf, int32_t n_steps, pdecg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Misc
 * ------------------------------------------------------------------------------------------ */
const char* pdecg_last_error(void);
int32_t pdecg_abi_version(void);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
int64_t pdecg_launch_count(void);

/* Peak-probe kernels for the roofline denominators that MEASURED_PEAKS.json does not carry
 * (non-FMA FP64 / FP32 issue rate).  `iters` dependent-chain iterations per thread over `n_chains`
 * independent chains; returns elapsed milliseconds in *ms (measured with CUDA events on `stream`) and
 * the number of floating-point instructions executed per thread in *ops_per_thread. */
int pdecg_probe_fp_issue(int32_t dtype, int32_t blocks, int32_t threads, int32_t iters, float* ms,
                         double* total_ops, pdecg_stream_t stream);

/* Self-test of the kernels' "division by a loop-invariant divisor" helper: out_fast[i] = helper(x[i], d),
 * out_ref[i] = IEEE x[i] / d; the two must be bit-identical for every input (tests/test_gpu_constdiv.py). */
int pdecg_selftest_constdiv(int32_t dtype, int64_t n, const void* x, double d, void* out_fast, void* out_ref,
                            pdecg_stream_t stream);

/* Self-test of the kernels' deferred-check division (Traffic-ARZ / BrainTumor inner loops): out_fast[i] = the
 * straight-line fast path of a[i] / b[i], out_ok[i] = 1 when both operands lie inside the exponent windows in which the
 * kernels trust it, out_ref[i] = IEEE a[i] / b[i]; wherever out_ok is 1 the two quotients must be bit-identical
 * (tests/test_gpu_constdiv.py). */
int pdecg_selftest_fastdiv(int32_t dtype, int64_t n, const void* a, const void* b, void* out_fast, uint8_t* out_ok,
                           void* out_ref, pdecg_stream_t stream);

/* Self-test of the correctly rounded exp() the BrainTumor kernels use for the radiation term exp(-alpha * BED)
 * (brain_tumor_env.py:259-265): out[i] = exp_cr(x[i]) (tests/test_gpu_exp.py: bit-identical to mpmath's rounding). */
int pdecg_selftest_exp(int64_t n, const double* x, double* out, pdecg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PDECG_H */
