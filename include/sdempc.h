/*
 * sdempc.h — C ABI of the B200-native neural-SDE MPC solve.
 *
 * This is the drop-in boundary for ONE path of wuwushrek/sde4mbrl_px4: the
 * gradient-based MPC solve the `mpc_controller` node calls every control tick.
 * Every entry point cites the reference interface it replaces (file:line are
 * into the reference repository, `sde4mbrl_px4/mpc_controller/sde_control.py`
 * unless stated otherwise).  The reference obtains four Python callables from
 * the un-vendored package `sde4mbrlExamples` (sde_control.py:12-13, 685):
 *
 *      cfg_dict, (m_reset, m_mpc), state_from_traj, _ = load_mpc_from_cfgfile(path, convert_to_enu=True)
 *
 * and this library is what those callables bind to (see INTEGRATION.md for the
 * ctypes stub a maintainer would add on the reference side).
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature; all pointers are HOST
 *     pointers owned by the caller unless the name says `_dev`;
 *   - every function returns 0 on success and a negative SDEMPC_E* code on
 *     failure; the message is available from sdempc_last_error() (thread local);
 *   - state layout x[13] = [p(3), v(3), q=(qw,qx,qy,qz), w(3)], float32
 *     (sde_control.py:246);
 *   - the CUDA context is created lazily by the first call that needs the GPU
 *     (the node forks the solver process AFTER building the solver objects,
 *     sde_control.py:66-75, 723-728), never by sdempc_create();
 *   - there is no CPU fallback: if no CUDA device is usable the compute calls
 *     fail with SDEMPC_ECUDA.
 */
#ifndef SDEMPC_H_
#define SDEMPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEMPC_NX 13          /* state dimension (sde_control.py:246)            */
#define SDEMPC_MAX_NU 8       /* iris 4, hexa 6 (launch/hexa_sitl_traj_mpc.yaml:7) */
#define SDEMPC_MAX_H 32       /* array bound; horizon <= SDEMPC_MAX_H - 1 = 31 (rows 0..H of the window fit one warp); every shipped config uses 20 */
#define SDEMPC_NNOISE 6       /* noisy state rows: v(3), w(3)                     */
#define SDEMPC_TRACE_W 8      /* floats per iteration in the decision trace       */

#define SDEMPC_MODEL_MAGIC 0x4D454453u /* 'SDEM' */
#define SDEMPC_MODEL_VERSION 1u

/* error codes */
#define SDEMPC_OK 0
#define SDEMPC_EINVAL (-1)   /* bad argument / unsupported configuration */
#define SDEMPC_ECUDA (-2)    /* CUDA runtime failure or no device         */
#define SDEMPC_ENOMEM (-3)
#define SDEMPC_ESTATE (-4)   /* call order (e.g. trajectory mode without table) */

/* flags for sdempc_config.flags */
#define SDEMPC_F_FRAME_ENU 1u      /* external frame is ENU/FLU (convert_to_enu=True, sde_control.py:685) */
#define SDEMPC_F_NO_SHIFT 2u       /* do not shift the plan by one step at the start of a solve */
#define SDEMPC_F_SPECULATIVE_LS 4u /* force latency mode: all line-search trials evaluated concurrently on sibling warps */
#define SDEMPC_F_SEQUENTIAL_LS 8u  /* force the one-warp-per-problem kernel (sequential line search) */
#define SDEMPC_F_GROUP 16u         /* force the throughput kernel (several problems per warp) */
#define SDEMPC_F_NO_CLUSTER 32u    /* latency kernel on one SM (8 warps) instead of a thread-block cluster of 2-16 SMs */
#define SDEMPC_F_TENSOR 64u        /* batched sdempc_solve_ex / sdempc_solve (m_mpc) and sdempc_rollout (value_and_grad) with 1, 2, 4 ... 32       \
                                      particles: network layers and their adjoints on the tensor cores (tcgen05, TF32 operands, fp32         \
                                      accumulation, tanh.approx), the whole APG loop on that mapping (mpc_tcsolve.cuh).  NOT bit-identical   \
                                      to the FP32 path: compared with the oracle teacher-forced and at cost level within the bound stated    \
                                      in DESIGN.md section 5.2.  The solve takes the soft input-rate constraint (a build of its own), the        \
                                      rollout entry point refuses it.  Pays off from a few thousand                                           \
                                      rollout rows (problems x particles) per launch; a single tick stays on the FP32 kernels. */
/* Default kernel choice: latency kernels when the batch fits one problem per SM (or per cluster), the throughput
 * kernel for large batches (> ~13 problems per SM; P = 1, width 32), one warp per (problem, particle) otherwise.
 * All of them produce bit-identical results. */

/*
 * Solver configuration == the YAML schema of launch/iris_sitl_traj_mpc.yaml:1-85
 * (R2 in SURVEY.md section 8a), flattened.
 */
typedef struct sdempc_config {
    int32_t nu;                      /* len(input_constr.input_id), yaml:10                */
    int32_t horizon;                 /* yaml:44; 1 .. SDEMPC_MAX_H - 1                     */
    int32_t num_particles;           /* yaml:52                                            */
    int32_t max_iter;                /* apg_mpc.max_iter, yaml:59                          */
    int32_t max_no_improvement_iter; /* yaml:60                                            */
    int32_t maxls;                   /* apg_mpc.linesearch.maxls, yaml:85                  */
    int32_t reset_option;            /* 0 = conservative, 1 = increase, yaml:84            */
    uint32_t flags;                  /* SDEMPC_F_*                                         */
    float dt[SDEMPC_MAX_H];          /* cfg_dict['_time_steps'] (sde_control.py:167)       */
    float discount;                  /* yaml:49                                            */
    float u_lo[SDEMPC_MAX_NU];       /* input_constr.input_bound, yaml:11                  */
    float u_hi[SDEMPC_MAX_NU];
    float uref[SDEMPC_MAX_NU];       /* cost_params.uref, yaml:33                          */
    float uerr;                      /* yaml:34 */
    float perr[3], verr[3], qerr[3], werr[3]; /* yaml:35-38 */
    float res_mult;                  /* yaml:40 */
    float u_slew_coeff;              /* yaml:41 */
    float init_stepsize, max_stepsize, coef, decrease_factor, increase_factor; /* yaml:76-80 */
    float atol, rtol;                /* yaml:72-73 */
    float beta_init;                 /* yaml:69: initial momentum of the adaptive rule (moment_scale); the classical rule has beta_1 = 1/4 */
    /* Soft input-rate constraint of the position-control configuration
     * (cost_params.u_slew_constr / u_slew_constr_coeff, iris_sitl_posctrl_mpc.yaml:40-41): with
     * ds = u_t[i] - u_{t-1}[i] and e = ds - hi (ds > hi), ds - lo (ds < lo), 0 otherwise, every stage adds
     * u_slew_constr_coeff * e^2 per input.  Coefficient 0 disables the term. */
    float u_slew_constr_coeff;
    float u_slew_lo[SDEMPC_MAX_NU];
    float u_slew_hi[SDEMPC_MAX_NU];
    /* apg_mpc.moment_scale (yaml:63-66: "the adaptive coefficient to scale the momentum ... between 0 and 1"; null = the
     * classical beta_k = k / (k + 3)).  0 = null.  [SPEC] for mu in (0, 1]: the momentum starts at beta_init (yaml:69) after a
     * reset or a rejected step and is scaled by 1 / mu at every accepted step, capped at 1:
     * beta_k = min(1, beta_init / mu^(k-1)), k = accepted steps since the last restart + 1 (mu = 1: constant momentum). */
    float moment_scale;
} sdempc_config;

/*
 * Learned-SDE model blob (`learned_model_params`, yaml:3).  The reference loads a
 * Haiku pickle through the un-vendored sde4mbrl package; this library takes the
 * flat little-endian blob below (the Python layer builds it from an .npz).
 * After the header come, for the drift net and then the diffusion net:
 *   W1[width][n_in], b1[width], W2[width][width], b2[width], W3[6][width], b3[6]
 * all float32, row-major `W[out][in]`.
 */
typedef struct sdempc_model_header {
    uint32_t magic, version;
    int32_t nu, n_in, width, n_hidden, n_out; /* n_in = 6 + nu, n_hidden = 2, n_out = 6 */
    float mass, gravity, k_thrust;            /* T_i = k_thrust * u_i^2                  */
    float inertia[3];                         /* diagonal of J                           */
    float mixer[3 * SDEMPC_MAX_NU];           /* M_b = mixer[3][MAX_NU] . T              */
    float sigma_prior[SDEMPC_NNOISE];         /* prior diffusion on (v, w)               */
} sdempc_model_header;

/* Optimiser telemetry == the fields the node reads from opt_state
 * (sde_control.py:444-450) == msg/OptMPCState.msg:5-24. */
typedef struct sdempc_info {
    float avg_linesearch;
    float stepsize;      /* in/out: carried from tick to tick */
    float num_steps;
    float grad_sqr;
    float avg_stepsize;
    float init_cost;
    float opt_cost;
    float solve_time_us; /* device time of the launch that solved this problem */
} sdempc_info;

typedef struct sdempc_handle sdempc_t;

/* Arguments of one batched solve: B independent MPC problems.
 * Reference selection, first non-NULL wins:
 *   xref_win [B][H+1][13]  explicit reference window per problem
 *   curr_t   [B]           trajectory mode: xref[t] = traj(curr_t + sum_{s<t} dt_s)
 *                          (m_mpc(..., curr_t=traj_time, ...), sde_control.py:412)
 *   xdes     [B][13]       position mode: xref[t] = xdes (sde_control.py:416)
 */
typedef struct sdempc_solve_args {
    int32_t B;
    const float* x;          /* [B][13] current state                            */
    const float* curr_t;     /* [B] or NULL                                      */
    const float* xdes;       /* [B][13] or NULL                                  */
    const float* xref_win;   /* [B][H+1][13] or NULL                             */
    const uint64_t* rng;     /* [B][2] = (seed, tick counter)                    */
    float* u_plan;           /* in/out [B][H][nu]                                */
    float* x_evol;           /* out [B][H+1][13] predicted mean trajectory       */
    sdempc_info* info;       /* in/out [B]                                       */
    const float* xi_override;/* NULL or [B][P][H][6] standard normals (test hook) */
    float* trace;            /* NULL or out [B][max_iter][SDEMPC_TRACE_W]        */
} sdempc_solve_args;

/* load_mpc_from_cfgfile(path, convert_to_enu=True)   (sde_control.py:685)
 * Parses nothing itself: the host layer reads the YAML and the model file and
 * passes both here.  Host-only: no CUDA call is made (fork rule).            */
int sdempc_create(const sdempc_config* cfg, const void* model_blob, size_t nbytes,
                  int device, sdempc_t** out);

/* trajectory_path of the YAML (yaml:6) -> state_from_traj (sde_control.py:694).
 * table[T][14] = (t, 13-state in the external frame), t strictly increasing.   */
int sdempc_set_trajectory(sdempc_t* h, const float* table, int T);

/* state_from_traj(t)   (sde_control.py:206, 694): host-side interpolation,
 * external frame, out[n][13]. */
int sdempc_state_from_traj(const sdempc_t* h, const float* t, int n, float* out);

/* m_reset(x=, rng=, xdes=) -> opt_state   (sde_control.py:345-346, 389-394, 702)
 * u_plan[B][H][nu] <- clip(uref); info <- zeros with stepsize = init_stepsize.
 * Host-only. */
int sdempc_reset(sdempc_t* h, int B, const float* x, const float* xdes,
                 float* u_plan, sdempc_info* info);

/* m_mpc(x, rng, opt_state, curr_t=, xdes=) -> (uopt, opt_state, rng, x_evol)
 * (sde_control.py:400-416, 713-719).  Host buffers in, host buffers out,
 * synchronous: returns when the results are in the caller's memory, i.e. the
 * reference's `.block_until_ready()` point (sde_control.py:420).
 * Zero-copy staging: arrays the caller has page-locked (sdempc_host_register,
 * cudaHostAlloc) are copied straight between its memory and the device; pageable
 * arrays go through the library's pinned staging blocks (one more host copy each
 * way).  Detected per call (cudaPointerGetAttributes); same results either way.  */
int sdempc_solve_ex(sdempc_t* h, const sdempc_solve_args* args);

/* Positional form of sdempc_solve_ex. */
int sdempc_solve(sdempc_t* h, int B, const float* x, const float* curr_t, const float* xdes,
                 const uint64_t* rng, float* u_plan, float* x_evol, sdempc_info* info,
                 const float* xi_override);

/* value_and_grad of the MPC objective at a given control sequence: the
 * building block of m_mpc (particle rollout + cost + adjoint; R6-R8 of
 * SURVEY.md section 8a).  u[B][H][nu], u_prev[B][nu] (slew reference),
 * cost[B], grad[B][H][nu] or NULL, x_evol[B][H+1][13] or NULL.
 * Reference selection as in sdempc_solve_args. */
int sdempc_rollout(sdempc_t* h, int B, const float* x, const float* curr_t, const float* xdes,
                   const float* xref_win, const uint64_t* rng, const float* xi_override,
                   const float* u, const float* u_prev,
                   float* cost, float* grad, float* x_evol);

/* Monte-Carlo closed loop (BASELINE config 5): R rollouts x `ticks` control
 * ticks of the learned-SDE plant under the trajectory MPC, entirely on device.
 * Each tick k: solve at (x_k, t0 + k*dt0) -> apply u*[0] -> one plant EM step
 * (independent Philox stream) -> warm-start shift.  Mirrors the `traj` branch of
 * mpc_process_fn (sde_control.py:410-412) driven at the plan rate.
 * x_hist[R][ticks+1][13] or NULL, u_hist[R][ticks][nu] or NULL,
 * stats[R][4] = (rms position error, max position error, mean opt_cost, mean num_steps). */
int sdempc_closed_loop(sdempc_t* h, int R, int ticks, const float* x0, const float* t0,
                       const uint64_t* rng, float* x_hist, float* u_hist, float* stats);

/* ---- staged API: same solve with the copies split out, for measurement ---- */
/* H2D of the inputs of `args` into the handle's device buffers (async on the handle's stream). */
int sdempc_stage(sdempc_t* h, const sdempc_solve_args* args);
/* Launch the solve kernel `n` times on the staged inputs (each launch restarts
 * from the staged plan).  flush_l2 != 0 overwrites a 256 MiB scratch buffer
 * before each launch.  ms[n] receives the CUDA-event time of each launch,
 * measured on the launching stream. */
int sdempc_launch_timed(sdempc_t* h, int n, int flush_l2, float* ms);
/* Wait for the handle's stream to drain. */
int sdempc_sync(sdempc_t* h);
/* Device address and size of the OUT block of the staged solve: x_evol[B][H+1][13] | u_plan[B][H][nu] |
 * info[B], each sub-array 16-byte aligned, valid until the next stage/solve on this handle.  For callers that
 * forward results device-to-device (the multi-GPU result gather) instead of through host memory. */
int sdempc_device_out(sdempc_t* h, void** dev_ptr, size_t* nbytes);
/* D2H of the outputs of the last launch into `args` and stream synchronise. */
int sdempc_fetch(sdempc_t* h, const sdempc_solve_args* args);

/* sdempc_fetch without the staging copy: device -> host straight into args->x_evol / u_plan / info.  Meant for page-locked
 * destinations (cudaHostRegister'ed, e.g. every rank's slice of a shared-memory result array in the multi-GPU gather);
 * pageable destinations work but are slower than sdempc_fetch. */
int sdempc_fetch_direct(sdempc_t* h, const sdempc_solve_args* args);
/* Page-lock / release a caller-owned host range for sdempc_solve_ex / sdempc_stage / sdempc_fetch_direct (cudaHostRegister /
 * cudaHostUnregister on the current device's context; a failure leaves no sticky CUDA error and the caller keeps using sdempc_fetch). */
int sdempc_host_register(void* p, size_t nbytes);
int sdempc_host_unregister(void* p);

/* CUDA-event duration (ms) of the most recent kernel launch of this handle, on its own stream. */
float sdempc_last_launch_ms(const sdempc_t* h);
/* Kernels launched by this handle so far (for bench.py's gpu_launches). */
int64_t sdempc_launch_count(const sdempc_t* h);
/* Static properties of the compiled solve kernel chosen for this handle:
 * out[0]=threads per CTA, out[1]=dynamic smem bytes, out[2]=problems per CTA,
 * out[3]=registers per thread, out[4]=CTAs launched last time, out[5]=SM count. */
int sdempc_kernel_info(sdempc_t* h, int32_t out[6]);

/* Measurement aid (no reference counterpart): what the FP32 FMA pipe of `device` sustains right now, in TFLOP/s
 * (an FFMA-bound kernel, 16 warps per SM, best of three launches, CUDA-event timed).  bench.py uses it as the
 * denominator of the FP32 roofline, measured in the same run as the solve. */
int sdempc_probe_fp32(int device, float* tflops);

void sdempc_destroy(sdempc_t* h);
const char* sdempc_last_error(void);
const char* sdempc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SDEMPC_H_ */
