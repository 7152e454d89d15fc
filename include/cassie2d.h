/* libcassie2d -- B200-native drop-in for CassieRL/cassierl's bin/libcassie2d.so.
 *
 * Part 1 (legacy ABI) is byte-compatible with the ten extern "C" symbols of the reference
 * (src/Cassie2d/Cassie2d.cpp:15-27) and its POD structs (src/Cassie2d/RobotInterface.h:14-50,
 * mirrored by rllab/envs/cassie2d_structs.py:5-51): one env per handle, host structs,
 * synchronous.  The unmodified ctypes wrappers rllab/envs/cassie2d.py:22-50 and
 * cassie_stand2d.py:20-48 bind these symbols as they are.
 *
 * Part 2 (batch ABI) is the same operator set over a batch dimension: N independent envs per
 * handle, state resident in HBM, one fused CUDA launch per call.  The reference defines no
 * batch interface; each entry point cites the legacy call it generalises.
 *
 * All entry points are plain C: pointers, sizes, ints.  No CUDA or torch types appear; a
 * stream is passed as void* (cudaStream_t, NULL = the legacy default stream).
 * There is no CPU fallback: every call fails (batch: returns < 0, legacy: aborts like the
 * reference's mju_error, Cassie2d.cpp:49-52) when no CUDA device is usable.
 */
#ifndef CASSIE2D_H_
#define CASSIE2D_H_

#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ Part 1: legacy ABI */

/* RobotInterface.h:14-50 -- all double, no padding */
typedef struct { double torques[6]; } ControllerTorque;                      /* :14-16 */
typedef struct { double left_force[3]; double right_force[3]; } ControllerForce; /* :18-21  (Fx, Fz, My) */
typedef struct { double body_xdd[2]; double left_xdd[2]; double right_xdd[2]; double pitch_add; } ControllerOsc; /* :23-28 */
typedef struct { double angles[6]; } ControllerPd;                           /* :30-32 */
typedef struct {                                                             /* :34-41 */
  double base_pos[3]; double base_vel[3];
  double left_pos[5]; double left_vel[5];
  double right_pos[5]; double right_vel[5];
} StateGeneral;
typedef struct {                                                             /* :43-50 */
  double body_x[3]; double body_xd[3];
  double left_x[3]; double left_xd[3];
  double right_x[3]; double right_xd[3];
} StateOperationalSpace;

typedef struct Cassie2d Cassie2d;

Cassie2d* Cassie2dInit(void);                                          /* Cassie2d.cpp:17 */
void Reset(Cassie2d* cassie, StateGeneral* state);                     /* :18 */
void StepOsc(Cassie2d* cassie, ControllerOsc* action);                 /* :19 */
void StepTorque(Cassie2d* cassie, ControllerTorque* action);           /* :20 */
void StepJacobian(Cassie2d* cassie, ControllerForce* action);          /* :21 */
void StepPd(Cassie2d* cassie, ControllerPd* action);                   /* :22 */
void GetGeneralState(Cassie2d* cassie, StateGeneral* state);           /* :23 */
void GetOperationalSpaceState(Cassie2d* cassie, StateOperationalSpace* state); /* :24 */
void Display(Cassie2d* cassie, bool display);                          /* :25 (no-op: no viewer) */
void Render(Cassie2d* cassie);                                         /* :26 (no-op; tolerant of a
                                                                          truncated handle, the Python
                                                                          wrapper declares no argtypes) */

/* ------------------------------------------------------------------ Part 2: batch ABI */

typedef struct CassieBatch CassieBatch;

enum { CASSIE_F32 = 32, CASSIE_F64 = 64 };
/* control mode of a step: which legacy Step* it batches */
enum { CASSIE_MODE_TORQUE = 0,   /* StepTorque   Cassie2d.cpp:86-94,   action dim 6 */
       CASSIE_MODE_PD = 1,       /* StepPd       Cassie2d.cpp:96-117,  action dim 6 */
       CASSIE_MODE_JACOBIAN = 2, /* StepJacobian Cassie2d.cpp:119-177, action dim 6 */
       CASSIE_MODE_OSC = 3 };    /* StepOsc      Cassie2d.cpp:179-209, action dim 7 */
/* task (observation / reward / termination of the Python env) */
enum { CASSIE_TASK_STAND = 0,    /* rllab/envs/cassie_stand2d.py:86-137: 17-d obs */
       CASSIE_TASK_IMITATE = 1 };/* rllab/envs/cassie2d.py:97-225:       26-d obs */
/* flags of Cassie2dBatchEnvStep */
enum { CASSIE_AUTO_RESET = 1,        /* done envs are reset to the standing pose (cassie2d.py:78-88) */
       CASSIE_FRESH_OBS_ON_RESET = 2,/* fix SURVEY App. D.2: recompute the lagged op-space state on reset */
       CASSIE_LIVE_QSTATE = 4,       /* fix SURVEY App. D.4: imitation reward reads the live joint angles */
       CASSIE_TERMINAL_OBS = 8 };    /* with AUTO_RESET: return the terminal observation of a done env instead of
                                        the observation env.reset() returns for its new episode (the default,
                                        because the caller's next action must be computed from the new episode) */
/* QP status in stats[.][3]: 0 optimal, 1 iteration cap, 2 factorisation failed, 3 = the env's state went
 * non-finite (or beyond 1e10) and the env was reset: mj_checkPos / mj_checkVel / mj_checkAcc [EXT].  Such an env
 * reports done = 1, reward 0 and the reset observation; its solver warm start and QP partition are cleared. */
enum { CASSIE_STATUS_DIVERGED = 3 };

/* Last error message of the calling thread ("" if none). */
const char* CassieGetLastError(void);

/* Creates n_envs envs on CUDA device `device`, all at the constructor's standing pose
 * (Cassie2d.cpp:56-64).  xml_path NULL = the packaged cassie2d_stiff.xml (env CASSIE2D_XML
 * overrides).  precision = CASSIE_F32 | CASSIE_F64.  Returns NULL on failure. */
CassieBatch* Cassie2dBatchInit(int n_envs, int device, const char* xml_path, int precision);
void Cassie2dBatchDestroy(CassieBatch* h);
int Cassie2dBatchNumEnvs(const CassieBatch* h);
int Cassie2dBatchPrecision(const CassieBatch* h);
int Cassie2dBatchDevice(const CassieBatch* h);
/* bytes of one real (4 or 8): every `real` buffer below has the handle's precision */
int Cassie2dBatchRealSize(const CassieBatch* h);

/* Reset (Cassie2d.cpp:78-82): writes qpos/qvel of the envs whose mask byte is non-zero
 * (mask NULL = all).  state26 = one StateGeneral in memory order (host pointer, 26 doubles;
 * NULL = the Python reset pose, cassie2d.py:79-85).  As in the reference, time, solver warm
 * start and the lagged op-space state are NOT touched.  mask is a DEVICE pointer [n]. */
int Cassie2dBatchReset(CassieBatch* h, const uint8_t* mask_dev, const double* state26_host, void* stream);
/* per-env states: DEVICE pointer, real [n][26] in StateGeneral memory order */
int Cassie2dBatchSetState(CassieBatch* h, const void* state26_dev, void* stream);
/* GetGeneralState (Cassie2d.cpp:213-216): DEVICE pointer, real [n][26] */
int Cassie2dBatchGetGeneralState(CassieBatch* h, void* state26_dev, void* stream);
/* GetOperationalSpaceState (Cassie2d.cpp:218-237): DEVICE pointer, real [n][18] */
int Cassie2dBatchGetOperationalSpaceState(CassieBatch* h, void* state18_dev, void* stream);

/* n_substeps consecutive legacy Step* calls with the same action (the Python envs' inner
 * loop, cassie2d.py:115-122).  action: DEVICE pointer, real [n][action_dim(mode)].
 * contact_mask_dev (optional, DEVICE uint32 [n]): floor-contact bit mask of the LAST substep
 * (bit 2g+e: geom g of the MJCF file, capsule end e). */
int Cassie2dBatchStep(CassieBatch* h, int mode, const void* action_dev, int n_substeps,
                      uint32_t* contact_mask_dev, void* stream);

/* One policy step of the Python env for every env, fused in one launch: n_substeps Step*,
 * observation, reward, termination, optional auto-reset.
 *   obs_dev    real [n][17] (stand) or [n][26] (imitate)     reward_dev real [n]
 *   done_dev   uint8 [n]                                     (all DEVICE pointers) */
int Cassie2dBatchEnvStep(CassieBatch* h, int task, int mode, const void* action_dev, int n_substeps,
                         int flags, void* obs_dev, void* reward_dev, uint8_t* done_dev, void* stream);
/* env.reset() observation of every env (stale op-space state as in the reference unless
 * CASSIE_FRESH_OBS_ON_RESET): resets all envs to the standing pose, zeroes episode clocks. */
int Cassie2dBatchEnvReset(CassieBatch* h, int task, int flags, void* obs_dev, void* stream);
/* reference trajectory table for CASSIE_TASK_IMITATE: host pointer, [n_rows][13] doubles =
 * Cassie2dTraj.qpos (cassie2d_trajectory.py:31-134), plus t_max = time[-1] and the row count
 * used by state(t) (:16-19). */
int Cassie2dBatchSetTrajectory(CassieBatch* h, const double* qpos_rows_host, int n_rows, double t_max);
/* optional companions of the table for the random-phase reset: Cassie2dTraj.qvel ([n_rows][13]) and .time ([n_rows]);
 * either may be NULL (zero velocities / time[i] = i * t_max / n_rows) */
int Cassie2dBatchSetTrajectoryDetail(CassieBatch* h, const double* qvel_rows_host, const double* time_rows_host, int n_rows);
/* Random-phase reset = Cassie3dTraj.sample() (cassie2d_trajectory.py:26-28: i = randrange(len(time)); time[i], qpos[i],
 * qvel[i]) for every env (mask NULL) or the masked envs, on the device: row i = Philox4x32-10(seed; global env id =
 * first_global_env + e, draw) mod n_rows, so the draw does not depend on the launch partition or the GPU count.  The env
 * restarts at that row: qpos / qvel from the table, env clock = time[i], episode length 0, lagged operational-space
 * state refreshed, warm start and QP partition cleared.  index_out_dev (int32 [n]) and obs_dev (the observation reset()
 * returns: reference slots zero) may be NULL. */
int Cassie2dBatchEnvResetSampled(CassieBatch* h, int task, unsigned long long seed, unsigned int first_global_env,
                                 unsigned int draw, const uint8_t* mask_dev, int32_t* index_out_dev, void* obs_dev, void* stream);

/* Rollout collection (BASELINE configs[4]): what rllab's sampler does around the reference env
 * (rllab/envs/trpo_cassie.py:13-55: GaussianMLPPolicy(hidden_sizes=(32,32)).get_action ->
 * normalize(env).step -> Cassie2dEnv.step(a, n=10)), n_policy_steps per launch for every env.
 *   params_dev  real [n_params]: W1(obs x 32) b1 W2(32 x 32) b2 W3(32 x adim) b3 log_std(adim), the order of
 *               rllab's get_param_values(); tanh hidden layers, linear mean, a = mean + exp(log_std) eps,
 *               eps from Philox4x32-10 keyed by (seed, first_global_env + env, policy step of the env)
 *   normalize   != 0: NormalizedEnv affine map [-1,1] -> [lb,ub] + clip (trpo_cassie.py:13), else clip only
 *   outputs     obs [T][n][odim], action [T][n][adim] (raw policy output, as rllab stores it),
 *               mean [T][n][adim], reward [T][n], done uint8 [T][n] (1 = env terminated, 2 = max_path_length);
 *               an env whose path ended is reset to the standing pose and keeps stepping. */
int Cassie2dBatchRollout(CassieBatch* h, int task, int mode, const void* params_dev, int n_params, int n_policy_steps,
                         int n_substeps, int max_path_length, int flags, int normalize, unsigned long long seed,
                         unsigned int first_global_env, void* obs_dev, void* action_dev, void* mean_dev, void* reward_dev,
                         uint8_t* done_dev, void* stream);

/* Discounted returns of the [T][n] reward / done buffers a rollout wrote, per path (the sampler-side
 * pre-processing of rllab's BatchPolopt [EXT], discount 0.99 at trpo_cassie.py:38):
 * R_t = r_t + gamma R_{t+1}, restarted after every done != 0; tail_dev (optional real [n]) bootstraps the
 * unfinished last path of each env. */
int Cassie2dBatchDiscountedReturns(CassieBatch* h, const void* reward_dev, const uint8_t* done_dev, const void* tail_dev,
                                   double gamma, int n_policy_steps, void* returns_dev, void* stream);

/* LinearFeatureBaseline and GAE advantages over the [T][n] rollout buffers (rllab [EXT]; trpo_cassie.py:29-41
 * constructs LinearFeatureBaseline, TRPO(discount=0.99)).  Features of a sample: [clip(o,-10,10), clip(o)^2,
 * al, al^2, al^3, 1], al = in-path step index / 100, D = 2 odim + 4.
 * BaselineMoments: path_index_dev int32 [T][n] (out) = in-path step index (path_start_dev int32 [n], optional,
 *   = index at which each env enters this buffer); moments_dev double [D(D+1)/2 + D] (out) = packed upper
 *   triangle of F'F row by row, then F'returns.  Solve (F'F + reg I) w = F'y on the host (38 x 38).
 * Advantages: delta_t = r_t + gamma V_{t+1} - V_t, A_t = delta_t + gamma lambda A_{t+1}, V = features . coeffs,
 *   restarted at every done != 0; values_dev (optional) receives V. */
int Cassie2dBatchBaselineMoments(CassieBatch* h, int task, const void* obs_dev, const void* returns_dev, const uint8_t* done_dev,
                                 const int32_t* path_start_dev, int n_policy_steps, int32_t* path_index_dev,
                                 double* moments_dev, void* stream);
int Cassie2dBatchAdvantages(CassieBatch* h, int task, const void* obs_dev, const void* reward_dev, const uint8_t* done_dev,
                            const int32_t* path_index_dev, const void* coeffs_dev, double gamma, double gae_lambda,
                            int n_policy_steps, void* advantages_dev, void* values_dev, void* stream);

/* The squatting.py loop (squatting.py:8-16) on device: n_steps iterations of
 * standing_controller_jacobian (mode JACOBIAN, cassie2d.py:297-331) or
 * standing_controller_osc (mode OSC, cassie2d.py:263-295) with height target
 * 0.7 + 0.25 sin(w t + phase_e), w = 0.5*3.1415, t accumulated by 0.0005 per step and kept
 * per env across calls.  phase_dev: DEVICE real [n] or NULL (0). */
int Cassie2dBatchSquat(CassieBatch* h, int mode, int n_steps, const void* phase_dev,
                       uint32_t* contact_mask_dev, void* stream);

/* Host-buffer variants: same semantics, HOST pointers (pinned or pageable); the call copies
 * host->device, launches, copies device->host and synchronises before returning. */
int Cassie2dBatchStepHost(CassieBatch* h, int mode, const void* action_host, int n_substeps,
                          void* state26_host /* out, may be NULL */);
int Cassie2dBatchEnvStepHost(CassieBatch* h, int task, int mode, const void* action_host, int n_substeps,
                             int flags, void* obs_host, void* reward_host, uint8_t* done_host);
int Cassie2dBatchSquatHost(CassieBatch* h, int mode, int n_steps, const void* phase_host,
                           void* state26_host /* out, may be NULL */);

/* solver statistics of the last Step/EnvStep/Squat call, DEVICE int32 [n][4]:
 * constraint rows, PGS sweeps, QP iterations, QP status of the last substep */
int Cassie2dBatchGetStats(CassieBatch* h, int32_t* stats_dev, void* stream);
/* policy steps since the last env reset of every env (the sampler's path position), DEVICE int32 [n] */
int Cassie2dBatchGetEpisodeLengths(CassieBatch* h, int32_t* ep_len_dev, void* stream);
/* solver warm start (mjData.qacc_warmstart, carried across steps and resets): DEVICE real [n][13].
 * Together with the general state this is the complete per-env simulator state. */
int Cassie2dBatchSetWarmStart(CassieBatch* h, const void* qacc_dev, void* stream);
int Cassie2dBatchGetWarmStart(CassieBatch* h, void* qacc_dev, void* stream);
int Cassie2dBatchSync(CassieBatch* h);

/* Measures the non-tensor FP32 FMA throughput of the device (TFLOP/s) with a register-only
 * FFMA kernel: the denominator of the step kernel's roofline (bench.py). */
double CassieMeasureFp32Peak(int device);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
long long CassieKernelLaunchCount(void);

#ifdef __cplusplus
}
#endif
#endif /* CASSIE2D_H_ */
