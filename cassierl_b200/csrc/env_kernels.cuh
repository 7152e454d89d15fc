// CUDA kernels of the batched Cassie2d engine: one env per thread, env state in registers for
// the whole launch (all substeps fused), constraint rows staged in thread-local memory and the
// constraint system (A, b, f) in registers for the PGS sweeps, model constants in the kernel-parameter
// constant bank (warp-uniform reads), 64-thread CTAs kept in lock step by one barrier per sim step (32 for OSC).
// DESIGN.md sections 4-5 have the mapping rationale and the measured optimisation log.
#pragma once
#include "batch_state.h"
#include "cassie_step.cuh"

namespace cassie {

#ifndef CASSIE_BLOCK
#define CASSIE_BLOCK 64
#endif
constexpr int kBlock = CASSIE_BLOCK;
// resident CTAs per SM the register allocation is sized for: 1 = all 255 registers (one warp per scheduler at
// 16384 envs anyway); larger values trade spills for occupancy on very large batches (profiles/r1_variants.txt)
#ifndef CASSIE_MIN_BLOCKS
#define CASSIE_MIN_BLOCKS 1
#endif
// With several warps per CTA, a barrier per simulator step keeps them in lock step so that they share
// instruction-cache fills (instruction fetch is the top stall of the step kernels).
// The OSC kernels run one warp per CTA instead: they are bound by the L1 hit rate of the controller's thread-local
// arrays, and 512 single-warp CTAs spread more evenly over the 148 SMs (3-4 warps each instead of 2-4): +2.5 %.
#ifndef CASSIE_OSC_BLOCK
#define CASSIE_OSC_BLOCK 32
#endif
constexpr int block_threads(int mode) { return mode == kModeOsc ? CASSIE_OSC_BLOCK : kBlock; }
template <int B>
__device__ __forceinline__ void step_barrier() {
  if constexpr (B > 32) __syncthreads();
}

// controller model of a mode: OSC runs in double in every build, the other modes in T
template <int MODE, typename T>
__device__ __forceinline__ const auto& ctrl_model(const ModelPair<T>& mp) {
  if constexpr (MODE == kModeOsc) return mp.ctrl_d;
  else return mp.ctrl;
}

template <typename T>
__device__ __forceinline__ void load_env(const BatchView<T>& v, int e, T q[kNV], T qd[kNV], T w[kNV]) {
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < kNV; i++) {
    q[i] = v.qpos[(size_t)i * n + e];
    qd[i] = v.qvel[(size_t)i * n + e];
    w[i] = v.warm[(size_t)i * n + e];
  }
}
template <typename T>
__device__ __forceinline__ void store_env(const BatchView<T>& v, int e, const T q[kNV], const T qd[kNV], const T w[kNV]) {
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < kNV; i++) {
    v.qpos[(size_t)i * n + e] = q[i];
    v.qvel[(size_t)i * n + e] = qd[i];
    v.warm[(size_t)i * n + e] = w[i];
  }
}
template <typename T>
__device__ __forceinline__ void load_op(const BatchView<T>& v, int e, OpState<T>& op) {
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    op.body[i] = v.op[(size_t)i * n + e];
    op.left[i] = v.op[(size_t)(4 + i) * n + e];
    op.right[i] = v.op[(size_t)(8 + i) * n + e];
  }
}
template <typename T>
__device__ __forceinline__ void store_op(const BatchView<T>& v, int e, const OpState<T>& op) {
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v.op[(size_t)i * n + e] = op.body[i];
    v.op[(size_t)(4 + i) * n + e] = op.left[i];
    v.op[(size_t)(8 + i) * n + e] = op.right[i];
  }
}
__device__ __forceinline__ void store_stats(int32_t* stats, int n, int e, const StepStats& st, const OscStats& qs) {
  (void)n;
  reinterpret_cast<int4*>(stats)[e] = make_int4(st.nrows, st.sweeps, qs.iters, qs.status);
}

// StateGeneral memory order (RobotInterface.h:34-41, converters :98-128) <-> qpos/qvel
template <typename T>
__device__ __forceinline__ void state26_to_q(const T* s, T q[kNV], T qd[kNV]) {
#pragma unroll
  for (int i = 0; i < 3; i++) { q[i] = s[i]; qd[i] = s[3 + i]; }
#pragma unroll
  for (int i = 0; i < 5; i++) { q[3 + i] = s[6 + i]; qd[3 + i] = s[11 + i]; q[8 + i] = s[16 + i]; qd[8 + i] = s[21 + i]; }
}

// ---------------------------------------------------------------------------------------
// n_substeps x Step* (Cassie2d.cpp:86-209) with a held action
template <typename T, int MODE>
__global__ void __launch_bounds__(block_threads(MODE), CASSIE_MIN_BLOCKS) k_step(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v,
                                                  const T* __restrict__ action, int n_sub, uint32_t* mask) {
  const int e = blockIdx.x * block_threads(MODE) + threadIdx.x;
  if (e >= v.n) return;
  T q[kNV], qd[kNV], w[kNV], act[7], u[kNU];
  load_env(v, e, q, qd, w);
  constexpr int adim = action_dim(MODE);
#pragma unroll
  for (int i = 0; i < adim; i++) act[i] = action[(size_t)e * adim + i];
  alignas(16) Rows<T> rows;
  OpState<T> op;
  StepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  unsigned qps = v.qp_set[e];
  for (int s = 0; s < n_sub; s++)
    controller_step<MODE>(mp.phys, mp.phys_d, ctrl_model<MODE>(mp), mp.ctrl_d, q, qd, w, act, rows, u, s == n_sub - 1 ? &op : nullptr, &st, &qs, &qps);
  v.qp_set[e] = qps;
  store_env(v, e, q, qd, w);
  if (n_sub > 0) {
    store_op(v, e, op);
    store_stats(v.stats, v.n, e, st, qs);
    if (mask) mask[e] = st.contact_mask;
  }
}

// ---------------------------------------------------------------------------------------
// One policy step of the Python env (cassie_stand2d.py:86-137 / cassie2d.py:97-225)
template <typename T>
struct EnvStepDev {
  int task, flags, n_sub;
  const T* action;
  T* obs;
  T* reward;
  uint8_t* done;
  T reset_state[26];
};

// Cassie2dTraj.state(t) (cassie2d_trajectory.py:16-19) in IEEE double like Python
__device__ __forceinline__ int traj_index(double t, double tmax, int rows) {
  int i = (int)(fmod(t, tmax) / tmax * (double)rows);
  return i < rows ? i : rows - 1;
}

// zero_ref: the observation env.reset() returns has zeros in its reference slots 17..25 (cassie2d.py:78-95 returns
// operational_state_array_to_pos_invariant_array(s), cassie2d_structs.py:68-75; only step() fills them, :158-166)
template <typename T>
__device__ __forceinline__ void write_obs(const BatchView<T>& v, int task, int e, const T o18[18], double t, T* obs, T ref9[9],
                                          bool zero_ref = false) {
  T o[17];
  pos_invariant_obs(o18, o);
  if (task == kTaskStand) {
#pragma unroll
    for (int i = 0; i < 17; i++) obs[(size_t)e * 17 + i] = o[i];
  } else {
#pragma unroll
    for (int i = 0; i < 17; i++) obs[(size_t)e * 26 + i] = o[i];
    const int idx9[9] = {0, 1, 2, 3, 4, 6, 8, 9, 11};
    const int row = v.traj ? traj_index(t, v.traj_tmax, v.traj_rows) : 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      ref9[i] = v.traj ? (T)v.traj[(size_t)row * 13 + idx9[i]] : T(0);
      obs[(size_t)e * 26 + 17 + i] = zero_ref ? T(0) : ref9[i];
    }
  }
}

// mj_checkPos / mj_checkVel / mj_checkAcc [EXT]: MuJoCo resets the data when the state is not finite or has left
// any sane range.  A NaN env would otherwise never report done (every comparison with NaN is false) and would
// poison the batch statistics; here it is reported done, reset, and flagged in stats[3] (kStatusDiverged).
constexpr int kStatusDiverged = 3;
template <typename T>
__device__ __forceinline__ bool state_diverged(const T q[kNV], const T qd[kNV]) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < kNV; i++) bad = bad || !(Num<T>::abs_(q[i]) < T(1e10)) || !(Num<T>::abs_(qd[i]) < T(1e10));
  return bad;
}

// The part of Cassie2dEnv.step() after the n substeps (cassie_stand2d.py:104-137 / cassie2d.py:124-225) plus the
// batch conventions: observation, reward, termination, divergence guard, auto-reset.  One env; called by one lane.
//   flags: CASSIE_AUTO_RESET 1, CASSIE_FRESH_OBS_ON_RESET 2, CASSIE_LIVE_QSTATE 4, CASSIE_TERMINAL_OBS 8.
// With auto-reset the observation returned for a done env is the one env.reset() returns (the caller's next action
// must see the new episode), unless CASSIE_TERMINAL_OBS asks for the terminal one.
template <typename T>
__device__ __forceinline__ void env_finish(const ModelPair<T>& mp, const BatchView<T>& v, int task, int flags, int e,
                                           T q[kNV], T qd[kNV], T w[kNV], OpState<T>& op, double& t, unsigned& qps,
                                           const T* act, int adim, const T reset_state[26], T* obs_out, T* rew_out,
                                           uint8_t* done_out, OscStats& qs) {
  const bool diverged = state_diverged(q, qd);
  T o18[18], ref9[9], r;
  int done;
  op_state_array(op, q, qd, o18);
  if (diverged) {
#pragma unroll
    for (int i = 0; i < 18; i++) o18[i] = T(0);
  }
  write_obs(v, task, e, o18, t, obs_out, ref9);
  if (task == kTaskStand) {
    T o[17];
    pos_invariant_obs(o18, o);
    stand_reward(o18, o, act, adim, r, done);
  } else {
    const T jsum = (flags & 4) ? q[3] + q[4] + q[6] + q[8] + q[9] + q[11] : v.jsum0[e];
    imitate_reward(o18, ref9, jsum, r, done);
  }
  if (diverged) { r = T(0); done = 1; qs.status = kStatusDiverged; }
  rew_out[e] = r;
  done_out[e] = (uint8_t)done;
  if ((done && (flags & 1)) || diverged) {
    // env.reset(): Reset() + self.time = 0 (cassie2d.py:78-95).  Warm start and, unless
    // CASSIE_FRESH_OBS_ON_RESET, the lagged op-space state survive.
    state26_to_q(reset_state, q, qd);
    t = 0.0;
    v.jsum0[e] = q[3] + q[4] + q[6] + q[8] + q[9] + q[11];
    if ((flags & 2) || diverged) {
      Kin<T> kc;
      forward_kinematics(mp.ctrl, q, qd, kc);
      op_state_from_kin(mp.ctrl, kc, q, op);
    }
    if (diverged) {   // mj_resetData [EXT]: the solver state goes too
#pragma unroll
      for (int i = 0; i < kNV; i++) w[i] = T(0);
      qps = 0u;
    }
    if (!(flags & 8) || diverged) {
      op_state_array(op, q, qd, o18);
      write_obs(v, task, e, o18, 0.0, obs_out, ref9, true);
    }
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(block_threads(MODE), CASSIE_MIN_BLOCKS) k_env_step(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v,
                                                      const __grid_constant__ EnvStepDev<T> a) {
  const int e_raw = blockIdx.x * block_threads(MODE) + threadIdx.x;
  const bool active = e_raw < v.n;   // inactive lanes shadow the last env (they must reach the barriers)
  const int e = active ? e_raw : v.n - 1;
  T q[kNV], qd[kNV], w[kNV], act[7], u[kNU];
  load_env(v, e, q, qd, w);
  constexpr int adim = action_dim(MODE);
#pragma unroll
  for (int i = 0; i < adim; i++) act[i] = a.action[(size_t)e * adim + i];
  alignas(16) Rows<T> rows;
  OpState<T> op;
  StepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  if (a.n_sub <= 0) load_op(v, e, op);
  double t = v.clock[e];
  unsigned qps = v.qp_set[e];
  for (int s = 0; s < a.n_sub; s++) {
    step_barrier<block_threads(MODE)>();
    controller_step<MODE>(mp.phys, mp.phys_d, ctrl_model<MODE>(mp), mp.ctrl_d, q, qd, w, act, rows, u, s == a.n_sub - 1 ? &op : nullptr, &st, &qs, &qps);
    t += 0.0005;  // cassie2d.py:122
  }
  if (!active) return;
  env_finish(mp, v, a.task, a.flags, e, q, qd, w, op, t, qps, act, adim, a.reset_state, a.obs, a.reward, a.done, qs);
  v.clock[e] = t;
  v.qp_set[e] = qps;
  store_env(v, e, q, qd, w);
  store_op(v, e, op);
  store_stats(v.stats, v.n, e, st, qs);
}

// env.reset() for every env + its observation
template <typename T>
__global__ void __launch_bounds__(128) k_env_reset(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v, int task,
                                                   int flags, const T* __restrict__ state26, T* obs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n) return;
  T q[kNV], qd[kNV];
  state26_to_q(state26, q, qd);
  OpState<T> op;
  if (flags & 2) {
    Kin<T> kc;
    forward_kinematics(mp.ctrl, q, qd, kc);
    op_state_from_kin(mp.ctrl, kc, q, op);
    store_op(v, e, op);
  } else {
    load_op(v, e, op);
  }
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < kNV; i++) { v.qpos[(size_t)i * n + e] = q[i]; v.qvel[(size_t)i * n + e] = qd[i]; }
  v.clock[e] = 0.0;
  v.ep_len[e] = 0;
  v.jsum0[e] = q[3] + q[4] + q[6] + q[8] + q[9] + q[11];
  if (obs) {
    T o18[18], ref9[9];
    op_state_array(op, q, qd, o18);
    write_obs(v, task, e, o18, 0.0, obs, ref9, true);
  }
}

// ---------------------------------------------------------------------------------------
// squatting.py:8-16 with standing_controller_jacobian / standing_controller_osc in the loop
template <typename T, int MODE>
__global__ void __launch_bounds__(block_threads(MODE), CASSIE_MIN_BLOCKS) k_squat(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v,
                                                   const T* __restrict__ phase, int n_steps, uint32_t* mask) {
  const int e_raw = blockIdx.x * block_threads(MODE) + threadIdx.x;
  const bool active = e_raw < v.n;   // inactive lanes shadow the last env (they must reach the barriers)
  const int e = active ? e_raw : v.n - 1;
  T q[kNV], qd[kNV], w[kNV], act[7], u[kNU];
  load_env(v, e, q, qd, w);
  alignas(16) Rows<T> rows;
  OpState<T> op;
  load_op(v, e, op);
  StepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  double t = v.clock[e];
  const double ph = phase ? (double)phase[e] : 0.0;
  const double wq = 0.5 * 3.1415;  // squatting.py:9
  unsigned qps = v.qp_set[e];
  for (int s = 0; s < n_steps; s++) {
    step_barrier<block_threads(MODE)>();
    T o18[18];
    op_state_array(op, q, qd, o18);
    double sn, cs;
    sincos(wq * t + ph, &sn, &cs);
    const T zt = (T)(0.7 + 0.25 * sn), zdt = (T)(0.25 * cs);
    if (MODE == kModeJacobian) squat_jacobian_action(o18, zt, zdt, act);
    else squat_osc_action(o18, zt, zdt, act);
    controller_step<MODE>(mp.phys, mp.phys_d, ctrl_model<MODE>(mp), mp.ctrl_d, q, qd, w, act, rows, u, &op, &st, &qs, &qps);
    t = t + 0.0005;  // squatting.py:15
  }
  if (!active) return;
  v.clock[e] = t;
  v.qp_set[e] = qps;
  store_env(v, e, q, qd, w);
  store_op(v, e, op);
  if (n_steps > 0) {
    store_stats(v.stats, v.n, e, st, qs);
    if (mask) mask[e] = st.contact_mask;
  }
}

// ---------------------------------------------------------------------------------------
// Reset (Cassie2d.cpp:78-82): qpos/qvel only
template <typename T>
__global__ void __launch_bounds__(128) k_reset(const BatchView<T> v, const T* __restrict__ state26, int per_env,
                                               const uint8_t* __restrict__ mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n) return;
  if (mask && !mask[e]) return;
  T q[kNV], qd[kNV];
  state26_to_q(state26 + (per_env ? (size_t)e * 26 : 0), q, qd);
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < kNV; i++) { v.qpos[(size_t)i * n + e] = q[i]; v.qvel[(size_t)i * n + e] = qd[i]; }
}

template <typename T>
__global__ void __launch_bounds__(128) k_refresh_op(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v,
                                                    const uint8_t* __restrict__ mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n) return;
  if (mask && !mask[e]) return;
  T q[kNV], qd[kNV];
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < kNV; i++) { q[i] = v.qpos[(size_t)i * n + e]; qd[i] = v.qvel[(size_t)i * n + e]; }
  Kin<T> kc;
  forward_kinematics(mp.ctrl, q, qd, kc);
  OpState<T> op;
  op_state_from_kin(mp.ctrl, kc, q, op);
  store_op(v, e, op);
}

template <typename T>
__global__ void __launch_bounds__(128) k_get_general(const BatchView<T> v, T* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n) return;
  const int n = v.n;
  T* s = out + (size_t)e * 26;
#pragma unroll
  for (int i = 0; i < 3; i++) { s[i] = v.qpos[(size_t)i * n + e]; s[3 + i] = v.qvel[(size_t)i * n + e]; }
#pragma unroll
  for (int i = 0; i < 5; i++) {
    s[6 + i] = v.qpos[(size_t)(3 + i) * n + e];
    s[11 + i] = v.qvel[(size_t)(3 + i) * n + e];
    s[16 + i] = v.qpos[(size_t)(8 + i) * n + e];
    s[21 + i] = v.qvel[(size_t)(8 + i) * n + e];
  }
}

// qacc_warmstart in/out, real [n][13] (checkpoint / teacher-forced parity tests)
template <typename T>
__global__ void __launch_bounds__(128) k_warm_io(const BatchView<T> v, T* __restrict__ buf, int write_to_env) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n) return;
  const int n = v.n;
#pragma unroll
  for (int i = 0; i < kNV; i++) {
    if (write_to_env) v.warm[(size_t)i * n + e] = buf[(size_t)e * kNV + i];
    else buf[(size_t)e * kNV + i] = v.warm[(size_t)i * n + e];
  }
}

template <typename T>
__global__ void __launch_bounds__(128) k_get_op(const BatchView<T> v, T* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n) return;
  const int n = v.n;
  OpState<T> op;
  load_op(v, e, op);
  T q[kNV], qd[kNV];
#pragma unroll
  for (int i = 0; i < kNV; i++) { q[i] = T(0); qd[i] = T(0); }
  q[2] = v.qpos[(size_t)2 * n + e];
  qd[2] = v.qvel[(size_t)2 * n + e];
  T o[18];
  op_state_array(op, q, qd, o);
#pragma unroll
  for (int i = 0; i < 18; i++) out[(size_t)e * 18 + i] = o[i];
}

// ---------------------------------------------------------------------------------------
inline unsigned grid_for(int n, int block) { return (unsigned)((n + block - 1) / block); }

// thread-per-env launchers (launch.cuh picks between these and the quad engine's, quad_kernels.cuh)
template <typename T>
cudaError_t thread_step(const ModelPair<T>& mp, const BatchView<T>& v, const StepArgs& a, cudaStream_t s) {
  const T* act = (const T*)a.action;
  const int bt = block_threads(a.mode);
  const unsigned g = grid_for(v.n, bt);
  switch (a.mode) {
    case kModeTorque: k_step<T, kModeTorque><<<g, bt, 0, s>>>(mp, v, act, a.n_substeps, a.contact_mask); break;
    case kModePd: k_step<T, kModePd><<<g, bt, 0, s>>>(mp, v, act, a.n_substeps, a.contact_mask); break;
    case kModeJacobian: k_step<T, kModeJacobian><<<g, bt, 0, s>>>(mp, v, act, a.n_substeps, a.contact_mask); break;
    case kModeOsc: k_step<T, kModeOsc><<<g, bt, 0, s>>>(mp, v, act, a.n_substeps, a.contact_mask); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}

template <typename T>
EnvStepDev<T> make_env_step_dev(const EnvStepArgs& a) {
  EnvStepDev<T> d;
  d.task = a.task; d.flags = a.flags; d.n_sub = a.n_substeps;
  d.action = (const T*)a.action; d.obs = (T*)a.obs; d.reward = (T*)a.reward; d.done = a.done;
  // Python reset pose, cassie2d.py:79-85 / cassie_stand2d.py:76
  const double qi[26] = {0.0, 0.939, 0.0, 0.0, 0.0, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                         0.0, 0.0, 0.0, 0.0, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                         0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 26; i++) d.reset_state[i] = (T)qi[i];
  return d;
}
template <typename T>
cudaError_t thread_env_step(const ModelPair<T>& mp, const BatchView<T>& v, const EnvStepArgs& a, cudaStream_t s) {
  const EnvStepDev<T> d = make_env_step_dev<T>(a);
  const int bt = block_threads(a.mode);
  const unsigned g = grid_for(v.n, bt);
  switch (a.mode) {
    case kModeTorque: k_env_step<T, kModeTorque><<<g, bt, 0, s>>>(mp, v, d); break;
    case kModePd: k_env_step<T, kModePd><<<g, bt, 0, s>>>(mp, v, d); break;
    case kModeOsc: k_env_step<T, kModeOsc><<<g, bt, 0, s>>>(mp, v, d); break;
    default: return cudaErrorInvalidValue;  // the Python envs have no Jacobian action space
  }
  count_launch();
  return cudaGetLastError();
}

template <typename T>
cudaError_t thread_squat(const ModelPair<T>& mp, const BatchView<T>& v, const SquatArgs& a, cudaStream_t s) {
  const int bt = block_threads(a.mode);
  const unsigned g = grid_for(v.n, bt);
  if (a.mode == kModeJacobian) k_squat<T, kModeJacobian><<<g, bt, 0, s>>>(mp, v, (const T*)a.phase, a.n_steps, a.contact_mask);
  else if (a.mode == kModeOsc) k_squat<T, kModeOsc><<<g, bt, 0, s>>>(mp, v, (const T*)a.phase, a.n_steps, a.contact_mask);
  else return cudaErrorInvalidValue;
  count_launch();
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::reset(const BatchView<T>& v, const T* state26, int per_env, const uint8_t* mask, cudaStream_t s) {
  k_reset<T><<<grid_for(v.n, 128), 128, 0, s>>>(v, state26, per_env, mask);
  count_launch();
  return cudaGetLastError();
}
template <typename T>
cudaError_t Launch<T>::refresh_op(const ModelPair<T>& mp, const BatchView<T>& v, const uint8_t* mask, cudaStream_t s) {
  k_refresh_op<T><<<grid_for(v.n, 128), 128, 0, s>>>(mp, v, mask);
  count_launch();
  return cudaGetLastError();
}
template <typename T>
cudaError_t Launch<T>::get_general(const BatchView<T>& v, T* state26, cudaStream_t s) {
  k_get_general<T><<<grid_for(v.n, 128), 128, 0, s>>>(v, state26);
  count_launch();
  return cudaGetLastError();
}
template <typename T>
cudaError_t Launch<T>::get_op(const BatchView<T>& v, T* state18, cudaStream_t s) {
  k_get_op<T><<<grid_for(v.n, 128), 128, 0, s>>>(v, state18);
  count_launch();
  return cudaGetLastError();
}
template <typename T>
cudaError_t Launch<T>::warm_io(const BatchView<T>& v, T* buf, int write_to_env, cudaStream_t s) {
  k_warm_io<T><<<grid_for(v.n, 128), 128, 0, s>>>(v, buf, write_to_env);
  count_launch();
  return cudaGetLastError();
}
template <typename T>
cudaError_t Launch<T>::env_reset(const ModelPair<T>& mp, const BatchView<T>& v, int task, int flags, const T* state26,
                                 T* obs, cudaStream_t s) {
  k_env_reset<T><<<grid_for(v.n, 128), 128, 0, s>>>(mp, v, task, flags, state26, obs);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cassie
