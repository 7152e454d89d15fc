// CUDA kernels of the quad engine (four lanes per env, eight envs per warp, one warp per CTA): the same launches as
// env_kernels.cuh -- n x Step* with a held action, the Python env's policy step, the squatting.py loop -- on the
// cooperative mapping of quad_engine.cuh / quad_ctrl.cuh.  Env state is staged in shared memory for the whole launch
// (all substeps fused); per-env scratch is a shared block that the controller and the physics step use in turn.
#pragma once
#include "env_kernels.cuh"
#include "quad_ctrl.cuh"

namespace cassie {
namespace quad {

constexpr int kEnvsPerWarp = 8;
// Warps per CTA.  The physics-only modes run one warp per CTA (fourteen independent CTAs per SM); the controller modes
// run all the warps an SM can hold (seven, shared memory bound) as ONE CTA kept in lock step by a barrier per simulator
// step: their once-per-step code is straight-line and several times the size of the instruction cache, and in lock step
// one fill serves seven warps (profiles/r2k: squat_jacobian +27 %, squat_osc +8 %; pd_env -6 % with the barrier).
#ifndef CASSIE_QUAD_WARPS
#define CASSIE_QUAD_WARPS 1
#endif
#ifndef CASSIE_QUAD_WARPS_CTRL
#define CASSIE_QUAD_WARPS_CTRL 7
#endif
constexpr int quad_warps(int mode) { return mode >= kModeJacobian ? CASSIE_QUAD_WARPS_CTRL : CASSIE_QUAD_WARPS; }
constexpr int quad_block(int mode) { return 32 * quad_warps(mode); }

// scratch bytes per env: the controller's block only exists in the Jacobian / OSC modes
template <typename T>
constexpr size_t scratch_bytes(int mode) {
  const size_t p = sizeof(T) * PhysLayout::end, c = sizeof(TC) * CtrlLayout::end;
  return mode >= kModeJacobian ? (p > c ? p : c) : p;
}
template <typename T>
constexpr size_t warp_bytes(int mode) {
  return kEnvsPerWarp * (sizeof(T) * StateLayout::end + scratch_bytes<T>(mode));
}
// resident CTAs per SM the register allocation is sized for (228 KB of shared memory per SM, 1 KB reserved per CTA)
template <typename T>
constexpr int min_blocks(int mode) {
  const size_t per = quad_warps(mode) * warp_bytes<T>(mode) + 1024;
  const int b = (int)((228 * 1024) / per);
  return b > 14 ? 14 : (b < 1 ? 1 : b);   // 16384 envs = 13.8 warps per SM: more resident slots than that only cost registers
}

// several warps per CTA (CASSIE_QUAD_WARPS): a barrier per simulator step keeps them in the same code region, so that
// one instruction-cache fill serves all of them (the once-per-step code is straight-line and far larger than the I-cache)
template <int MODE>
__device__ __forceinline__ void step_sync() {
  if (quad_warps(MODE) > 1) __syncthreads();
}

struct QuadEnv {
  int e, ei;
  bool active;
  Lane ln;
};
template <int MODE, typename T>
__device__ __forceinline__ QuadEnv quad_env(const BatchView<T>& v) {
  QuadEnv q;
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  q.ei = lane >> 2;
  const int e_raw = ((int)blockIdx.x * quad_warps(MODE) + warp) * kEnvsPerWarp + q.ei;
  q.active = e_raw < v.n;      // inactive quads shadow the last env (they take part in the warp collectives)
  q.e = q.active ? e_raw : v.n - 1;
  q.ln = lane_id();
  return q;
}
template <typename T, int MODE>
__device__ __forceinline__ unsigned char* warp_smem() {
  extern __shared__ __align__(16) unsigned char quad_smem[];
  return quad_smem + (threadIdx.x >> 5) * warp_bytes<T>(MODE);
}

// global SoA [field][env] <-> the state block; the four lanes of a quad take every fourth field
template <typename T>
__device__ __forceinline__ void load_state(const BatchView<T>& v, const QuadEnv& qe, SV<T> St) {
  typedef StateLayout X;
  const size_t n = (size_t)v.n;
  for (int i = qe.ln.ql; i < kNV; i += 4) {
    St[X::q + i] = v.qpos[i * n + qe.e];
    St[X::qd + i] = v.qvel[i * n + qe.e];
    St[X::warm + i] = v.warm[i * n + qe.e];
  }
  for (int i = qe.ln.ql; i < 12; i += 4) St[X::op + i] = v.op[i * n + qe.e];
  __syncwarp();
}
template <typename T>
__device__ __forceinline__ void store_state(const BatchView<T>& v, const QuadEnv& qe, SV<T> St) {
  typedef StateLayout X;
  __syncwarp();
  if (!qe.active) return;
  const size_t n = (size_t)v.n;
  for (int i = qe.ln.ql; i < kNV; i += 4) {
    v.qpos[i * n + qe.e] = St[X::q + i];
    v.qvel[i * n + qe.e] = St[X::qd + i];
    v.warm[i * n + qe.e] = St[X::warm + i];
  }
  for (int i = qe.ln.ql; i < 12; i += 4) v.op[i * n + qe.e] = St[X::op + i];
}

template <typename T, int MODE, bool CTA = false>
__device__ __forceinline__ void quad_step(const ModelPair<T>& mp, const QuadEnv& qe, SV<T> St, unsigned char* wb, const T* act,
                                          bool want_op, QStepStats* st, OscStats* qs, unsigned* qps) {
  quad_controller_step<MODE, CTA>(mp.phys, mp.phys_d, mp.ctrl_d, qe.ln, St, wb + kEnvsPerWarp * sizeof(T) * StateLayout::end, qe.ei,
                             act, want_op, st, qs, qps);
}

// ---------------------------------------------------------------------------------------
// n_substeps x Step* (Cassie2d.cpp:86-209) with a held action
template <typename T, int MODE>
__global__ void __launch_bounds__(quad_block(MODE), min_blocks<T>(MODE))
k_qstep(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v, const T* __restrict__ action, int n_sub, uint32_t* mask) {
  const QuadEnv qe = quad_env<MODE>(v);
  unsigned char* wb = warp_smem<T, MODE>();
  const SV<T> St{reinterpret_cast<T*>(wb) + qe.ei};
  load_state(v, qe, St);
  constexpr int adim = action_dim(MODE);
  T act[7];
#pragma unroll
  for (int i = 0; i < adim; i++) act[i] = action[(size_t)qe.e * adim + i];
  QStepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  unsigned qps = v.qp_set[qe.e];
  for (int s = 0; s < n_sub; s++) {
    step_sync<MODE>();
    quad_step<T, MODE, true>(mp, qe, St, wb, act, s == n_sub - 1, &st, &qs, &qps);
  }
  store_state(v, qe, St);
  if (qe.active && qe.ln.ql == 0) {
    v.qp_set[qe.e] = qps;
    if (n_sub > 0) {
      store_stats(v.stats, v.n, qe.e, StepStats{st.nrows, st.sweeps, st.contact_mask}, qs);
      if (mask) mask[qe.e] = st.contact_mask;
    }
  }
}

// ---------------------------------------------------------------------------------------
// squatting.py:8-16 with standing_controller_jacobian / standing_controller_osc in the loop
template <typename T, int MODE>
__global__ void __launch_bounds__(quad_block(MODE), min_blocks<T>(MODE))
k_qsquat(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v, const T* __restrict__ phase, int n_steps, uint32_t* mask) {
  const QuadEnv qe = quad_env<MODE>(v);
  unsigned char* wb = warp_smem<T, MODE>();
  const SV<T> St{reinterpret_cast<T*>(wb) + qe.ei};
  load_state(v, qe, St);
  QStepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  double t = v.clock[qe.e];
  const double ph = phase ? (double)phase[qe.e] : 0.0;
  const double wq = 0.5 * 3.1415;  // squatting.py:9
  unsigned qps = v.qp_set[qe.e];
  for (int s = 0; s < n_steps; s++) {
    step_sync<MODE>();
    T o18[18], act[7];
    quad_op_array(St, o18);
    double sn, cs;
    sincos(wq * t + ph, &sn, &cs);
    const T zt = (T)(0.7 + 0.25 * sn), zdt = (T)(0.25 * cs);
    if (MODE == kModeJacobian) squat_jacobian_action(o18, zt, zdt, act);
    else squat_osc_action(o18, zt, zdt, act);
    __syncwarp();   // every lane has read the lagged op-space state before the step rewrites it
    quad_step<T, MODE>(mp, qe, St, wb, act, true, &st, &qs, &qps);
    t = t + 0.0005;  // squatting.py:15
  }
  store_state(v, qe, St);
  if (qe.active && qe.ln.ql == 0) {
    v.clock[qe.e] = t;
    v.qp_set[qe.e] = qps;
    if (n_steps > 0) {
      store_stats(v.stats, v.n, qe.e, StepStats{st.nrows, st.sweeps, st.contact_mask}, qs);
      if (mask) mask[qe.e] = st.contact_mask;
    }
  }
}

// ---------------------------------------------------------------------------------------
// One policy step of the Python env (cassie_stand2d.py:86-137 / cassie2d.py:97-225): the substeps by the quad, the
// observation / reward / termination / auto-reset arithmetic (env_kernels.cuh env_finish) by its lane 0
template <typename T, int MODE>
__global__ void __launch_bounds__(quad_block(MODE), min_blocks<T>(MODE))
k_qenv_step(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v, const __grid_constant__ EnvStepDev<T> a) {
  typedef StateLayout X;
  const QuadEnv qe = quad_env<MODE>(v);
  unsigned char* wb = warp_smem<T, MODE>();
  const SV<T> St{reinterpret_cast<T*>(wb) + qe.ei};
  load_state(v, qe, St);
  constexpr int adim = action_dim(MODE);
  T act[7];
#pragma unroll
  for (int i = 0; i < adim; i++) act[i] = a.action[(size_t)qe.e * adim + i];
  QStepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  double t = v.clock[qe.e];
  unsigned qps = v.qp_set[qe.e];
  for (int s = 0; s < a.n_sub; s++) {
    step_sync<MODE>();
    quad_step<T, MODE, true>(mp, qe, St, wb, act, s == a.n_sub - 1, &st, &qs, &qps);
    t += 0.0005;  // cassie2d.py:122
  }
  __syncwarp();
  if (qe.active && qe.ln.ql == 0) {
    T q[kNV], qd[kNV], w[kNV];
    OpState<T> op;
#pragma unroll
    for (int i = 0; i < kNV; i++) { q[i] = St[X::q + i]; qd[i] = St[X::qd + i]; w[i] = St[X::warm + i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) { op.body[i] = St[X::op + i]; op.left[i] = St[X::op + 4 + i]; op.right[i] = St[X::op + 8 + i]; }
    env_finish(mp, v, a.task, a.flags, qe.e, q, qd, w, op, t, qps, act, adim, a.reset_state, a.obs, a.reward, a.done, qs);
#pragma unroll
    for (int i = 0; i < kNV; i++) { St[X::q + i] = q[i]; St[X::qd + i] = qd[i]; St[X::warm + i] = w[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) { St[X::op + i] = op.body[i]; St[X::op + 4 + i] = op.left[i]; St[X::op + 8 + i] = op.right[i]; }
    v.clock[qe.e] = t;
    v.qp_set[qe.e] = qps;
    store_stats(v.stats, v.n, qe.e, StepStats{st.nrows, st.sweeps, st.contact_mask}, qs);
  }
  store_state(v, qe, St);
}

// ---------------------------------------------------------------------------------------
template <typename K>
inline void prefer_shared(K kernel, size_t dyn_bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (dyn_bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes);
}

inline unsigned quad_grid(int n, int mode) { return (unsigned)((n + kEnvsPerWarp * quad_warps(mode) - 1) / (kEnvsPerWarp * quad_warps(mode))); }

template <typename T, int MODE>
inline cudaError_t launch_qstep(const ModelPair<T>& mp, const BatchView<T>& v, const T* act, int n_sub, uint32_t* mask, cudaStream_t s) {
  static bool once = (prefer_shared(k_qstep<T, MODE>, quad_warps(MODE) * warp_bytes<T>(MODE)), true);
  (void)once;
  k_qstep<T, MODE><<<quad_grid(v.n, MODE), quad_block(MODE), quad_warps(MODE) * warp_bytes<T>(MODE), s>>>(mp, v, act, n_sub, mask);
  return cudaGetLastError();
}
template <typename T, int MODE>
inline cudaError_t launch_qsquat(const ModelPair<T>& mp, const BatchView<T>& v, const T* phase, int n_steps, uint32_t* mask, cudaStream_t s) {
  static bool once = (prefer_shared(k_qsquat<T, MODE>, quad_warps(MODE) * warp_bytes<T>(MODE)), true);
  (void)once;
  k_qsquat<T, MODE><<<quad_grid(v.n, MODE), quad_block(MODE), quad_warps(MODE) * warp_bytes<T>(MODE), s>>>(mp, v, phase, n_steps, mask);
  return cudaGetLastError();
}
template <typename T, int MODE>
inline cudaError_t launch_qenv_step(const ModelPair<T>& mp, const BatchView<T>& v, const EnvStepDev<T>& d, cudaStream_t s) {
  static bool once = (prefer_shared(k_qenv_step<T, MODE>, quad_warps(MODE) * warp_bytes<T>(MODE)), true);
  (void)once;
  k_qenv_step<T, MODE><<<quad_grid(v.n, MODE), quad_block(MODE), quad_warps(MODE) * warp_bytes<T>(MODE), s>>>(mp, v, d);
  return cudaGetLastError();
}

}  // namespace quad
}  // namespace cassie
