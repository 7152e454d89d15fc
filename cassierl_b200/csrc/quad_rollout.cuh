// Rollout collection on the quad engine (BASELINE configs[4]; rollout_kernels.cuh has the thread-per-env version and the
// description of what rllab's sampler does around the reference env, rllab/envs/trpo_cassie.py:13-55).
// Per policy step and env: observation -> tanh MLP (26|17 -> 32 -> 32 -> adim), each lane of the quad evaluating eight
// hidden units per layer with the activations exchanged through the shared scratch block -> a = mean + exp(log_std) eps
// (Philox4x32-10 keyed by seed / global env id / policy step) -> NormalizedEnv affine map + clip -> n_substeps quad steps
// -> reward / done / max_path_length bookkeeping and reset by lane 0.  Weights are read through the read-only path
// (8.7 KB, L1 resident): staging them in shared memory would cost the physics-only modes their occupancy.
#pragma once
#include "quad_kernels.cuh"
#include "rollout_kernels.cuh"

namespace cassie {
namespace quad {

// scratch words (type T) used by the policy between two env steps; they overlay the step's scratch block
struct PolicyLayout {
  static constexpr int obs = 0, h1 = 26, h2 = h1 + kHidden, act = h2 + kHidden, end = act + 8;
};
static_assert(PolicyLayout::end <= PhysLayout::end, "the policy's activations must fit in the physics scratch");

template <typename T, int MODE>
__global__ void __launch_bounds__(quad_block(MODE), min_blocks<T>(MODE))
k_qrollout(const __grid_constant__ ModelPair<T> mp, const BatchView<T> v, const __grid_constant__ RolloutDev<T> a) {
  typedef StateLayout X;
  typedef PolicyLayout Y;
  constexpr int adim = action_dim(MODE);
  const int odim = a.task == kTaskStand ? 17 : 26;
  const QuadEnv qe = quad_env<MODE>(v);
  unsigned char* wb = warp_smem<T, MODE>();
  const SV<T> St{reinterpret_cast<T*>(wb) + qe.ei};
  const SV<T> W{reinterpret_cast<T*>(wb + kEnvsPerWarp * sizeof(T) * X::end) + qe.ei};
  load_state(v, qe, St);
  const T* __restrict__ W1 = a.params; const T* __restrict__ b1 = W1 + odim * kHidden; const T* __restrict__ W2 = b1 + kHidden;
  const T* __restrict__ b2 = W2 + kHidden * kHidden; const T* __restrict__ W3 = b2 + kHidden; const T* __restrict__ b3 = W3 + kHidden * adim;
  const T* __restrict__ lstd = b3 + adim;
  const int e = qe.e, l = qe.ln.ql;
  const size_t n = (size_t)v.n;
  QStepStats st = {0, 0, 0u};
  OscStats qs = {0, 0};
  double t = v.clock[e];
  unsigned qps = v.qp_set[e];
  int ep_len = v.ep_len[e];
  uint32_t pstep = (uint32_t)v.policy_step[e];
  int n_diverged = 0;

  for (int k = 0; k < a.T_steps; k++) {
    // ---- observation of the current state (what the previous step / reset returned), by lane 0
    const bool fresh = ep_len == 0;   // the observation env.reset() returned: reference slots are zero (cassie2d.py:78-95)
    if (l == 0) {
      T o18[18], o[17], ref9[9];
      quad_op_array(St, o18);
      if (qe.active) write_obs(v, a.task, e, o18, t, a.obs + (size_t)k * n * odim, ref9, fresh);
      else {
        const int idx9[9] = {0, 1, 2, 3, 4, 6, 8, 9, 11};
        const int row = v.traj ? traj_index(t, v.traj_tmax, v.traj_rows) : 0;
#pragma unroll
        for (int i = 0; i < 9; i++) ref9[i] = (v.traj && a.task != kTaskStand) ? (T)v.traj[(size_t)row * 13 + idx9[i]] : T(0);
      }
      pos_invariant_obs(o18, o);
#pragma unroll
      for (int i = 0; i < 17; i++) W[Y::obs + i] = o[i];
      if (a.task != kTaskStand) {
#pragma unroll
        for (int i = 0; i < 9; i++) W[Y::obs + 17 + i] = fresh ? T(0) : ref9[i];
      }
    }
    __syncwarp();
    // ---- GaussianMLPPolicy forward: tanh hidden layers, linear mean head (trpo_cassie.py:21-27); lane l owns units 8 l .. 8 l + 7
    {
      T h[8];
#pragma unroll
      for (int j = 0; j < 8; j++) h[j] = __ldg(b1 + 8 * l + j);
      for (int i = 0; i < odim; i++) {
        const T x = W[Y::obs + i];
#pragma unroll
        for (int j = 0; j < 8; j++) h[j] += x * __ldg(W1 + i * kHidden + 8 * l + j);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) W[Y::h1 + 8 * l + j] = tanh(h[j]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; j++) h[j] = __ldg(b2 + 8 * l + j);
#pragma unroll 4
      for (int i = 0; i < kHidden; i++) {
        const T x = W[Y::h1 + i];
#pragma unroll
        for (int j = 0; j < 8; j++) h[j] += x * __ldg(W2 + i * kHidden + 8 * l + j);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) W[Y::h2 + 8 * l + j] = tanh(h[j]);
      __syncwarp();
    }
    T act[7];
    {
      // mean head: every lane evaluates all outputs (adim <= 7 dots of 32: cheaper than another exchange)
      T mu[adim];
#pragma unroll
      for (int c = 0; c < adim; c++) mu[c] = __ldg(b3 + c);
#pragma unroll 4
      for (int i = 0; i < kHidden; i++) {
        const T x = W[Y::h2 + i];
#pragma unroll
        for (int c = 0; c < adim; c++) mu[c] += x * __ldg(W3 + i * adim + c);
      }
      // ---- a = mean + exp(log_std) * eps ; NormalizedEnv: lb + (a + 1) / 2 (ub - lb), clipped (trpo_cassie.py:13)
      T eps[8];
      normal8(a.seed, a.env0 + (uint32_t)e, pstep, eps);
#pragma unroll
      for (int c = 0; c < adim; c++) {
        const T raw = mu[c] + exp(__ldg(lstd + c)) * eps[c];
        if (l == 0 && qe.active) {
          a.act[((size_t)k * n + e) * adim + c] = raw;
          a.mean[((size_t)k * n + e) * adim + c] = mu[c];
        }
        T x = raw;
        if (a.normalize) x = a.act_lo[c] + (raw + T(1)) * T(0.5) * (a.act_hi[c] - a.act_lo[c]);
        act[c] = x < a.act_lo[c] ? a.act_lo[c] : (x > a.act_hi[c] ? a.act_hi[c] : x);
      }
    }
    __syncwarp();   // the activations are dead: the step may reuse the scratch block
    // ---- Cassie2dEnv.step(action, n)
    for (int s = 0; s < a.n_sub; s++) {
      step_sync<MODE>();
      quad_step<T, MODE, true>(mp, qe, St, wb, act, s == a.n_sub - 1, &st, &qs, &qps);
      t += 0.0005;
    }
    __syncwarp();
    // ---- reward, termination, sampler bookkeeping, reset: lane 0 on the state block
    ep_len++;
    pstep++;
    int flag = 0;
    bool diverged = false;
    if (l == 0) {
      T q[kNV], qd[kNV];
#pragma unroll
      for (int i = 0; i < kNV; i++) { q[i] = St[X::q + i]; qd[i] = St[X::qd + i]; }
      diverged = state_diverged(q, qd);   // mj_checkPos/Vel/Acc [EXT]: report done, reset, flag in stats
      T o18[18], r;
      int done;
      quad_op_array(St, o18);
      if (a.task == kTaskStand) {
        T o[17];
        pos_invariant_obs(o18, o);
        stand_reward(o18, o, act, adim, r, done);
      } else {
        T ref9[9];
        const int idx9[9] = {0, 1, 2, 3, 4, 6, 8, 9, 11};
        const int row = v.traj ? traj_index(t, v.traj_tmax, v.traj_rows) : 0;
#pragma unroll
        for (int i = 0; i < 9; i++) ref9[i] = v.traj ? (T)v.traj[(size_t)row * 13 + idx9[i]] : T(0);
        const T jsum = (a.flags & 4) ? q[3] + q[4] + q[6] + q[8] + q[9] + q[11] : v.jsum0[e];
        imitate_reward(o18, ref9, jsum, r, done);
      }
      if (diverged) { r = T(0); done = 1; }
      flag = done ? 1 : (ep_len >= a.max_path_length ? 2 : 0);
      if (qe.active) {
        a.rew[(size_t)k * n + e] = r;
        a.done[(size_t)k * n + e] = (uint8_t)flag;
      }
      if (flag) {  // rollout(): the path ends, the sampler calls env.reset()
        state26_to_q(a.reset_state, q, qd);
#pragma unroll
        for (int i = 0; i < kNV; i++) { St[X::q + i] = q[i]; St[X::qd + i] = qd[i]; }
        if (qe.active) v.jsum0[e] = q[3] + q[4] + q[6] + q[8] + q[9] + q[11];
        if ((a.flags & 2) || diverged) {
          Kin<T> kc;
          OpState<T> op;
          forward_kinematics(mp.ctrl, q, qd, kc);
          op_state_from_kin(mp.ctrl, kc, q, op);
#pragma unroll
          for (int i = 0; i < 4; i++) { St[X::op + i] = op.body[i]; St[X::op + 4 + i] = op.left[i]; St[X::op + 8 + i] = op.right[i]; }
        }
        if (diverged) {
#pragma unroll
          for (int i = 0; i < kNV; i++) St[X::warm + i] = T(0);
        }
      }
    }
    flag = __shfl_sync(0xffffffffu, flag, 0, 4);
    diverged = __shfl_sync(0xffffffffu, (int)diverged, 0, 4) != 0;
    if (flag) { t = 0.0; ep_len = 0; }
    if (diverged) { qps = 0u; n_diverged++; qs.status = kStatusDiverged; }
    __syncwarp();
  }
  store_state(v, qe, St);
  if (qe.active && l == 0) {
    v.clock[e] = t;
    v.qp_set[e] = qps;
    v.ep_len[e] = ep_len;
    v.policy_step[e] = (int32_t)pstep;
    if (n_diverged) qs.status = kStatusDiverged;   // sticky for the launch: the env diverged and was reset at least once
    store_stats(v.stats, v.n, e, StepStats{st.nrows, st.sweeps, st.contact_mask}, qs);
  }
}

template <typename T, int MODE>
inline cudaError_t launch_qrollout(const ModelPair<T>& mp, const BatchView<T>& v, const RolloutDev<T>& d, cudaStream_t s) {
  static bool once = (prefer_shared(k_qrollout<T, MODE>, quad_warps(MODE) * warp_bytes<T>(MODE)), true);
  (void)once;
  k_qrollout<T, MODE><<<quad_grid(v.n, MODE), quad_block(MODE), quad_warps(MODE) * warp_bytes<T>(MODE), s>>>(mp, v, d);
  return cudaGetLastError();
}

}  // namespace quad
}  // namespace cassie
