// Device-resident state of a batch of Cassie2d envs (structure-of-arrays in HBM) and the
// launchers the C-ABI (cassie2d_api.cu) calls.  Replaces the per-instance mjData + RBDL state
// + qpOASES hot-start state owned by the reference's Cassie2d object
// (CassieRL/cassierl src/Cassie2d/Cassie2d.h:18-50).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "planar_model.h"

namespace cassie {

template <typename T>
struct ModelPair {
  PlanarModel<T> phys;  // MuJoCo's view of the MJCF (mj_loadXML, Cassie2d.cpp:48)
  PlanarModel<T> ctrl;  // the RBDL loader's view (DynamicModel::LoadModel, Cassie2d.cpp:43)
  PlanarModel<double> phys_d;  // phys in double: the position pass of every step runs in double (planar_engine.cuh)
  PlanarModel<double> ctrl_d;  // the same in double: the OSC controller always runs in double (osc_qp.cuh)
};

// All arrays are [field][env] (env fastest) so that a warp's 32 envs read one 128-byte line
// per field.
template <typename T>
struct BatchView {
  int n;
  T* qpos;          // [13][n]  mjData.qpos
  T* qvel;          // [13][n]  mjData.qvel
  T* warm;          // [13][n]  mjData.qacc_warmstart
  T* op;            // [12][n]  lagged op-space state (RBDL state of the last Step*, App. D.1)
  double* clock;    // [n]      squatting.py's t  /  cassie2d.py's self.time (seconds, double)
  T* jsum0;         // [n]      frozen qstate joint sum of the imitation reward (App. D.4)
  uint32_t* qp_set; // [n]      OSC QP partition (free / at-lower / at-upper) carried across steps, like
                    //          the qpOASES hot start that survives resets in the reference (App. D.3)
  int32_t* ep_len;      // [n]  policy steps since the last reset (max_path_length bookkeeping of the sampler)
  int32_t* policy_step; // [n]  policy steps since init: the Philox counter of the action noise
  int32_t* stats;   // [n][4]   rows, PGS sweeps, QP iterations, QP status of the last substep
  const double* traj;  // [traj_rows][13] reference qpos (device), may be null
  int traj_rows;
  double traj_tmax;
};

struct StepArgs {
  int mode, n_substeps;
  const void* action;       // real [n][adim]
  uint32_t* contact_mask;   // [n] or null
};
struct EnvStepArgs {
  int task, mode, n_substeps, flags;
  const void* action;
  void* obs;
  void* reward;
  uint8_t* done;
};
struct SquatArgs {
  int mode, n_steps;
  const void* phase;        // real [n] or null
  uint32_t* contact_mask;
};

// launchers, one explicit instantiation per precision (kernels_f32.cu / kernels_f64.cu)
template <typename T>
struct Launch {
  static cudaError_t step(const ModelPair<T>& mp, const BatchView<T>& v, const StepArgs& a, cudaStream_t s);
  static cudaError_t env_step(const ModelPair<T>& mp, const BatchView<T>& v, const EnvStepArgs& a, cudaStream_t s);
  static cudaError_t squat(const ModelPair<T>& mp, const BatchView<T>& v, const SquatArgs& a, cudaStream_t s);
  // Reset: state26 (device, T[26] single state or T[n][26] per env when per_env), mask may be null
  static cudaError_t reset(const BatchView<T>& v, const T* state26, int per_env, const uint8_t* mask, cudaStream_t s);
  // refresh the lagged op-space state from the current qpos/qvel (constructor's setState, Cassie2d.cpp:64)
  static cudaError_t refresh_op(const ModelPair<T>& mp, const BatchView<T>& v, const uint8_t* mask, cudaStream_t s);
  static cudaError_t get_general(const BatchView<T>& v, T* state26, cudaStream_t s);
  static cudaError_t get_op(const BatchView<T>& v, T* state18, cudaStream_t s);
  static cudaError_t warm_io(const BatchView<T>& v, T* buf, int write_to_env, cudaStream_t s);
  static cudaError_t env_reset(const ModelPair<T>& mp, const BatchView<T>& v, int task, int flags, const T* state26,
                               T* obs, cudaStream_t s);
};

struct RolloutArgs;
template <typename T>
cudaError_t launch_rollout(const ModelPair<T>& mp, const BatchView<T>& v, const RolloutArgs& a, cudaStream_t s);

template <typename T>
cudaError_t launch_discounted_returns(const void* rew, const uint8_t* done, const void* tail, double gamma, int T_steps, int n,
                                      void* ret, cudaStream_t s);

struct BaselineArgs;
template <typename T>
cudaError_t launch_baseline_moments(const BaselineArgs& a, cudaStream_t s);
template <typename T>
cudaError_t launch_advantages(const BaselineArgs& a, cudaStream_t s);

long long kernel_launch_count();
void count_launch();

}  // namespace cassie
