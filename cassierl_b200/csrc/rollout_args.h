// Host-side arguments of the fused rollout launch (rollout_kernels.cuh)
#pragma once
#include <cstdint>
namespace cassie {
struct RolloutArgs {
  int task, mode, n_substeps, T_steps, max_path_length, flags, normalize;
  uint64_t seed;
  uint32_t env0;
  const void* params;
  void* obs; void* act; void* mean; void* rew; uint8_t* done;
  double act_lo[7], act_hi[7];
};
// arguments of the baseline-moment / advantage launches (rollout_kernels.cuh)
struct BaselineArgs {
  const void* obs; const void* rew; const void* ret; const uint8_t* done; const int32_t* start; int32_t* idx;
  const void* coeffs; void* adv; void* value; double* moments; int odim, T_steps, n; double gamma, lambda;
};
}  // namespace cassie
