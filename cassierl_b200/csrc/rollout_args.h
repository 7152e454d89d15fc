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
}  // namespace cassie
