// Engine selection for the fused rollout launch (see launch.cuh)
#pragma once
#include "launch.cuh"
#include "quad_rollout.cuh"

namespace cassie {

template <typename T>
cudaError_t launch_rollout(const ModelPair<T>& mp, const BatchView<T>& v, const RolloutArgs& a, cudaStream_t s) {
  if (!engine_for_mode(a.mode) || v.n < kQuadMinEnvs) return thread_rollout<T>(mp, v, a, s);
  const RolloutDev<T> d = make_rollout_dev<T>(a);
  cudaError_t e;
  switch (a.mode) {
    case kModeTorque: e = quad::launch_qrollout<T, kModeTorque>(mp, v, d, s); break;
    case kModePd: e = quad::launch_qrollout<T, kModePd>(mp, v, d, s); break;
    case kModeOsc: e = quad::launch_qrollout<T, kModeOsc>(mp, v, d, s); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return e;
}

}  // namespace cassie
