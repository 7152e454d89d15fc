// Engine selection for the step launches.  Two complete engines live in this library:
//   quad    four lanes per env, shared-memory scratch        -- quad_engine.cuh, quad_ctrl.cuh, quad_kernels.cuh
//   thread  one env per thread, thread-local scratch (round 1) -- planar_engine.cuh ... env_kernels.cuh
// A batch of fewer than kQuadMinEnvs envs (the legacy batch-of-one ABI) always runs the thread engine.  Both engines
// are checked against the oracle by the same tests (tests/test_gpu_engines.py re-runs the parity files per engine).
#pragma once
#include <cstdlib>
#include <cstring>
#include "quad_kernels.cuh"

namespace cassie {

constexpr int kQuadMinEnvs = 2;
// CASSIE_ENGINE = quad | thread forces one engine for every mode; unset = the measured default per control mode
// (profiles/r2_engines.txt): the quad engine for the controller modes (Jacobian, OSC), the thread engine for the
// physics-only modes (torque, PD), where robots thrown around by random actions spend their time in the general
// constraint tier that the quad engine steps serially.
inline int engine_choice() {
  static const int c = [] {
    const char* e = getenv("CASSIE_ENGINE");
    if (e && strcmp(e, "thread") == 0) return 0;
    if (e && strcmp(e, "quad") == 0) return 1;
    return 2;
  }();
  return c;
}
inline bool engine_for_mode(int mode) {
  const int c = engine_choice();
  return c == 1 || (c == 2 && mode >= kModeJacobian);
}

template <typename T>
cudaError_t Launch<T>::step(const ModelPair<T>& mp, const BatchView<T>& v, const StepArgs& a, cudaStream_t s) {
  if (!engine_for_mode(a.mode) || v.n < kQuadMinEnvs) return thread_step<T>(mp, v, a, s);
  const T* act = (const T*)a.action;
  cudaError_t e;
  switch (a.mode) {
    case kModeTorque: e = quad::launch_qstep<T, kModeTorque>(mp, v, act, a.n_substeps, a.contact_mask, s); break;
    case kModePd: e = quad::launch_qstep<T, kModePd>(mp, v, act, a.n_substeps, a.contact_mask, s); break;
    case kModeJacobian: e = quad::launch_qstep<T, kModeJacobian>(mp, v, act, a.n_substeps, a.contact_mask, s); break;
    case kModeOsc: e = quad::launch_qstep<T, kModeOsc>(mp, v, act, a.n_substeps, a.contact_mask, s); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return e;
}

template <typename T>
cudaError_t Launch<T>::env_step(const ModelPair<T>& mp, const BatchView<T>& v, const EnvStepArgs& a, cudaStream_t s) {
  if (!engine_for_mode(a.mode) || v.n < kQuadMinEnvs) return thread_env_step<T>(mp, v, a, s);
  const EnvStepDev<T> d = make_env_step_dev<T>(a);
  cudaError_t e;
  switch (a.mode) {
    case kModeTorque: e = quad::launch_qenv_step<T, kModeTorque>(mp, v, d, s); break;
    case kModePd: e = quad::launch_qenv_step<T, kModePd>(mp, v, d, s); break;
    case kModeOsc: e = quad::launch_qenv_step<T, kModeOsc>(mp, v, d, s); break;
    default: return cudaErrorInvalidValue;  // the Python envs have no Jacobian action space
  }
  count_launch();
  return e;
}

template <typename T>
cudaError_t Launch<T>::squat(const ModelPair<T>& mp, const BatchView<T>& v, const SquatArgs& a, cudaStream_t s) {
  if (!engine_for_mode(a.mode) || v.n < kQuadMinEnvs) return thread_squat<T>(mp, v, a, s);
  cudaError_t e;
  if (a.mode == kModeJacobian) e = quad::launch_qsquat<T, kModeJacobian>(mp, v, (const T*)a.phase, a.n_steps, a.contact_mask, s);
  else if (a.mode == kModeOsc) e = quad::launch_qsquat<T, kModeOsc>(mp, v, (const T*)a.phase, a.n_steps, a.contact_mask, s);
  else return cudaErrorInvalidValue;
  count_launch();
  return e;
}

}  // namespace cassie
