// placeholder until the device QP lands
#pragma once
#include "controllers.cuh"
namespace cassie {
struct OscStats { int iters; int status; };
template <typename T>
CASSIE_HD void osc_control(const PlanarModel<T>& m, const Kin<T>& k, const T* qd, const T a[7], T u[kNU], OscStats* st) {
  CASSIE_UNROLL
  for (int i = 0; i < kNU; i++) u[i] = T(0);
  if (st) { st->iters = 0; st->status = -1; }
}
}  // namespace cassie
