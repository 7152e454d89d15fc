// TEST-ONLY: the quad engine (four lanes per env, cassierl_b200/csrc/quad_*.cuh) compiled for the CPU.  The four
// lanes of one env run as four threads that meet at a spin barrier for every shuffle / warp sync (quad_rt.cuh), so
// the cooperative code itself -- not a serial twin -- is checked against the oracle by the `-m "not gpu"` tests.
#define CASSIE_HOST_HARNESS 1
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "../../cassierl_b200/csrc/mjcf_flatten.h"
#include "../../cassierl_b200/csrc/quad_ctrl.cuh"
#include "../../cassierl_b200/csrc/cassie_step.cuh"

using namespace cassie;
bool cassie_force_general_path = false;
bool cassie_force_tier1 = false;
bool cassie_force_tier2 = false;

static FlatModels g_models;
static std::string g_err;

template <typename T>
static void run_steps(int n, double* q, double* qd, double* warm, const double* u, int* nrows, int* sweeps, unsigned* mask) {
  PlanarModel<T> m = cast_model<T>(g_models.phys);
  std::vector<T> sbuf(quad::StateLayout::end, T(0)), buf(quad::PhysLayout::end, T(0));
  quad::SV<T> S{sbuf.data()}, W{buf.data()};
  typedef quad::StateLayout P;
  for (int i = 0; i < kNV; i++) { S[P::q + i] = (T)q[i]; S[P::qd + i] = (T)qd[i]; S[P::warm + i] = (T)warm[i]; }
  for (int s = 0; s < n; s++) {
    for (int i = 0; i < kNU; i++) S[P::u + i] = (T)u[s * kNU + i];
    quad::QStepStats st[4];
    quad::run_quad([&](int l) {
      quad::quad_physics_step(m, g_models.phys, quad::lane_id(), S, W, &st[l]);
    });
    if (nrows) nrows[s] = st[0].nrows;
    if (sweeps) sweeps[s] = st[0].sweeps;
    if (mask) mask[s] = st[0].contact_mask;
  }
  for (int i = 0; i < kNV; i++) { q[i] = S[P::q + i]; qd[i] = S[P::qd + i]; warm[i] = S[P::warm + i]; }
}

// mode: 0 torque 1 pd 2 jacobian 3 osc; the same outputs as harness.cpp ctrl_step
template <typename T>
static void ctrl_step(int mode, int n, double* q, double* qd, double* warm, const double* act, int adim,
                      double* u_out, double* op_out, double* traj_out, unsigned* mask_out, int* qp_out, unsigned* qp_set_io) {
  PlanarModel<T> mp = cast_model<T>(g_models.phys);
  typedef quad::StateLayout X;
  const size_t scratch_bytes = std::max(sizeof(T) * quad::PhysLayout::end, sizeof(double) * quad::CtrlLayout::end);
  std::vector<double> scratch(scratch_bytes / 8 + 1, 0.0);
  std::vector<T> sbuf(X::end, T(0));
  quad::SV<T> S{sbuf.data()};
  for (int i = 0; i < kNV; i++) { S[X::q + i] = (T)q[i]; S[X::qd + i] = (T)qd[i]; S[X::warm + i] = (T)warm[i]; }
  unsigned qp_set = qp_set_io ? *qp_set_io : 0u;
  for (int s = 0; s < n; s++) {
    T a[8];
    for (int i = 0; i < adim; i++) a[i] = (T)act[s * adim + i];
    quad::QStepStats st[4];
    OscStats qs[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    unsigned qps[4] = {qp_set, qp_set, qp_set, qp_set};
    quad::run_quad([&](int l) {
      const quad::Lane ln = quad::lane_id();
      if (mode == 0) quad::quad_controller_step<kModeTorque>(mp, g_models.phys, g_models.ctrl, ln, S, scratch.data(), 0, a, true, &st[l], &qs[l], &qps[l]);
      else if (mode == 1) quad::quad_controller_step<kModePd>(mp, g_models.phys, g_models.ctrl, ln, S, scratch.data(), 0, a, true, &st[l], &qs[l], &qps[l]);
      else if (mode == 2) quad::quad_controller_step<kModeJacobian>(mp, g_models.phys, g_models.ctrl, ln, S, scratch.data(), 0, a, true, &st[l], &qs[l], &qps[l]);
      else quad::quad_controller_step<kModeOsc>(mp, g_models.phys, g_models.ctrl, ln, S, scratch.data(), 0, a, true, &st[l], &qs[l], &qps[l]);
    });
    qp_set = qps[0];
    if (u_out) for (int i = 0; i < kNU; i++) u_out[s * kNU + i] = S[X::u + i];
    if (op_out) {
      T o[18];
      quad::quad_op_array(S, o);
      for (int i = 0; i < 18; i++) op_out[s * 18 + i] = o[i];
    }
    if (traj_out) for (int i = 0; i < kNV; i++) { traj_out[s * 26 + i] = S[X::q + i]; traj_out[s * 26 + 13 + i] = S[X::qd + i]; }
    if (mask_out) mask_out[s] = st[0].contact_mask;
    if (qp_out) { qp_out[2 * s] = qs[0].iters; qp_out[2 * s + 1] = qs[0].status; }
  }
  if (qp_set_io) *qp_set_io = qp_set;
  for (int i = 0; i < kNV; i++) { q[i] = S[X::q + i]; qd[i] = S[X::qd + i]; warm[i] = S[X::warm + i]; }
}

extern "C" {
void qh_ctrl_steps_f64(int mode, int n, double* q, double* qd, double* warm, const double* act, int adim,
                       double* u_out, double* op_out, double* traj_out, unsigned* mask_out, int* qp_out, unsigned* qp_set_io) {
  ctrl_step<double>(mode, n, q, qd, warm, act, adim, u_out, op_out, traj_out, mask_out, qp_out, qp_set_io);
}
void qh_ctrl_steps_f32(int mode, int n, double* q, double* qd, double* warm, const double* act, int adim,
                       double* u_out, double* op_out, double* traj_out, unsigned* mask_out, int* qp_out, unsigned* qp_set_io) {
  ctrl_step<float>(mode, n, q, qd, warm, act, adim, u_out, op_out, traj_out, mask_out, qp_out, qp_set_io);
}
int qh_load(const char* path) { return flatten_mjcf_file(path, &g_models, &g_err) ? 0 : -1; }
const char* qh_error() { return g_err.c_str(); }
void qh_force_general_path(int on) { cassie_force_general_path = on != 0; }
void qh_force_tier1(int on) { cassie_force_tier1 = on != 0; }
void qh_force_tier2(int on) { cassie_force_tier2 = on != 0; }
void qh_steps_f64(int n, double* q, double* qd, double* warm, const double* u, int* nrows, int* sweeps, unsigned* mask) {
  run_steps<double>(n, q, qd, warm, u, nrows, sweeps, mask);
}
void qh_steps_f32(int n, double* q, double* qd, double* warm, const double* u, int* nrows, int* sweeps, unsigned* mask) {
  run_steps<float>(n, q, qd, warm, u, nrows, sweeps, mask);
}
}
