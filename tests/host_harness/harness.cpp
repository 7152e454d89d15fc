// TEST-ONLY: compiles the device engine (__host__ __device__ headers under
// cassierl_b200/csrc) and the host flattener for the CPU, so that `-m "not gpu"` tests can
// check the kernel's arithmetic against the oracle without a GPU.  Never loaded by the
// product package.
#define CASSIE_HOST_HARNESS 1
#include <cstring>
#include <string>
#include "../../cassierl_b200/csrc/mjcf_flatten.h"
#include "../../cassierl_b200/csrc/cassie_step.cuh"

using namespace cassie;
bool cassie_force_general_path = false;

// ---- operation-counting scalar: the engine instantiated on it yields the exact algorithmic
// FLOP count of one step (bench.py's roofline numerator)
struct OpCount { long add = 0, mul = 0, div = 0, sqrt_ = 0, trig = 0, cmp = 0; };
static thread_local OpCount g_ops;
struct CountD {
  double v;
  CountD() : v(0) {}
  CountD(double x) : v(x) {}
  CountD(float x) : v(x) {}
  CountD(int x) : v(x) {}
  explicit operator double() const { return v; }
};
static inline CountD operator+(CountD a, CountD b) { g_ops.add++; return CountD(a.v + b.v); }
static inline CountD operator-(CountD a, CountD b) { g_ops.add++; return CountD(a.v - b.v); }
static inline CountD operator*(CountD a, CountD b) { g_ops.mul++; return CountD(a.v * b.v); }
static inline CountD operator/(CountD a, CountD b) { g_ops.div++; return CountD(a.v / b.v); }
static inline CountD operator-(CountD a) { return CountD(-a.v); }
static inline CountD& operator+=(CountD& a, CountD b) { g_ops.add++; a.v += b.v; return a; }
static inline CountD& operator-=(CountD& a, CountD b) { g_ops.add++; a.v -= b.v; return a; }
static inline CountD& operator*=(CountD& a, CountD b) { g_ops.mul++; a.v *= b.v; return a; }
static inline bool operator<(CountD a, CountD b) { g_ops.cmp++; return a.v < b.v; }
static inline bool operator>(CountD a, CountD b) { g_ops.cmp++; return a.v > b.v; }
static inline bool operator<=(CountD a, CountD b) { g_ops.cmp++; return a.v <= b.v; }
static inline bool operator>=(CountD a, CountD b) { g_ops.cmp++; return a.v >= b.v; }
static inline bool operator==(CountD a, CountD b) { g_ops.cmp++; return a.v == b.v; }
namespace cassie {
template <> struct Num<CountD> {
  static void sincos_(CountD a, CountD* s, CountD* c) { g_ops.trig++; s->v = sin(a.v); c->v = cos(a.v); }
  static CountD sqrt_(CountD a) { g_ops.sqrt_++; return CountD(sqrt(a.v)); }
  static CountD abs_(CountD a) { return CountD(fabs(a.v)); }
  static CountD pow_(CountD a, CountD b) { g_ops.trig++; return CountD(pow(a.v, b.v)); }
  static CountD exp_(CountD a) { g_ops.trig++; return CountD(exp(a.v)); }
  static CountD max_(CountD a, CountD b) { return CountD(fmax(a.v, b.v)); }
  static CountD min_(CountD a, CountD b) { return CountD(fmin(a.v, b.v)); }
  static CountD rcp_(CountD a) { return CountD(1.0) / a; }
  static constexpr bool kExactConeTest = true;
};
}  // namespace cassie

static FlatModels g_models;
static std::string g_err;

template <typename T>
static void run_steps(int n, double* q, double* qd, double* warm, const double* u, int* nrows, int* sweeps, unsigned* mask) {
  PlanarModel<T> m = cast_model<T>(g_models.phys);
  T tq[kNV], tv[kNV], tw[kNV], tu[kNU];
  for (int i = 0; i < kNV; i++) { tq[i] = (T)q[i]; tv[i] = (T)qd[i]; tw[i] = (T)warm[i]; }
  static thread_local Rows<T> rows;
  for (int s = 0; s < n; s++) {
    for (int i = 0; i < kNU; i++) tu[i] = (T)u[s * kNU + i];
    StepStats st;
    physics_step(m, g_models.phys, tq, tv, tw, tu, rows, &st);
    if (nrows) nrows[s] = st.nrows;
    if (sweeps) sweeps[s] = st.sweeps;
    if (mask) mask[s] = st.contact_mask;
  }
  for (int i = 0; i < kNV; i++) { q[i] = tq[i]; qd[i] = tv[i]; warm[i] = tw[i]; }
}

template <typename T, typename TC>
static void ctrl_step(int mode, int n, double* q, double* qd, double* warm, const double* act, int adim,
                      double* u_out, double* op_out, double* traj_out, unsigned* mask_out, int* qp_out) {
  PlanarModel<T> mp = cast_model<T>(g_models.phys);
  PlanarModel<TC> mc = cast_model<TC>(g_models.ctrl);
  unsigned qp_set = 0u;
  T tq[kNV], tv[kNV], tw[kNV];
  for (int i = 0; i < kNV; i++) { tq[i] = (T)q[i]; tv[i] = (T)qd[i]; tw[i] = (T)warm[i]; }
  static thread_local Rows<T> rows;
  for (int s = 0; s < n; s++) {
    T a[8], u[kNU];
    for (int i = 0; i < adim; i++) a[i] = (T)act[s * adim + i];
    OpState<T> op;
    StepStats st;
    OscStats qs = {0, 0};
    controller_step_dyn(mp, g_models.phys, mc, g_models.ctrl, mode, tq, tv, tw, a, rows, u, &op, &st, &qs, &qp_set);
    if (u_out) for (int i = 0; i < kNU; i++) u_out[s * kNU + i] = u[i];
    if (op_out) {
      T o[18];
      op_state_array(op, tq, tv, o);
      for (int i = 0; i < 18; i++) op_out[s * 18 + i] = o[i];
    }
    if (traj_out) for (int i = 0; i < kNV; i++) { traj_out[s * 26 + i] = tq[i]; traj_out[s * 26 + 13 + i] = tv[i]; }
    if (mask_out) mask_out[s] = st.contact_mask;
    if (qp_out) { qp_out[2 * s] = qs.iters; qp_out[2 * s + 1] = qs.status; }
  }
  for (int i = 0; i < kNV; i++) { q[i] = tq[i]; qd[i] = tv[i]; warm[i] = tw[i]; }
}

// the squatting.py loop (squatting.py:8-16): mode 2 = standing_controller_jacobian, 3 = _osc
template <typename T, typename TC>
static void squat(int mode, int n, double phase, double* q, double* qd, double* warm, double* traj_out, double* u_out, int* qp_out) {
  PlanarModel<T> mp = cast_model<T>(g_models.phys);
  PlanarModel<TC> mc = cast_model<TC>(g_models.ctrl);
  unsigned qp_set = 0u;
  T tq[kNV], tv[kNV], tw[kNV];
  for (int i = 0; i < kNV; i++) { tq[i] = (T)q[i]; tv[i] = (T)qd[i]; tw[i] = (T)warm[i]; }
  static thread_local Rows<T> rows;
  OpState<T> op;
  {
    PlanarModel<T> mct = cast_model<T>(g_models.ctrl);
    Kin<T> kc;
    forward_kinematics(mct, tq, tv, kc);
    op_state_from_kin(mct, kc, tq, op);
  }
  const double w = 0.5 * 3.1415;
  double t = 0.0;
  for (int s = 0; s < n; s++) {
    T o[18], a[8], u[kNU];
    op_state_array(op, tq, tv, o);
    const T zt = (T)(0.7 + 0.25 * sin(w * t + phase)), zdt = (T)(0.25 * cos(w * t + phase));
    if (mode == kModeJacobian) squat_jacobian_action(o, zt, zdt, a);
    else squat_osc_action(o, zt, zdt, a);
    OscStats qs = {0, 0};
    controller_step_dyn(mp, g_models.phys, mc, g_models.ctrl, mode, tq, tv, tw, a, rows, u, &op, (StepStats*)nullptr, &qs, &qp_set);
    if (qp_out) { qp_out[2 * s] = qs.iters; qp_out[2 * s + 1] = qs.status; }
    t = t + 0.0005;
    if (traj_out) for (int i = 0; i < kNV; i++) { traj_out[s * 26 + i] = tq[i]; traj_out[s * 26 + 13 + i] = tv[i]; }
    if (u_out) for (int i = 0; i < kNU; i++) u_out[s * kNU + i] = u[i];
  }
  for (int i = 0; i < kNV; i++) { q[i] = tq[i]; qd[i] = tv[i]; warm[i] = tw[i]; }
}

static_assert(sizeof(PlanarModel<CountD>) == sizeof(PlanarModel<double>), "CountD must wrap exactly one double");

extern "C" {

// exact operation counts of ONE Step* call at the given state: out[0..5] = add/sub, mul, div, sqrt,
// transcendental (sincos/pow/exp), compare ; out[6] = constraint rows, out[7] = PGS sweeps,
// out[8] = QP iterations (the QP itself runs in plain double and is counted analytically), out[9] = partition
void hh_count_ops(int mode, const double* q, const double* qd, const double* warm, const double* act, int adim, long* out) {
  PlanarModel<CountD> mp, mc, mg;
  std::memcpy((void*)&mg, &g_models.phys, sizeof(mg));
  std::memcpy((void*)&mp, &g_models.phys, sizeof(mp));
  std::memcpy((void*)&mc, &g_models.ctrl, sizeof(mc));
  CountD tq[kNV], tv[kNV], tw[kNV], a[8], u[kNU];
  for (int i = 0; i < kNV; i++) { tq[i].v = q[i]; tv[i].v = qd[i]; tw[i].v = warm[i]; }
  for (int i = 0; i < adim; i++) a[i].v = act[i];
  static thread_local Rows<CountD> rows;
  OpState<CountD> op;
  StepStats st;
  g_ops = OpCount();
  OscStats qs = {0, 0};
  unsigned qp_set = (unsigned)out[8];   // in: warm-start partition, out: QP iterations
  controller_step_dyn(mp, mg, mc, mc, mode, tq, tv, tw, a, rows, u, &op, &st, &qs, &qp_set);
  out[8] = qs.iters; out[9] = qp_set;
  out[0] = g_ops.add; out[1] = g_ops.mul; out[2] = g_ops.div; out[3] = g_ops.sqrt_; out[4] = g_ops.trig; out[5] = g_ops.cmp;
  out[6] = st.nrows; out[7] = st.sweeps;
}

void hh_force_general_path(int on) { cassie_force_general_path = on != 0; }
// the two pseudo-inverse routines of controllers.cuh (fast Cholesky / QR path + certified Jacobi fallback)
void hh_sym4_pinv(const double* A, double tol, double* P, int f32) {
  if (f32) { float a[4][4], p[4][4]; for (int i = 0; i < 16; i++) a[i / 4][i % 4] = (float)A[i]; sym4_pinv(a, (float)tol, p); for (int i = 0; i < 16; i++) P[i] = p[i / 4][i % 4]; }
  else { double a[4][4], p[4][4]; for (int i = 0; i < 16; i++) a[i / 4][i % 4] = A[i]; sym4_pinv(a, tol, p); for (int i = 0; i < 16; i++) P[i] = p[i / 4][i % 4]; }
}
void hh_pinv13x6_apply(const double* B, double tol, const double* rhs, double* u) {
  double b[kNV][kNU], r[kNV], uu[kNU];
  for (int i = 0; i < kNV; i++) { r[i] = rhs[i]; for (int j = 0; j < kNU; j++) b[i][j] = B[i * kNU + j]; }
  pinv13x6_apply(b, tol, r, uu);
  for (int j = 0; j < kNU; j++) u[j] = uu[j];
}
int hh_load(const char* path) { return flatten_mjcf_file(path, &g_models, &g_err) ? 0 : -1; }
const char* hh_error() { return g_err.c_str(); }
const void* hh_model(int ctrl) { return ctrl ? &g_models.ctrl : &g_models.phys; }
int hh_model_size() { return (int)sizeof(PlanarModel<double>); }
double hh_total_mass() { return g_models.phys.total_mass; }

// u: [n][6] torques per step
void hh_steps_f64(int n, double* q, double* qd, double* warm, const double* u, int* nrows, int* sweeps, unsigned* mask) {
  run_steps<double>(n, q, qd, warm, u, nrows, sweeps, mask);
}
void hh_steps_f32(int n, double* q, double* qd, double* warm, const double* u, int* nrows, int* sweeps, unsigned* mask) {
  run_steps<float>(n, q, qd, warm, u, nrows, sweeps, mask);
}

// dynamics pieces at (q, qd) on the physics (ctrl=0) or controller (ctrl=1) model
void hh_dynamics(int ctrl, const double* q, const double* qd, double* M /*13x13 full*/, double* bias) {
  const PlanarModel<double>& m = ctrl ? g_models.ctrl : g_models.phys;
  Kin<double> k;
  forward_kinematics(m, q, qd, k);
  double Mm[kNV][kNV];
  std::memset(Mm, 0, sizeof(Mm));
  mass_matrix(m, k, Mm);
  for (int i = 0; i < kNV; i++)
    for (int j = 0; j <= i; j++) { M[i * kNV + j] = Mm[i][j]; M[j * kNV + i] = Mm[i][j]; }
  bias_forces(m, k, bias);
}

// controller-model pieces (DynamicState.cpp:45-91): Jeq rows (4x13: Lx Lz Rx Rz), JeqdotQdot is
// folded into gamma; returns bias (C+G+Dqd), gamma, and Nc applied to the identity (13x13)
void hh_ctrl_dynamics(const double* q, const double* qd, double* bias, double* Jeq, double* gamma, double* Nc) {
  const PlanarModel<double>& m = g_models.ctrl;
  Kin<double> k;
  forward_kinematics(m, q, qd, k);
  static CtrlDyn<double> d;
  ctrl_dynamics(m, k, qd, d);
  for (int i = 0; i < kNV; i++) { bias[i] = d.bias[i]; gamma[i] = d.gamma[i]; }
  for (int r = 0; r < 4; r++) {
    double e[kNV];
    expand_row(d.Jeq[r], r / 2, e);
    for (int i = 0; i < kNV; i++) Jeq[r * kNV + i] = e[i];
  }
  for (int c = 0; c < kNV; c++) {
    double x[kNV] = {0};
    x[c] = 1;
    apply_Nc(d, x);
    for (int i = 0; i < kNV; i++) Nc[i * kNV + c] = x[i];
  }
}

// mode: 0 torque 1 pd 2 jacobian 3 osc ; act [n][adim]; u_out [n][6]; op_out [n][18] = the
// GetOperationalSpaceState a caller would read AFTER each step; traj_out [n][26] qpos,qvel
void hh_ctrl_steps_f64(int mode, int n, double* q, double* qd, double* warm, const double* act, int adim,
                       double* u_out, double* op_out, double* traj_out, unsigned* mask_out, int* qp_out) {
  ctrl_step<double, double>(mode, n, q, qd, warm, act, adim, u_out, op_out, traj_out, mask_out, qp_out);
}
void hh_ctrl_steps_f32(int mode, int n, double* q, double* qd, double* warm, const double* act, int adim,
                       double* u_out, double* op_out, double* traj_out, unsigned* mask_out, int* qp_out) {
  // the fp32 product build runs the OSC controller in double (cassie_step.cuh)
  if (mode == kModeOsc) ctrl_step<float, double>(mode, n, q, qd, warm, act, adim, u_out, op_out, traj_out, mask_out, qp_out);
  else ctrl_step<float, float>(mode, n, q, qd, warm, act, adim, u_out, op_out, traj_out, mask_out, qp_out);
}
void hh_squat_f64(int mode, int n, double phase, double* q, double* qd, double* warm, double* traj_out, double* u_out, int* qp_out) {
  squat<double, double>(mode, n, phase, q, qd, warm, traj_out, u_out, qp_out);
}
void hh_squat_f32(int mode, int n, double phase, double* q, double* qd, double* warm, double* traj_out, double* u_out, int* qp_out) {
  if (mode == kModeOsc) squat<float, double>(mode, n, phase, q, qd, warm, traj_out, u_out, qp_out);
  else squat<float, float>(mode, n, phase, q, qd, warm, traj_out, u_out, qp_out);
}
}
