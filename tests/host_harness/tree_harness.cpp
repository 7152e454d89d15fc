// TEST INFRASTRUCTURE: the 3-D tree engine (cassierl_b200/csrc/tree_engine.cuh) compiled for the CPU as a one-lane
// tile, so that tests/test_tree_host.py can check it against the oracle without a GPU.
#include <cmath>
#include <cstring>
#include <string>
#include "../../cassierl_b200/csrc/mjcf_flatten.h"
#include "../../cassierl_b200/csrc/tree_engine.cuh"

// operation-counting scalar: add/sub, mul, div, sqrt count 1 each (an FMA therefore 2); sin/cos/pow listed apart
struct Cnt {
  double v;
  Cnt() : v(0) {}
  Cnt(double x) : v(x) {}
  Cnt(int x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
};
static long g_ops[4];   // add, mul, div+sqrt, transcendental
inline Cnt operator+(Cnt a, Cnt b) { g_ops[0]++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { g_ops[0]++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { g_ops[1]++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { g_ops[2]++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt& operator+=(Cnt& a, Cnt b) { g_ops[0]++; a.v += b.v; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { g_ops[0]++; a.v -= b.v; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { g_ops[1]++; a.v *= b.v; return a; }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
inline bool operator==(Cnt a, Cnt b) { return a.v == b.v; }
inline bool operator!=(Cnt a, Cnt b) { return a.v != b.v; }
inline Cnt sqrt(Cnt a) { g_ops[2]++; return Cnt(std::sqrt(a.v)); }
inline Cnt sin(Cnt a) { g_ops[3]++; return Cnt(std::sin(a.v)); }
inline Cnt cos(Cnt a) { g_ops[3]++; return Cnt(std::cos(a.v)); }
inline Cnt pow(Cnt a, Cnt b) { g_ops[3]++; return Cnt(std::pow(a.v, b.v)); }
inline Cnt fmax(Cnt a, Cnt b) { return a.v > b.v ? a : b; }

using namespace cassie;
using namespace cassie::tree;

static TreeModel<double> g_m64;
static TreeModel<float> g_m32;
static std::string g_err;

template <typename T> static const TreeModel<T>& model();
template <> const TreeModel<double>& model<double>() { return g_m64; }
template <> const TreeModel<float>& model<float>() { return g_m32; }

template <typename T>
static void steps(int n, double* q, double* qd, double* warm, const double* u, int* stats) {
  const TreeModel<T>& m = model<T>();
  static Scratch<T> s;
  const Tile<1> tl = Tile<1>::make();
  // user (MuJoCo) order <-> internal depth order
  for (int i = 0; i < 7; i++) s.q[i] = (T)q[i];
  for (int d = 6; d < m.nv; d++) s.q[d + 1] = (T)q[m.user_dof[d] + 1];
  for (int d = 0; d < m.nv; d++) { s.qd[d] = (T)qd[m.user_dof[d]]; s.warm[d] = (T)warm[m.user_dof[d]]; }
  T uu[kMaxAct];
  for (int a = 0; a < m.nu; a++) uu[a] = (T)u[a];
  TreeStats st = {0, 0, 0, 0};
  for (int k = 0; k < n; k++) tree_step(tl, m, s, uu, &st);
  for (int i = 0; i < 7; i++) q[i] = (double)s.q[i];
  for (int d = 6; d < m.nv; d++) q[m.user_dof[d] + 1] = (double)s.q[d + 1];
  for (int d = 0; d < m.nv; d++) { qd[m.user_dof[d]] = (double)s.qd[d]; warm[m.user_dof[d]] = (double)s.warm[d]; }
  if (stats) { stats[0] = st.nefc; stats[1] = st.ncon; stats[2] = st.sweeps; stats[3] = st.dropped; }
}

extern "C" {
const char* th_error() { return g_err.c_str(); }
int th_load(const char* path) {
  if (!flatten_tree_file(path, &g_m64, &g_err)) return -1;
  cast_tree_model(&g_m32, g_m64);
  return 0;
}
void th_sizes(int* out) {
  out[0] = g_m64.nl; out[1] = g_m64.nv; out[2] = g_m64.nq; out[3] = g_m64.nu; out[4] = g_m64.ng; out[5] = g_m64.npair;
  out[6] = g_m64.neq; out[7] = g_m64.nlevels; out[8] = (int)sizeof(Scratch<float>); out[9] = (int)sizeof(Scratch<double>);
}
void th_consts(double* dof_invweight, double* meaninertia, double* pair_invweight, double* eq_invweight, double* qpos0,
               double* link_mass) {
  for (int d = 0; d < g_m64.nv; d++) dof_invweight[g_m64.user_dof[d]] = g_m64.dof_invweight[d];
  *meaninertia = g_m64.meaninertia;
  for (int p = 0; p < g_m64.npair; p++) pair_invweight[p] = g_m64.pair_invweight[p];
  for (int e = 0; e < g_m64.neq; e++) eq_invweight[e] = g_m64.eq_invweight[e];
  for (int i = 0; i < 7; i++) qpos0[i] = g_m64.qpos0[i];
  for (int d = 6; d < g_m64.nv; d++) qpos0[g_m64.user_dof[d] + 1] = g_m64.qpos0[d + 1];
  for (int l = 0; l < g_m64.nl; l++) link_mass[l] = g_m64.mass[l];
}
// mass matrix [nv x nv] and bias [nv] at (q, qd)
void th_dynamics(const double* q, const double* qd, double* M, double* bias) {
  static Scratch<double> s;
  const Tile<1> tl = Tile<1>::make();
  const int* ud = g_m64.user_dof;
  for (int i = 0; i < 7; i++) s.q[i] = q[i];
  for (int d = 6; d < g_m64.nv; d++) s.q[d + 1] = q[ud[d] + 1];
  for (int d = 0; d < g_m64.nv; d++) s.qd[d] = qd[ud[d]];
  dynamics(tl, g_m64, s);
  for (int i = 0; i < g_m64.nv; i++) {
    bias[ud[i]] = s.tmp[i];
    for (int j = 0; j < g_m64.nv; j++) M[ud[i] * g_m64.nv + ud[j]] = i == j ? s.Mdiag[i] : (i < j ? s.L[i * kLD + j] : s.L[j * kLD + i]);
  }
}
// constraint rows at (q, qd) in user dof order: returns nefc; J [nefc x nv], pos, R, aref, type, id
int th_rows(const double* q, const double* qd, double* J, double* pos, double* R, double* aref, int* type, int* id) {
  static Scratch<double> s;
  const Tile<1> tl = Tile<1>::make();
  const int* ud = g_m64.user_dof;
  const int nv = g_m64.nv;
  for (int i = 0; i < 7; i++) s.q[i] = q[i];
  for (int d = 6; d < nv; d++) s.q[d + 1] = q[ud[d] + 1];
  for (int d = 0; d < nv; d++) s.qd[d] = qd[ud[d]];
  dynamics(tl, g_m64, s);
  collide(tl, g_m64, s);
  make_rows(tl, g_m64, s);
  for (int r = 0; r < s.nefc; r++) {
    for (int d = 0; d < nv; d++) J[r * nv + ud[d]] = s.J[r * kLD + d];
    pos[r] = s.r_pos[r]; R[r] = s.r_R[r]; aref[r] = s.r_aref[r]; type[r] = s.r_type[r];
    id[r] = s.r_type[r] == kRowLimit ? ud[s.r_id[r]] : s.r_id[r];
  }
  return s.nefc;
}
// one step from (q, qd, warm) with controls u, returning the constraint forces and qacc (user order) for debugging
int th_debug_step(const double* q, const double* qd, const double* warm, const double* u, double* f, double* qacc, double* b, double* f0) {
  static Scratch<double> s;
  const Tile<1> tl = Tile<1>::make();
  const int* ud = g_m64.user_dof;
  const int nv = g_m64.nv;
  for (int i = 0; i < 7; i++) s.q[i] = q[i];
  for (int d = 6; d < nv; d++) s.q[d + 1] = q[ud[d] + 1];
  for (int d = 0; d < nv; d++) { s.qd[d] = qd[ud[d]]; s.warm[d] = warm[ud[d]]; }
  TreeStats st = {0, 0, 0, 0};
  tree_step(tl, g_m64, s, u, &st);
  for (int r = 0; r < s.nefc; r++) { f[r] = s.r_f[r]; b[r] = s.r_b[r]; }
  for (int d = 0; d < nv; d++) qacc[ud[d]] = s.qacc[d];
  (void)f0;
  return s.sweeps;
}
// algorithmic operation count of ONE step from (q, qd, warm) with controls u: out = add, mul, div+sqrt, transcendental,
// constraint rows, PGS sweeps
void th_count_ops(const double* q, const double* qd, const double* warm, const double* u, long* out) {
  static TreeModel<Cnt> mc;
  static Scratch<Cnt> s;
  cast_tree_model(&mc, g_m64);
  const Tile<1> tl = Tile<1>::make();
  const int* ud = g_m64.user_dof;
  const int nv = g_m64.nv;
  for (int i = 0; i < 7; i++) s.q[i] = Cnt(q[i]);
  for (int d = 6; d < nv; d++) s.q[d + 1] = Cnt(q[ud[d] + 1]);
  for (int d = 0; d < nv; d++) { s.qd[d] = Cnt(qd[ud[d]]); s.warm[d] = Cnt(warm[ud[d]]); }
  Cnt uu[kMaxAct];
  for (int a = 0; a < g_m64.nu; a++) uu[a] = Cnt(u[a]);
  TreeStats st = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) g_ops[i] = 0;
  tree_step(tl, mc, s, uu, &st);
  for (int i = 0; i < 4; i++) out[i] = g_ops[i];
  out[4] = st.nefc; out[5] = st.sweeps;
}
// one step with the FAST capacity and abort_on_overflow (the kernels' first pass): returns 1 when the step was taken, 0 when
// it needs more rows / contacts (the state must then be untouched); stats as th_steps
int th_step_fast_f64(double* q, double* qd, double* warm, const double* u, int* stats) {
  const TreeModel<double>& m = g_m64;
  static Scratch<double, kFastRows, kFastCon> s;
  const Tile<1> tl = Tile<1>::make();
  for (int i = 0; i < 7; i++) s.q[i] = q[i];
  for (int d = 6; d < m.nv; d++) s.q[d + 1] = q[m.user_dof[d] + 1];
  for (int d = 0; d < m.nv; d++) { s.qd[d] = qd[m.user_dof[d]]; s.warm[d] = warm[m.user_dof[d]]; }
  TreeStats st = {0, 0, 0, 0};
  s.n_dropped = 0; s.overflow = 0;
  const bool ok = tree_step(tl, m, s, u, &st, true);
  for (int i = 0; i < 7; i++) q[i] = s.q[i];
  for (int d = 6; d < m.nv; d++) q[m.user_dof[d] + 1] = s.q[d + 1];
  for (int d = 0; d < m.nv; d++) { qd[m.user_dof[d]] = s.qd[d]; warm[m.user_dof[d]] = s.warm[d]; }
  if (stats) { stats[0] = s.nefc; stats[1] = s.ncon; stats[2] = st.sweeps; stats[3] = s.n_dropped; }
  return ok ? 1 : 0;
}
void th_steps_f64(int n, double* q, double* qd, double* warm, const double* u, int* stats) { steps<double>(n, q, qd, warm, u, stats); }
void th_steps_f32(int n, double* q, double* qd, double* warm, const double* u, int* stats) { steps<float>(n, q, qd, warm, u, stats); }
}
