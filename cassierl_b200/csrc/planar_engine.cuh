// Planar (x-z) rigid-body engine for the Cassie-2D model class: one env per thread, all
// per-env vectors in registers, constraint rows in thread-local memory.
//
// This is a from-scratch B200 restatement of what the reference obtains from MuJoCo's mj_step
// (CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:92,115,174,206; pipeline in SURVEY.md App. B)
// specialised to models whose motion is confined to the x-z plane (cassie2d_stiff.xml):
//   * every link is (mass, planar com, inertia about world-y); the 3-D model's y-offsets and
//     off-diagonal inertias do not enter planar dynamics (DESIGN.md §3 shows why),
//   * a `connect` keeps its x and z rows (the y row has a zero Jacobian),
//   * an elliptic condim-3 floor contact keeps its normal row and one tangent row (the other
//     tangent has a zero Jacobian and its force stays identically 0 in mj_solPGS).
// Functions are __host__ __device__ so that tests/host_harness can unit-test the very same
// code on the CPU; the shipped library only ever launches them on the GPU.
#pragma once
#include "planar_model.h"
#include <math.h>

#if defined(__CUDACC__)
#define CASSIE_HD __host__ __device__ __forceinline__
#define CASSIE_COLD __host__ __device__ __noinline__
#define CASSIE_UNROLL _Pragma("unroll")
#define CASSIE_ROLL _Pragma("unroll 1")   // keep a loop rolled: the straight-line code size is what limits the step kernel
#else
#define CASSIE_HD inline
#define CASSIE_COLD inline
#define CASSIE_UNROLL
#define CASSIE_ROLL
#endif

#ifdef CASSIE_HOST_HARNESS
extern bool cassie_force_general_path;
#endif

namespace cassie {

// sin and cos of a double argument of moderate size (joint angles, |a| < ~1e3): two-term Cody-Waite
// reduction by pi/2 and the fdlibm __kernel_sin / __kernel_cos minimax polynomials on [-pi/4, pi/4]
// (absolute error < 2e-16).  Used instead of the CUDA / libm sincos so that the host harness and
// the device agree bit for bit and the kernel does not inline 20-odd copies of the slow-path range
// reduction (13.7k of 44k SASS instructions before).
CASSIE_HD void sincos_reduced(double a, double* sn, double* cs) {
  const double k = rint(a * 6.36619772367581382433e-01);
  double r = a - k * 1.57079632673412561417e+00;
  r = r - k * 6.07710050650619224932e-11;
  const double z = r * r;
  const double ps = -1.66666666666666324348e-01 + z * (8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 +
                    z * (2.75573137070700676789e-06 + z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
  const double pc = 4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05 +
                    z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
  const double s = r + r * z * ps;
  const double c = 1.0 - 0.5 * z + z * z * pc;
  const int q = ((int)k) & 3;
  *sn = (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
  *cs = (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
}

template <typename T> struct Num;
template <> struct Num<float> {
  static CASSIE_HD void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
  static CASSIE_HD float sqrt_(float a) { return sqrtf(a); }
  static CASSIE_HD float abs_(float a) { return fabsf(a); }
  static CASSIE_HD float pow_(float a, float b) { return powf(a, b); }
  static CASSIE_HD float exp_(float a) { return expf(a); }
  static CASSIE_HD float max_(float a, float b) { return fmaxf(a, b); }   // one FMNMX, no predicate round trip
  static CASSIE_HD float min_(float a, float b) { return fminf(a, b); }
  // reciprocal of a value known to be in the normal range: one MUFU.RCP (a plain 1.0f / a carries ~8 instructions of
  // range scaling around it even with -prec-div=false)
  static CASSIE_HD float rcp_(float a) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
  }
  static constexpr bool kExactConeTest = false;
};
template <> struct Num<double> {
  static CASSIE_HD void sincos_(double a, double* s, double* c) { sincos_reduced(a, s, c); }
  static CASSIE_HD double sqrt_(double a) { return sqrt(a); }
  static CASSIE_HD double abs_(double a) { return fabs(a); }
  static CASSIE_HD double pow_(double a, double b) { return pow(a, b); }
  static CASSIE_HD double exp_(double a) { return exp(a); }
  static CASSIE_HD double max_(double a, double b) { return fmax(a, b); }
  static CASSIE_HD double min_(double a, double b) { return fmin(a, b); }
  static CASSIE_HD double rcp_(double a) { return 1.0 / a; }
  static constexpr bool kExactConeTest = true;
};

// j is an ancestor-or-self of dof i in the kinematic tree (j <= i)
CASSIE_HD constexpr bool dof_anc(int i, int j) {
  if (j > i) return false;
  if (j <= 2) return true;
  if (i < 3) return false;
  if ((i - 3) / 5 != (j - 3) / 5) return false;
  const int a = (i - 3) % 5, b = (j - 3) % 5;
  if (a == b) return true;
  if (b == kThigh) return true;
  if (a == kRod || b == kRod) return false;
  return b < a;
}
// leg-local: link b is an ancestor-or-self of link a
CASSIE_HD constexpr bool link_anc(int a, int b) {
  if (a == b) return true;
  if (b == kThigh) return true;
  if (a == kRod || b == kRod) return false;
  return b < a;
}
CASSIE_HD constexpr int link_parent(int a) { return a == kThigh ? -1 : (a == kRod ? kThigh : a - 1); }

// ---------------------------------------------------------------------------------------
// kinematic state; positions are RELATIVE TO THE PELVIS PIVOT (keeps fp32 well conditioned
// however far the robot has walked), velocities are absolute.
template <typename T>
struct Kin {
  T c0, s0, w0, v0x, v0z;  // pelvis: cos/sin(pitch), pitch rate, pivot velocity
  T c[2][kLegLinks], s[2][kLegLinks];
  T px[2][kLegLinks], pz[2][kLegLinks];
  T dx[2][kLegLinks], dz[2][kLegLinks];  // pivot - parent's pivot (world axes): exact local lever arms
  T w[2][kLegLinks], vx[2][kLegLinks], vz[2][kLegLinks];
};

template <typename T>
CASSIE_HD void rot(T c, T s, T x, T z, T& ox, T& oz) {
  ox = x * c + z * s;
  oz = z * c - x * s;
}

// mj_kinematics + mj_comVel [EXT] / RBDL UpdateKinematics (DynamicModel.cpp:237-242), planar.
// Split in a position pass (angles -> cos/sin, pivots) and a velocity pass so that the position
// pass can run in a wider type than the rest of the step (physics_step).
template <typename T>
CASSIE_HD void fk_positions(const PlanarModel<T>& m, const T* q, Kin<T>& k) {
  Num<T>::sincos_(q[2] - m.pel_ref[2], &k.s0, &k.c0);
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T alpha[kLegLinks];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int p = link_parent(a);
      const int dof = 3 + 5 * L + a;
      T ap, cp, sp, ppx, ppz;
      if (p < 0) { ap = q[2] - m.pel_ref[2]; cp = k.c0; sp = k.s0; ppx = T(0); ppz = T(0); }
      else { ap = alpha[p]; cp = k.c[L][p]; sp = k.s[L][p]; ppx = k.px[L][p]; ppz = k.pz[L][p]; }
      alpha[a] = ap + m.sgn[L][a] * q[dof] + m.ang0[L][a];
      Num<T>::sincos_(alpha[a], &k.s[L][a], &k.c[L][a]);
      T dx, dz;
      rot(cp, sp, m.off[L][a][0], m.off[L][a][1], dx, dz);
      k.px[L][a] = ppx + dx;
      k.pz[L][a] = ppz + dz;
      k.dx[L][a] = dx;
      k.dz[L][a] = dz;
    }
  }
}
template <typename T>
CASSIE_HD void fk_velocities(const PlanarModel<T>& m, const T* qd, Kin<T>& k) {
  k.w0 = qd[2];
  k.v0x = qd[0];
  k.v0z = qd[1];
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int p = link_parent(a);
      const int dof = 3 + 5 * L + a;
      T wp, vpx, vpz;
      if (p < 0) { wp = k.w0; vpx = k.v0x; vpz = k.v0z; }
      else { wp = k.w[L][p]; vpx = k.vx[L][p]; vpz = k.vz[L][p]; }
      k.w[L][a] = wp + m.sgn[L][a] * qd[dof];
      k.vx[L][a] = vpx + wp * k.dz[L][a];   // w y^ x d = w (d.z, -d.x)
      k.vz[L][a] = vpz - wp * k.dx[L][a];
    }
  }
}
template <typename T>
CASSIE_HD void forward_kinematics(const PlanarModel<T>& m, const T* q, const T* qd, Kin<T>& k) {
  fk_positions(m, q, k);
  fk_velocities(m, qd, k);
}
template <typename T, typename TG>
CASSIE_HD void cast_kin_positions(const Kin<TG>& g, Kin<T>& k) {
  k.c0 = (T)g.c0; k.s0 = (T)g.s0;
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      k.c[L][a] = (T)g.c[L][a]; k.s[L][a] = (T)g.s[L][a];
      k.px[L][a] = (T)g.px[L][a]; k.pz[L][a] = (T)g.pz[L][a];
      k.dx[L][a] = (T)g.dx[L][a]; k.dz[L][a] = (T)g.dz[L][a];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Mass matrix (lower triangle, structural non-zeros only) + armature.
// mj_crb [EXT] / RBDL CompositeRigidBodyAlgorithm + rotor inertia (DynamicModel.cpp:267-272).
// Composite (m, h, I) of each subtree is kept ABOUT THAT LINK'S OWN PIVOT and moved to the parent's
// pivot by the local offset d = p_child - p_parent (parallel-axis), so no large cancelling terms
// appear in fp32.  For hinge i (pivot p_i) and an ancestor hinge j (pivot p_j):
//   M_ij = s_i s_j [ I_i + (p_i - p_j) . h_i ],   M_i,x = s_i h_i,z,   M_i,z = -s_i h_i,x.
template <typename T>
CASSIE_HD void mass_matrix(const PlanarModel<T>& m, const Kin<T>& k, T M[kNV][kNV]) {
  T tm = m.pel_mass, thx, thz, tI;
  {
    T cx, cz;
    rot(k.c0, k.s0, m.pel_com[0], m.pel_com[1], cx, cz);
    thx = m.pel_mass * cx; thz = m.pel_mass * cz;
    tI = m.pel_inertia + m.pel_mass * (cx * cx + cz * cz);
  }
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T cm[kLegLinks], hx[kLegLinks], hz[kLegLinks], cI[kLegLinks];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      T rx, rz;
      rot(k.c[L][a], k.s[L][a], m.com[L][a][0], m.com[L][a][1], rx, rz);
      cm[a] = m.mass[L][a];
      hx[a] = cm[a] * rx; hz[a] = cm[a] * rz;
      cI[a] = m.inertia[L][a] + cm[a] * (rx * rx + rz * rz);
    }
    // leaves to root: toe->tarsus->knee->thigh, rod->thigh
    CASSIE_UNROLL
    for (int step = 0; step < 4; step++) {
      const int c = step == 0 ? kToe : (step == 1 ? kTarsus : (step == 2 ? kKnee : kRod));
      const int p = link_parent(c);
      const T dx = k.dx[L][c], dz = k.dz[L][c];
      cI[p] += cI[c] + T(2) * (dx * hx[c] + dz * hz[c]) + cm[c] * (dx * dx + dz * dz);
      hx[p] += hx[c] + cm[c] * dx;
      hz[p] += hz[c] + cm[c] * dz;
      cm[p] += cm[c];
    }
    {  // thigh subtree -> whole-robot composite about the pelvis pivot
      const T dx = k.dx[L][kThigh], dz = k.dz[L][kThigh];
      tI += cI[kThigh] + T(2) * (dx * hx[kThigh] + dz * hz[kThigh]) + cm[kThigh] * (dx * dx + dz * dz);
      thx += hx[kThigh] + cm[kThigh] * dx;
      thz += hz[kThigh] + cm[kThigh] * dz;
      tm += cm[kThigh];
    }
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int i = 3 + 5 * L + a;
      const T sa = m.sgn[L][a];
      M[i][0] = sa * hz[a];
      M[i][1] = -sa * hx[a];
      M[i][2] = sa * (cI[a] + k.px[L][a] * hx[a] + k.pz[L][a] * hz[a]);
      M[i][i] = cI[a] + m.armature[i];
      // walk up the chain accumulating p_a - p_b
      T lx = T(0), lz = T(0);
      int b = a;
      CASSIE_UNROLL
      for (int hop = 0; hop < 3; hop++) {
        if (link_parent(b) >= 0) {
          lx += k.dx[L][b]; lz += k.dz[L][b];
          b = link_parent(b);
          M[i][3 + 5 * L + b] = sa * m.sgn[L][b] * (cI[a] + lx * hx[a] + lz * hz[a]);
        }
      }
    }
  }
  M[0][0] = tm + m.armature[0];
  M[1][0] = T(0);
  M[1][1] = tm + m.armature[1];
  M[2][0] = thz;
  M[2][1] = -thx;
  M[2][2] = tI + m.armature[2];
}

// RNE with qdd = 0: bias = C(q,qd) qd + G(q)  (mj_rne [EXT]; RBDL NonlinearEffects,
// DynamicModel.cpp:320-323).  Planar: angular accelerations vanish, pivots carry the
// centripetal terms.  Moments are accumulated about each link's OWN pivot and moved to the parent's
// pivot by the local offset (no large cancelling terms in fp32).
template <typename T>
CASSIE_HD void bias_forces(const PlanarModel<T>& m, const Kin<T>& k, T bias[kNV]) {
  const T g = -m.gravity_z;  // f = m (a_com + g z^)
  T Fx, Fz, N;
  {
    T rx, rz;
    rot(k.c0, k.s0, m.pel_com[0], m.pel_com[1], rx, rz);
    const T w2 = k.w0 * k.w0;
    Fx = m.pel_mass * (-w2 * rx);
    Fz = m.pel_mass * (-w2 * rz + g);
    N = rz * Fx - rx * Fz;
  }
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T ax[kLegLinks], az[kLegLinks];          // pivot accelerations
    T fx[kLegLinks], fz[kLegLinks], n[kLegLinks];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int p = link_parent(a);
      T apx, apz, wp;
      if (p < 0) { apx = T(0); apz = T(0); wp = k.w0; }
      else { apx = ax[p]; apz = az[p]; wp = k.w[L][p]; }
      const T w2p = wp * wp;
      ax[a] = apx - w2p * k.dx[L][a];
      az[a] = apz - w2p * k.dz[L][a];
      T rx, rz;
      rot(k.c[L][a], k.s[L][a], m.com[L][a][0], m.com[L][a][1], rx, rz);
      const T w2 = k.w[L][a] * k.w[L][a];
      fx[a] = m.mass[L][a] * (ax[a] - w2 * rx);
      fz[a] = m.mass[L][a] * (az[a] - w2 * rz + g);
      n[a] = rz * fx[a] - rx * fz[a];
    }
    CASSIE_UNROLL
    for (int step = 0; step < 4; step++) {
      const int c = step == 0 ? kToe : (step == 1 ? kTarsus : (step == 2 ? kKnee : kRod));
      const int p = link_parent(c);
      n[p] += n[c] + (k.dz[L][c] * fx[c] - k.dx[L][c] * fz[c]);
      fx[p] += fx[c];
      fz[p] += fz[c];
    }
    N += n[kThigh] + (k.dz[L][kThigh] * fx[kThigh] - k.dx[L][kThigh] * fz[kThigh]);
    Fx += fx[kThigh];
    Fz += fz[kThigh];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) bias[3 + 5 * L + a] = m.sgn[L][a] * n[a];
  }
  bias[0] = Fx;
  bias[1] = Fz;
  bias[2] = N;
}

// ---------------------------------------------------------------------------------------
// M = L^T D L in place (unit L stored below the diagonal, 1/D returned), tree-sparse
// backward elimination = mj_factorM [EXT]; no fill-in outside the ancestor pattern.
template <typename T>
CASSIE_HD void factor(T M[kNV][kNV], T Dinv[kNV]) {
  CASSIE_UNROLL
  for (int k = kNV - 1; k >= 0; k--) {
    const T inv = Num<T>::rcp_(M[k][k]);
    Dinv[k] = inv;
    CASSIE_UNROLL
    for (int i = kNV - 1; i >= 0; i--) {
      if (i < k && dof_anc(k, i)) {
        const T l = M[k][i] * inv;
        CASSIE_UNROLL
        for (int j = 0; j < kNV; j++)
          if (j <= i && dof_anc(k, j)) M[i][j] -= l * M[k][j];
        M[k][i] = l;
      }
    }
  }
}
// x <- M^-1 x   (mj_solveLD [EXT])
template <typename T>
CASSIE_HD void solve(const T M[kNV][kNV], const T Dinv[kNV], T x[kNV]) {
  CASSIE_UNROLL
  for (int k = kNV - 1; k >= 0; k--) {
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++)
      if (i < k && dof_anc(k, i)) x[i] -= M[k][i] * x[k];
  }
  CASSIE_UNROLL
  for (int k = 0; k < kNV; k++) x[k] *= Dinv[k];
  CASSIE_UNROLL
  for (int k = 0; k < kNV; k++) {
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++)
      if (i < k && dof_anc(k, i)) x[k] -= M[k][i] * x[i];
  }
}

// ---------------------------------------------------------------------------------------
// constraint rows (thread-local memory).  A row's Jacobian is non-zero only on the three base
// dofs and the five dofs of ONE leg: J8 = [x, z, pitch, thigh, knee, tarsus, toe, rod].
enum { kRowEq = 0, kRowLimit = 1, kRowNormal = 2, kRowTangent = 3 };

template <typename T>
struct Rows {
  int n;
  T J[kMaxRows][8];
  signed char leg[kMaxRows], type[kMaxRows];
  T R[kMaxRows], b[kMaxRows], f[kMaxRows];
  T A[kMaxRows][kMaxRows];
};

constexpr double kMinVal = 1e-15;

// general solimp power: out of line (powf/pow inline ~800 SASS instructions per call site, 16 k in total
// before; every MJCF in scope uses the default power 2)
template <typename T>
CASSIE_COLD T impedance_pow(const T* si, T x) {
  if (x <= si[3]) return Num<T>::pow_(x, si[4]) / Num<T>::pow_(si[3], si[4] - T(1));
  return T(1) - Num<T>::pow_(T(1) - x, si[4]) / Num<T>::pow_(T(1) - si[3], si[4] - T(1));
}
template <typename T>
CASSIE_HD T impedance(const T* si, T pos) {  // getimpedance [EXT], margin = 0
  if (si[0] == si[1] || si[2] <= T(kMinVal)) return T(0.5) * (si[0] + si[1]);
  T x = Num<T>::abs_(pos * Num<T>::rcp_(si[2]));
  if (x >= T(1)) return si[1];
  if (x <= T(0)) return si[0];
  T y;
  if (si[4] == T(1)) y = x;
  else if (si[4] == T(2)) y = x <= si[3] ? x * x * Num<T>::rcp_(si[3]) : T(1) - (T(1) - x) * (T(1) - x) * Num<T>::rcp_(T(1) - si[3]);
  else y = impedance_pow(si, x);
  return si[0] + y * (si[1] - si[0]);
}

// point Jacobian (x row, z row) of a point fixed to leg link `a` of leg L, given by its lever
// r = P - p_a from that link's pivot (world axes);  a = -1: pelvis, r from the pelvis pivot.
// mj_jac [EXT] / RBDL CalcPointJacobian, planar.  Lever arms to the ancestor pivots are built by
// adding the local pivot-to-pivot offsets, never by subtracting two far-away positions.
template <typename T>
CASSIE_HD void point_jac(const PlanarModel<T>& m, const Kin<T>& k, int L, int a, T rx, T rz, T Jx[8], T Jz[8]) {
  CASSIE_UNROLL
  for (int b = 0; b < kLegLinks; b++) { Jx[3 + b] = T(0); Jz[3 + b] = T(0); }
  T lx = rx, lz = rz;
  int b = a;
  CASSIE_UNROLL
  for (int hop = 0; hop < 4; hop++) {
    if (b >= 0) {
      CASSIE_UNROLL
      for (int c = 0; c < kLegLinks; c++) {
        if (c == b) { Jx[3 + c] = m.sgn[L][c] * lz; Jz[3 + c] = -m.sgn[L][c] * lx; }
      }
      lx += k.dx[L][b]; lz += k.dz[L][b];
      b = link_parent(b);
    }
  }
  Jx[0] = T(1); Jx[1] = T(0); Jx[2] = lz;
  Jz[0] = T(0); Jz[1] = T(1); Jz[2] = -lx;
}

template <typename T>
CASSIE_HD T dot8_dense(const T J8[8], int leg, const T x[kNV]) {
  T s = J8[0] * x[0] + J8[1] * x[1] + J8[2] * x[2];
  CASSIE_UNROLL
  for (int b = 0; b < kLegLinks; b++) s += J8[3 + b] * (leg ? x[8 + b] : x[3 + b]);
  return s;
}

// finishes a row: R, aref, stores J (mj_makeImpedance [EXT])
template <typename T>
CASSIE_HD void push_row(Rows<T>& r, const T J8[8], int leg, int type, T diag, T imp, T K, T Bd,
                        T pos, const T qd[kNV]) {
  const int i = r.n++;
  CASSIE_UNROLL
  for (int c = 0; c < 8; c++) r.J[i][c] = J8[c];
  r.leg[i] = (signed char)leg;
  r.type[i] = (signed char)type;
  const T Rv = (T(1) - imp) * diag * Num<T>::rcp_(imp);   // imp in [1e-4, 1 - 1e-4]
  r.R[i] = Rv > T(kMinVal) ? Rv : T(kMinVal);
  const T vel = dot8_dense(J8, leg, qd);
  // aref, kept in b until qacc_smooth is known:  b = J qacc_smooth - aref
  r.b[i] = -(-Bd * vel - K * imp * pos);
}

template <typename T>
CASSIE_HD void kb_from_solref(const PlanarModel<T>& m, const T* solref, const T* solimp, T& K, T& Bd) {
  T tc = solref[0];
  if (tc > T(0) && tc < T(2) * m.timestep) tc = T(2) * m.timestep;  // refsafe
  const T dmax = solimp[1];
  T kd = dmax * dmax * tc * tc * solref[1] * solref[1];
  K = Num<T>::rcp_(kd > T(kMinVal) ? kd : T(kMinVal));
  T bd = dmax * tc;
  Bd = T(2) * Num<T>::rcp_(bd > T(kMinVal) ? bd : T(kMinVal));
}

// Position-level constraint violations: the loop-closure anchor mismatch and the signed distances of
// the collision geoms to the floor.  They are ~1e-4 m differences of ~1 m quantities and get
// multiplied by the constraint stiffness (1e4 .. 4e4 1/s^2), so the fp32 build evaluates them from
// the position pass in double (physics_step) -- the dominant single-step error otherwise.
template <typename T>
struct ConPos {
  T eq_rx[2], eq_rz[2];
  T sph_dist, cap_dist[kNumCaps][2];
};
template <typename TG, typename T>
CASSIE_HD void constraint_positions(const PlanarModel<TG>& m, const Kin<TG>& k, TG body_z, ConPos<T>& cp) {
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    TG ax, az, bx, bz;
    rot(k.c[L][kRod], k.s[L][kRod], m.eq_a1[L][0], m.eq_a1[L][1], ax, az);
    rot(k.c[L][kTarsus], k.s[L][kTarsus], m.eq_a2[L][0], m.eq_a2[L][1], bx, bz);
    // (p_rod + a) - (p_tarsus + b),  p_rod - p_tarsus = d_rod - d_knee - d_tarsus
    cp.eq_rx[L] = (T)((k.dx[L][kRod] - k.dx[L][kKnee] - k.dx[L][kTarsus]) + (ax - bx));
    cp.eq_rz[L] = (T)((k.dz[L][kRod] - k.dz[L][kKnee] - k.dz[L][kTarsus]) + (az - bz));
  }
  const TG height = body_z - m.pel_ref[1] + m.pel_org[1];  // world z of the pelvis pivot
  {
    TG cx, cz;
    rot(k.c0, k.s0, m.sph_c[0], m.sph_c[1], cx, cz);
    cp.sph_dist = (T)(height + cz - m.sph_r);
  }
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int g = 0; g < 4; g++) {
      const int cap = 4 * L + g;
      CASSIE_UNROLL
      for (int e = 0; e < 2; e++) {
        TG ex, ez;
        const TG* ep = e ? m.cap_from[cap] : m.cap_to[cap];
        rot(k.c[L][g], k.s[L][g], ep[0], ep[1], ex, ez);
        cp.cap_dist[cap][e] = (T)(height + k.pz[L][g] + ez - m.cap_r[cap]);
      }
    }
  }
}

// mj_collision + mj_makeConstraint + mj_makeImpedance [EXT], canonical row order:
// connects (L, R), joint limits (dof order), contacts (pelvis sphere, then per capsule the 'to'
// end before the 'from' end).  Returns the public contact bit mask.
// ALL = false builds only the rows of the common regime (connects + the four toe-capsule end spheres);
// the caller guarantees that nothing else is active (rare_rows_active), so both variants emit the
// same rows in the same order there.
template <bool ALL, typename T>
CASSIE_HD unsigned int make_rows(const PlanarModel<T>& m, const Kin<T>& k, const ConPos<T>& cp, const T* q, const T* qd,
                                  Rows<T>& r, int* nlimit) {
  r.n = 0;
  int nlim = 0;
  T K, Bd;
  // ---- connects
  kb_from_solref(m, m.eq_solref, m.eq_solimp, K, Bd);
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T ax, az, bx, bz;
    rot(k.c[L][kRod], k.s[L][kRod], m.eq_a1[L][0], m.eq_a1[L][1], ax, az);
    rot(k.c[L][kTarsus], k.s[L][kTarsus], m.eq_a2[L][0], m.eq_a2[L][1], bx, bz);
    T J1x[8], J1z[8], J2x[8], J2z[8];
    point_jac(m, k, L, kRod, ax, az, J1x, J1z);
    point_jac(m, k, L, kTarsus, bx, bz, J2x, J2z);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) { J1x[c] -= J2x[c]; J1z[c] -= J2z[c]; }
    const T rx = cp.eq_rx[L], rz = cp.eq_rz[L];
    const T imp = impedance(m.eq_solimp, Num<T>::sqrt_(rx * rx + rz * rz));
    push_row(r, J1x, L, kRowEq, m.eq_diag[L], imp, K, Bd, rx, qd);
    push_row(r, J1z, L, kRowEq, m.eq_diag[L], imp, K, Bd, rz, qd);
  }
  // ---- joint limits
  kb_from_solref(m, m.lim_solref, m.lim_solimp, K, Bd);
  CASSIE_UNROLL
  for (int j = 3; j < kNV; j++) {
    if (ALL && m.has_limit[j]) {
      const T dlo = q[j] - m.lim_lo[j], dhi = m.lim_hi[j] - q[j];
      CASSIE_UNROLL
      for (int side = 0; side < 2; side++) {
        const T dist = side ? dhi : dlo;
        if (dist < T(0)) {
          T J8[8];
          CASSIE_UNROLL
          for (int c = 0; c < 8; c++) J8[c] = T(0);
          J8[3 + (j - 3) % 5] = side ? T(-1) : T(1);
          push_row(r, J8, (j - 3) / 5, kRowLimit, m.lim_diag[j], impedance(m.lim_solimp, dist), K, Bd, dist, qd);
          nlim++;
        }
      }
    }
  }
  // ---- floor contacts (plane z = 0, normal +z).  Elliptic friction rows: K = 0, pos = 0.
  unsigned int mask = 0;
  kb_from_solref(m, m.con_solref, m.con_solimp, K, Bd);
  {
    T cx, cz;
    rot(k.c0, k.s0, m.sph_c[0], m.sph_c[1], cx, cz);
    const T dist = cp.sph_dist;
    if (ALL && !(dist > T(0))) {
      T Jx[8], Jz[8];
      point_jac(m, k, 0, -1, cx, cz - m.sph_r - T(0.5) * dist, Jx, Jz);
      const T imp = impedance(m.con_solimp, dist);
      push_row(r, Jz, 0, kRowNormal, m.sph_diag, imp, K, Bd, dist, qd);
      push_row(r, Jx, 0, kRowTangent, m.sph_diag, imp, T(0), Bd, T(0), qd);
      mask |= 1u << 2;
    }
  }
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int g = 0; g < 4; g++) {  // thigh, shin (knee link), tarsus, toe
      const int cap = 4 * L + g;
      CASSIE_UNROLL
      for (int e = 0; e < 2; e++) {
        T ex, ez;
        const T* ep = e ? m.cap_from[cap] : m.cap_to[cap];
        rot(k.c[L][g], k.s[L][g], ep[0], ep[1], ex, ez);
        const T dist = cp.cap_dist[cap][e];
        if ((ALL || g == kToe) && !(dist > T(0))) {
          T Jx[8], Jz[8];
          point_jac(m, k, L, g, ex, ez - m.cap_r[cap] - T(0.5) * dist, Jx, Jz);
          const T imp = impedance(m.con_solimp, dist);
          push_row(r, Jz, L, kRowNormal, m.cap_diag[cap], imp, K, Bd, dist, qd);
          push_row(r, Jx, L, kRowTangent, m.cap_diag[cap], imp, T(0), Bd, T(0), qd);
          mask |= 1u << (2 * (2 + cap) + e);
        }
      }
    }
  }
  if (nlimit) *nlimit = nlim;
  return mask;
}

template <typename T>
CASSIE_HD void expand_row(const T J8[8], int leg, T x[kNV]) {
  x[0] = J8[0]; x[1] = J8[1]; x[2] = J8[2];
  CASSIE_UNROLL
  for (int b = 0; b < kLegLinks; b++) {
    x[3 + b] = leg ? T(0) : J8[3 + b];
    x[8 + b] = leg ? J8[3 + b] : T(0);
  }
}

// planar restriction of mj_constraintUpdate [EXT] for one row / contact pair (PGS warm start)
template <typename T>
CASSIE_HD void warm_force(const PlanarModel<T>& m, Rows<T>& r, int i, const T* jar) {
  const int tp = r.type[i];
  if (tp == kRowEq) r.f[i] = -jar[i] / r.R[i];
  else if (tp == kRowLimit) r.f[i] = jar[i] < T(0) ? -jar[i] / r.R[i] : T(0);
  else if (tp == kRowNormal) {
    const T mu = m.con_mu / Num<T>::sqrt_(m.impratio);
    const T N = jar[i] * mu, U1 = jar[i + 1] * m.con_mu, Tn = Num<T>::abs_(U1);
    const T D0 = T(1) / r.R[i], D1 = T(1) / r.R[i + 1];
    if (N >= mu * Tn || (Tn <= T(0) && N >= T(0))) { r.f[i] = T(0); r.f[i + 1] = T(0); }
    else if (mu * N + Tn <= T(0) || (Tn <= T(0) && N < T(0))) { r.f[i] = -D0 * jar[i]; r.f[i + 1] = -D1 * jar[i + 1]; }
    else {
      T den = mu * mu * (T(1) + mu * mu);
      const T Dm = D0 / (den > T(kMinVal) ? den : T(kMinVal));
      const T NmT = N - mu * Tn;
      r.f[i] = -Dm * NmT * mu;
      r.f[i + 1] = -r.f[i] / Tn * U1 * m.con_mu;
    }
  }
}

// mj_solPGS [EXT], planar rows.  Returns the number of sweeps done.
template <typename T>
CASSIE_HD int solve_pgs(const PlanarModel<T>& m, Rows<T>& r) {
  const int n = r.n;
  const T scale = T(1) / (m.meaninertia * T(kNV));
  int iter = 0;
  while (iter < m.iterations) {
    T improvement = T(0);
    for (int i = 0; i < n;) {
      const int tp = r.type[i];
      T res0 = r.b[i];
      for (int c = 0; c < n; c++) res0 += r.A[i][c] * r.f[c];
      const T old0 = r.f[i];
      if (tp != kRowNormal) {
        T f = old0 - res0 / r.A[i][i];
        if (tp != kRowEq && f < T(0)) f = T(0);
        const T d = f - old0;
        const T change = T(0.5) * d * d * r.A[i][i] + d * res0;
        if (change > T(1e-10)) { f = old0; } else improvement -= change;
        r.f[i] = f;
        i += 1;
      } else {
        T res1 = r.b[i + 1];
        for (int c = 0; c < n; c++) res1 += r.A[i + 1][c] * r.f[c];
        const T old1 = r.f[i + 1];
        const T A00 = r.A[i][i], A01 = r.A[i][i + 1], A11 = r.A[i + 1][i + 1];
        T f0 = old0, f1 = old1;
        if (f0 < T(kMinVal)) {
          f0 -= res0 / A00;
          if (f0 < T(0)) f0 = T(0);
          f1 = T(0);
        } else {
          const T denom = f0 * (A00 * f0 + A01 * f1) + f1 * (A01 * f0 + A11 * f1);
          if (denom >= T(kMinVal)) {
            T x = -(f0 * res0 + f1 * res1) / denom;
            if (f0 + x * f0 < T(0)) x = T(-1);
            f0 += x * old0;
            f1 += x * old1;
          }
        }
        // friction update with the normal force fixed (mju_QCQP2 collapses to a clamp)
        const T bc = res1 - A11 * old1 + A01 * (f0 - old0);
        if (f0 < T(kMinVal)) f1 = T(0);
        else {
          T v = -bc / A11;
          const T vs = v / m.con_mu;  // QCQP works in the scaled variable x/mu, radius f0
          if (vs * vs - f0 * f0 >= T(1e-10)) v = v > T(0) ? m.con_mu * f0 : -m.con_mu * f0;
          f1 = v;
        }
        const T d0 = f0 - old0, d1 = f1 - old1;
        const T change = T(0.5) * (d0 * (A00 * d0 + A01 * d1) + d1 * (A01 * d0 + A11 * d1)) + d0 * res0 + d1 * res1;
        if (change > T(1e-10)) { f0 = old0; f1 = old1; } else improvement -= change;
        r.f[i] = f0; r.f[i + 1] = f1;
        i += 2;
      }
    }
    iter++;
    if (improvement * scale < m.tolerance) break;
  }
  return iter;
}

struct StepStats {
  int nrows;
  int sweeps;
  unsigned int contact_mask;
};

// ---------------------------------------------------------------------------------------
// Constraint solve, general path: any number of rows (<= kMaxRows), A and f in thread-local
// memory.  Returns the PGS sweep count and qfrc_constraint = J^T f in fc.
template <typename T>
CASSIE_HD int constraint_solve_general(const PlanarModel<T>& m, Rows<T>& r, const T LD[kNV][kNV], const T Dinv[kNV],
                                       const T qs[kNV], const T warm[kNV], T fc[kNV]) {
  const int n = r.n;
  // b = J qacc_smooth - aref ;  A = J M^-1 J^T + diag(R)
  for (int i = 0; i < n; i++) {
    T Bi[kNV];
    expand_row(r.J[i], r.leg[i], Bi);
    r.b[i] += dot8_dense(r.J[i], r.leg[i], qs);
    solve(LD, Dinv, Bi);
    for (int j = 0; j <= i; j++) {
      const T v = dot8_dense(r.J[j], r.leg[j], Bi);
      r.A[i][j] = v;
      r.A[j][i] = v;
    }
    r.A[i][i] += r.R[i];
  }
  // warm start: forces from qacc_warmstart, kept only if they beat zero (mj_fwdConstraint [EXT])
  {
    T jar[kMaxRows];
    for (int i = 0; i < n; i++) jar[i] = dot8_dense(r.J[i], r.leg[i], warm) + r.b[i] - dot8_dense(r.J[i], r.leg[i], qs);
    for (int i = 0; i < n; i++) {
      if (r.type[i] != kRowTangent) warm_force(m, r, i, jar);
    }
    T cost = T(0);
    for (int i = 0; i < n; i++) {
      T s = T(0);
      for (int c = 0; c < n; c++) s += r.A[i][c] * r.f[c];
      cost += r.f[i] * (T(0.5) * s + r.b[i]);
    }
    if (cost > T(0))
      for (int i = 0; i < n; i++) r.f[i] = T(0);
  }
  const int sweeps = solve_pgs(m, r);
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) fc[i] = T(0);
  for (int i = 0; i < n; i++) {
    const T f = r.f[i];
    const int leg = r.leg[i];
    fc[0] += r.J[i][0] * f; fc[1] += r.J[i][1] * f; fc[2] += r.J[i][2] * f;
    CASSIE_UNROLL
    for (int b = 0; b < kLegLinks; b++) {
      const T v = r.J[i][3 + b] * f;
      fc[3 + b] += leg ? T(0) : v;
      fc[8 + b] += leg ? v : T(0);
    }
  }
  return sweeps;
}

// ---------------------------------------------------------------------------------------
// Constraint solve, fast path: the standing / squatting / walking regime has the 4 connect rows
// and at most 4 floor contacts (no joint-limit rows), i.e. <= 12 rows in the fixed order
// [eq Lx, eq Lz, eq Rx, eq Rz, (normal, tangent) x 4].  Everything -- A (78 symmetric entries), b,
// f -- then lives in registers and the PGS sweep is fully unrolled and branch-free; missing
// contact pairs are padded with inert rows (J = 0, A_ii = 1, b = 0).  Same arithmetic as the
// general path (mj_solPGS [EXT]) except that 1/A_ii is precomputed.
constexpr int kFastRows = 12;
CASSIE_HD constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// NS scalar rows (4 connect rows + NS - 4 joint-limit slots, the latter clamped at f >= 0) followed by NC
// contact pairs.  nlimit = joint-limit rows actually present (<= NS - 4).
template <int NS, int NC, typename T>
CASSIE_HD int constraint_solve_fast(const PlanarModel<T>& m, Rows<T>& r, const T LD[kNV][kNV], const T Dinv[kNV],
                                    const T qs[kNV], const T warm[kNV], T fc[kNV], int nlimit = 0) {
  constexpr int NR = NS + 2 * NC;  // rows handled by this instantiation
  int n = r.n;
  if (NS > 4 && nlimit < NS - 4) {
    // move the contact rows up so that they start at slot NS; the freed limit slots become inert rows
    const int shift = NS - 4 - nlimit;
    for (int i = n - 1; i >= 4 + nlimit; i--) {
      CASSIE_UNROLL
      for (int c = 0; c < 8; c++) r.J[i + shift][c] = r.J[i][c];
      r.leg[i + shift] = r.leg[i]; r.type[i + shift] = r.type[i];
      r.R[i + shift] = r.R[i]; r.b[i + shift] = r.b[i];
    }
    for (int i = 4 + nlimit; i < NS; i++) {
      CASSIE_UNROLL
      for (int c = 0; c < 8; c++) r.J[i][c] = T(0);
      r.leg[i] = 0; r.type[i] = kRowLimit;
      r.R[i] = T(1); r.b[i] = T(0);
    }
    n += shift;
  }
  for (int i = n; i < NR; i++) {  // inert padding
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) r.J[i][c] = T(0);
    r.leg[i] = 0;
    r.R[i] = T(1);
    r.b[i] = T(0);
  }
  // Assembly in a ROLLED row loop (J rows read from, A written to thread-local memory with run-time
  // indices), then A / b / jar are pulled into registers with static indices for the sweeps.  Unrolled,
  // this block was 5 k straight-line instructions that ran ~3x slower per instruction than loop code:
  // instruction fetch stalls at every 128-byte line (profiles/r1m_squat_osc_final.txt).
  CASSIE_ROLL
  for (int i = 0; i < NR; i++) {
    T J8[8], Bi[kNV];
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) J8[c] = r.J[i][c];
    const int leg = r.leg[i];
    expand_row(J8, leg, Bi);
    const T maref = r.b[i];
    r.b[i] = maref + dot8_dense(J8, leg, qs);
    r.f[i] = maref + dot8_dense(J8, leg, warm);   // jar, parked in f until the warm start below
    if (i < n) solve(LD, Dinv, Bi);
    CASSIE_ROLL
    for (int j = 0; j <= i; j++) r.A[i][j] = dot8_dense(r.J[j], r.leg[j], Bi);
    r.A[i][i] += r.R[i];
  }
  T A[NR * (NR + 1) / 2], b[NR], f[NR], jar[NR];
  CASSIE_UNROLL
  for (int i = 0; i < NR; i++) {
    b[i] = r.b[i];
    jar[i] = r.f[i];
    CASSIE_UNROLL
    for (int j = 0; j < NR; j++)
      if (j <= i) A[tri(i, j)] = r.A[i][j];
  }
  // warm start (mj_constraintUpdate [EXT] on jar = J qacc_warmstart - aref)
  CASSIE_UNROLL
  for (int i = 0; i < NS; i++) {
    const T fw = -jar[i] * Num<T>::rcp_(r.R[i]);          // R >= kMinVal
    f[i] = (i >= 4 && !(jar[i] < T(0))) ? T(0) : fw;   // limit rows: force only when violated
  }
  {
    const T mu = m.con_mu / Num<T>::sqrt_(m.impratio);
    CASSIE_UNROLL
    for (int i = NS; i < NR; i += 2) {
      const T N = jar[i] * mu, U1 = jar[i + 1] * m.con_mu, Tn = Num<T>::abs_(U1);
      const T D0 = Num<T>::rcp_(r.R[i]), D1 = Num<T>::rcp_(r.R[i + 1]);
      T f0, f1;
      if (N >= mu * Tn || (Tn <= T(0) && N >= T(0))) { f0 = T(0); f1 = T(0); }
      else if (mu * N + Tn <= T(0) || (Tn <= T(0) && N < T(0))) { f0 = -D0 * jar[i]; f1 = -D1 * jar[i + 1]; }
      else {
        T den = mu * mu * (T(1) + mu * mu);
        const T Dm = D0 / (den > T(kMinVal) ? den : T(kMinVal));
        f0 = -Dm * (N - mu * Tn) * mu;
        f1 = -f0 / Tn * U1 * m.con_mu;
      }
      f[i] = f0; f[i + 1] = f1;
    }
  }
  {
    T cost = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < NR; i++) {
      T s = T(0);
      CASSIE_UNROLL
      for (int c = 0; c < NR; c++) s += A[tri(i, c)] * f[c];
      cost += f[i] * (T(0.5) * s + b[i]);
    }
    if (cost > T(0)) {
      CASSIE_UNROLL
      for (int i = 0; i < NR; i++) f[i] = T(0);
    }
  }
  T inv[NR];
  CASSIE_UNROLL
  for (int i = 0; i < NR; i++) inv[i] = Num<T>::rcp_(A[tri(i, i)]);
  const T scale = T(1) / (m.meaninertia * T(kNV));
  const T mu = m.con_mu, inv_mu = T(1) / mu;
  int iter = 0;
  // Latency-oriented sweep.  One accumulator per row carries its residual: acc_i = b_i + sum_c A_ic f_c with the
  // forces of the previous sweep for c >= i and of this sweep for c < i.  Whenever a force is final it is
  // "published": every other row adds its term (independent FMAs, they fill the issue slots under the
  // Gauss-Seidel dependency chain), and the row itself restarts its accumulator for the next sweep
  // (acc_i = b_i + A_ii f_i, to which the later columns then add in ascending order -- no incremental drift:
  // every residual is a fresh sum of <= NR terms).  The chain through the freshly updated forces is one FMA
  // per row plus the projection.  The per-block "cost went up -> revert" test of mj_solPGS is dropped here:
  // every block update is an exact minimisation of a convex sub-problem, so the change is <= 0 in exact
  // arithmetic and the test can only fire on rounding noise (the general path keeps it).
  T acc[NR], rden[NC > 0 ? NC : 1];
  CASSIE_UNROLL
  for (int i = 0; i < NR; i++) {
    T sacc = b[i];
    CASSIE_UNROLL
    for (int c = 0; c < NR; c++)
      if (c >= i) sacc += A[tri(i, c)] * f[c];
    acc[i] = sacc;
  }
  CASSIE_UNROLL
  for (int p = 0; p < NC; p++) {
    const int i = NS + 2 * p;
    const T o0 = f[i], o1 = f[i + 1];
    const T denom = o0 * (A[tri(i, i)] * o0 + A[tri(i + 1, i)] * o1) + o1 * (A[tri(i + 1, i)] * o0 + A[tri(i + 1, i + 1)] * o1);
    rden[p] = denom >= T(kMinVal) ? Num<T>::rcp_(denom) : T(0);
  }
  while (iter < m.iterations) {
    T improvement = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < NS; i++) {  // scalar rows: connects unbounded, joint limits f >= 0
      const T res = acc[i];
      T fn = f[i] - res * inv[i];
      if (i >= 4) fn = Num<T>::max_(fn, T(0));
      const T d = fn - f[i];
      f[i] = fn;
      improvement -= d * (T(0.5) * d * A[tri(i, i)] + res);
      CASSIE_UNROLL
      for (int rr = 0; rr < NR; rr++) acc[rr] = rr == i ? b[i] + A[tri(i, i)] * fn : acc[rr] + A[tri(rr, i)] * fn;
    }
    CASSIE_UNROLL
    for (int p = 0; p < NC; p++) {  // elliptic contact: normal + one tangent, updated as a pair
      const int i = NS + 2 * p;
      const T old0 = f[i], old1 = f[i + 1];
      const T A00 = A[tri(i, i)], A01 = A[tri(i + 1, i)], A11 = A[tri(i + 1, i + 1)];
      const T res0 = acc[i];
      const T res1 = acc[i + 1] + A01 * old0;  // row i+1 has not seen column i of this pair yet: old force
      // (a) normal / ray update
      const T fa = Num<T>::max_(old0 - res0 * inv[i], T(0));
      const T x = Num<T>::max_(-(old0 * res0 + old1 * res1) * rden[p], T(-1));
      const T f0 = old0 < T(kMinVal) ? fa : old0 + x * old0;
      // (b) friction update with the normal force fixed (mju_QCQP2 collapses to a clamp)
      const T bc = (res1 - A11 * old1 - A01 * old0) + A01 * f0;
      T v = -bc * inv[i + 1];
      const T lim = mu * f0;
      if (Num<T>::kExactConeTest) {
        const T vs = v * inv_mu;
        v = (vs * vs - f0 * f0 >= T(1e-10)) ? (v > T(0) ? lim : -lim) : v;
      } else {
        // fp32: plain clamp (two FMNMX instead of a compare -> predicate -> select chain).  It differs from the
        // test above only inside its 1e-10 guard band, |v| in [lim, mu sqrt(f0^2 + 1e-10)): below 1e-5 N.
        v = Num<T>::min_(Num<T>::max_(v, -lim), lim);
      }
      const T f1 = f0 < T(kMinVal) ? T(0) : v;
      const T d0 = f0 - old0, d1 = f1 - old1;
      improvement -= d0 * (T(0.5) * d0 * A00 + A01 * d1 + res0) + d1 * (T(0.5) * d1 * A11 + res1);
      f[i] = f0;
      f[i + 1] = f1;
      CASSIE_UNROLL
      for (int rr = 0; rr < NR; rr++)
        if (rr != i + 1) acc[rr] = rr == i ? b[i] + A00 * f0 : acc[rr] + A[tri(rr, i)] * f0;
      CASSIE_UNROLL
      for (int rr = 0; rr < NR; rr++) acc[rr] = rr == i + 1 ? b[i + 1] + A11 * f1 : acc[rr] + A[tri(rr, i + 1)] * f1;
      const T denom = f0 * (A00 * f0 + A01 * f1) + f1 * (A01 * f0 + A11 * f1);
      rden[p] = denom >= T(kMinVal) ? Num<T>::rcp_(denom) : T(0);
    }
    iter++;
    if (improvement * scale < m.tolerance) break;
  }
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) fc[i] = T(0);
  CASSIE_UNROLL
  for (int i = 0; i < NR; i++) {
    const T fi = f[i];
    const int leg = r.leg[i];
    fc[0] += r.J[i][0] * fi; fc[1] += r.J[i][1] * fi; fc[2] += r.J[i][2] * fi;
    CASSIE_UNROLL
    for (int bb = 0; bb < kLegLinks; bb++) {
      const T v = r.J[i][3 + bb] * fi;
      fc[3 + bb] += leg ? T(0) : v;
      fc[8 + bb] += leg ? v : T(0);
    }
  }
  return iter;
}

// Anything outside the common regime active?  (a violated joint limit, or a floor contact of the
// pelvis sphere / thigh / shin / tarsus capsules)
template <typename T>
CASSIE_HD bool rare_rows_active(const PlanarModel<T>& m, const ConPos<T>& cp, const T* q) {
  bool rare = !(cp.sph_dist > T(0));
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int g = 0; g < 3; g++) rare = rare || !(cp.cap_dist[4 * L + g][0] > T(0)) || !(cp.cap_dist[4 * L + g][1] > T(0));
  }
  CASSIE_UNROLL
  for (int j = 3; j < kNV; j++)
    if (m.has_limit[j]) rare = rare || (q[j] - m.lim_lo[j] < T(0)) || (m.lim_hi[j] - q[j] < T(0));
  return rare;
}

// Out-of-line slow path: all row kinds, any row count, thread-local memory.  Kept out of the hot
// instruction stream on purpose: inlined, its mere presence cost the Jacobian-mode kernel 30 % (register
// allocation + instruction fetch; profiles/r1_variants.txt).
template <typename T>
CASSIE_COLD void constraints_cold(const PlanarModel<T>& m, const Kin<T>& k, const ConPos<T>& cp, const T* q, const T* qd,
                                  Rows<T>& r, const T (&LD)[kNV][kNV], const T* Dinv, const T* qs, const T* warm, T* fc,
                                  int* sweeps, unsigned int* mask) {
  int nlimit = 0;
  *mask = make_rows<true>(m, k, cp, q, qd, r, &nlimit);
  // middle tier: up to 4 violated joint limits and up to 4 floor contacts of any geom still fit the
  // register solver (robots thrown around by random actions live here); everything else is general
#ifdef CASSIE_HOST_HARNESS
  const bool fast_ok = !cassie_force_general_path;
#else
  const bool fast_ok = true;
#endif
  const int ncontact = (r.n - 4 - nlimit) / 2;
  if (fast_ok && nlimit <= 4 && ncontact <= 4) *sweeps = constraint_solve_fast<8, 4>(m, r, LD, Dinv, qs, warm, fc, nlimit);
  else *sweeps = constraint_solve_general(m, r, LD, Dinv, qs, warm, fc);
}

// One mj_step [EXT] (Cassie2d.cpp:92): forward dynamics, constraint solve, semi-implicit Euler
// with implicit joint damping.  q, qd, warm are updated in place; u is in ctrl units.
// TG = type of the position pass (angles -> sin/cos -> pivots -> constraint violations): double in
// the fp32 build (mg = the same model in double), T in the fp64 build.
template <typename T, typename TG>
CASSIE_HD void physics_step(const PlanarModel<T>& m, const PlanarModel<TG>& mg, T q[kNV], T qd[kNV], T warm[kNV],
                            const T u[kNU], Rows<T>& r, StepStats* st) {
  Kin<T> k;
  ConPos<T> cp;
  {
    TG qg[kNV];
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) qg[i] = (TG)q[i];
    Kin<TG> kg;
    fk_positions(mg, qg, kg);
    constraint_positions(mg, kg, qg[1], cp);
    cast_kin_positions(kg, k);
  }
  fk_velocities(m, qd, k);
  T M[kNV][kNV], LD[kNV][kNV], Dinv[kNV];
  mass_matrix(m, k, M);
  T fs[kNV];  // qfrc_smooth = passive - bias + actuator
  bias_forces(m, k, fs);
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) fs[i] = -fs[i] - m.damping[i] * qd[i];
  CASSIE_UNROLL
  for (int a = 0; a < kNU; a++) {
    T c = u[a];
    c = c < m.act_lo[a] ? m.act_lo[a] : (c > m.act_hi[a] ? m.act_hi[a] : c);
    CASSIE_UNROLL
    for (int i = 3; i < kNV; i++)
      if (m.act_dof[a] == i) fs[i] += m.act_gear[a] * c;
  }
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < kNV; j++)
      if (dof_anc(i, j)) LD[i][j] = M[i][j];
  }
  factor(LD, Dinv);
  T qs[kNV];  // qacc_smooth
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) qs[i] = fs[i];
  solve(LD, Dinv, qs);

  T fc[kNV];  // qfrc_constraint = J^T f
  int sweeps;
  unsigned int mask;
#ifdef CASSIE_HOST_HARNESS  // test hook: lets the harness exercise / count the general path on small row counts
  const bool fast_ok = !cassie_force_general_path;
#else
  const bool fast_ok = true;
#endif
  // The tier is chosen per WARP: lanes of one warp that took different tiers would run them one after the other,
  // and the middle tier of constraints_cold solves the common regime as well (its limit slots stay empty).
  const bool rare = !fast_ok || rare_rows_active(m, cp, q);
#ifdef __CUDA_ARCH__
  const bool warp_rare = __any_sync(__activemask(), rare);
#else
  const bool warp_rare = rare;
#endif
  if (!warp_rare) {
    // common regime: 4 connect rows + <= 4 toe contacts, everything in registers
    mask = make_rows<false>(m, k, cp, q, qd, r, (int*)nullptr);
    // 8-row (two contacts) or 12-row variant, chosen per WARP so that lanes never run both
#ifdef __CUDA_ARCH__
    const bool wide = __any_sync(__activemask(), r.n > 8);
#else
    const bool wide = r.n > 8;
#endif
    if (wide) sweeps = constraint_solve_fast<4, 4>(m, r, LD, Dinv, qs, warm, fc);
    else sweeps = constraint_solve_fast<4, 2>(m, r, LD, Dinv, qs, warm, fc);
  } else {
    constraints_cold(m, k, cp, q, qd, r, LD, Dinv, qs, warm, fc, &sweeps, &mask);
  }
  const int n = r.n;
  // qacc = qacc_smooth + M^-1 qfrc_constraint
  T dq[kNV];
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) dq[i] = fc[i];
  solve(LD, Dinv, dq);
  // mj_Euler [EXT]: (M + h D) qacc' = qfrc_smooth + qfrc_constraint
  const T h = m.timestep;
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) {
    warm[i] = qs[i] + dq[i];
    fs[i] += fc[i];
    M[i][i] += h * m.damping[i];
  }
  factor(M, Dinv);
  solve(M, Dinv, fs);
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) {
    qd[i] += h * fs[i];
    q[i] += h * qd[i];
  }
  if (st) { st->nrows = n; st->sweeps = sweeps; st->contact_mask = mask; }
}

}  // namespace cassie
