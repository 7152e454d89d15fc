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
#define CASSIE_UNROLL _Pragma("unroll")
#else
#define CASSIE_HD inline
#define CASSIE_UNROLL
#endif

namespace cassie {

template <typename T> struct Num;
template <> struct Num<float> {
  static CASSIE_HD void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
  static CASSIE_HD float sqrt_(float a) { return sqrtf(a); }
  static CASSIE_HD float abs_(float a) { return fabsf(a); }
  static CASSIE_HD float pow_(float a, float b) { return powf(a, b); }
  static CASSIE_HD float exp_(float a) { return expf(a); }
};
template <> struct Num<double> {
  static CASSIE_HD void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
  static CASSIE_HD double sqrt_(double a) { return sqrt(a); }
  static CASSIE_HD double abs_(double a) { return fabs(a); }
  static CASSIE_HD double pow_(double a, double b) { return pow(a, b); }
  static CASSIE_HD double exp_(double a) { return exp(a); }
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
  T w[2][kLegLinks], vx[2][kLegLinks], vz[2][kLegLinks];
};

template <typename T>
CASSIE_HD void rot(T c, T s, T x, T z, T& ox, T& oz) {
  ox = x * c + z * s;
  oz = z * c - x * s;
}

// mj_kinematics + mj_comVel [EXT] / RBDL UpdateKinematics (DynamicModel.cpp:237-242), planar
template <typename T>
CASSIE_HD void forward_kinematics(const PlanarModel<T>& m, const T* q, const T* qd, Kin<T>& k) {
  Num<T>::sincos_(q[2] - m.pel_ref[2], &k.s0, &k.c0);
  k.w0 = qd[2];
  k.v0x = qd[0];
  k.v0z = qd[1];
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T alpha[kLegLinks];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int p = link_parent(a);
      const int dof = 3 + 5 * L + a;
      T ap, cp, sp, ppx, ppz, wp, vpx, vpz;
      if (p < 0) {
        ap = q[2] - m.pel_ref[2]; cp = k.c0; sp = k.s0; ppx = T(0); ppz = T(0);
        wp = k.w0; vpx = k.v0x; vpz = k.v0z;
      } else {
        ap = alpha[p]; cp = k.c[L][p]; sp = k.s[L][p]; ppx = k.px[L][p]; ppz = k.pz[L][p];
        wp = k.w[L][p]; vpx = k.vx[L][p]; vpz = k.vz[L][p];
      }
      alpha[a] = ap + m.sgn[L][a] * q[dof] + m.ang0[L][a];
      Num<T>::sincos_(alpha[a], &k.s[L][a], &k.c[L][a]);
      T dx, dz;
      rot(cp, sp, m.off[L][a][0], m.off[L][a][1], dx, dz);
      k.px[L][a] = ppx + dx;
      k.pz[L][a] = ppz + dz;
      k.w[L][a] = wp + m.sgn[L][a] * qd[dof];
      k.vx[L][a] = vpx + wp * dz;   // w y^ x d = w (d.z, -d.x)
      k.vz[L][a] = vpz - wp * dx;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Mass matrix (lower triangle, structural non-zeros only) + armature.
// mj_crb [EXT] / RBDL CompositeRigidBodyAlgorithm + rotor inertia (DynamicModel.cpp:267-272).
// Composite (m, h = m c, Io) per subtree about the pelvis pivot; for hinges i (ancestor) and j:
//   M_ij = s_i s_j [ Io(j) - (p_i + p_j).h(j) + m(j) p_i.p_j ]
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
    T cm[kLegLinks], chx[kLegLinks], chz[kLegLinks], cI[kLegLinks];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      T rx, rz;
      rot(k.c[L][a], k.s[L][a], m.com[L][a][0], m.com[L][a][1], rx, rz);
      const T cx = k.px[L][a] + rx, cz = k.pz[L][a] + rz;
      cm[a] = m.mass[L][a];
      chx[a] = cm[a] * cx; chz[a] = cm[a] * cz;
      cI[a] = m.inertia[L][a] + cm[a] * (cx * cx + cz * cz);
    }
    // accumulate subtrees: toe->tarsus->knee->thigh, rod->thigh
    cm[kTarsus] += cm[kToe]; chx[kTarsus] += chx[kToe]; chz[kTarsus] += chz[kToe]; cI[kTarsus] += cI[kToe];
    cm[kKnee] += cm[kTarsus]; chx[kKnee] += chx[kTarsus]; chz[kKnee] += chz[kTarsus]; cI[kKnee] += cI[kTarsus];
    cm[kThigh] += cm[kKnee] + cm[kRod]; chx[kThigh] += chx[kKnee] + chx[kRod];
    chz[kThigh] += chz[kKnee] + chz[kRod]; cI[kThigh] += cI[kKnee] + cI[kRod];
    tm += cm[kThigh]; thx += chx[kThigh]; thz += chz[kThigh]; tI += cI[kThigh];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int i = 3 + 5 * L + a;
      const T sa = m.sgn[L][a], pax = k.px[L][a], paz = k.pz[L][a];
      // slides and pitch (pitch pivot = origin of the relative frame, sign +1)
      M[i][0] = sa * (chz[a] - cm[a] * paz);
      M[i][1] = -sa * (chx[a] - cm[a] * pax);
      M[i][2] = sa * (cI[a] - (pax * chx[a] + paz * chz[a]));
      CASSIE_UNROLL
      for (int b = 0; b < kLegLinks; b++) {
        if (b <= a && link_anc(a, b)) {
          const T pbx = k.px[L][b], pbz = k.pz[L][b];
          T v = cI[a] - ((pax + pbx) * chx[a] + (paz + pbz) * chz[a]) + cm[a] * (pax * pbx + paz * pbz);
          v *= sa * m.sgn[L][b];
          if (a == b) v += m.armature[i];
          M[i][3 + 5 * L + b] = v;
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
// centripetal terms; torques are taken about the pelvis pivot.
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
      T apx, apz, wp, ppx, ppz;
      if (p < 0) { apx = T(0); apz = T(0); wp = k.w0; ppx = T(0); ppz = T(0); }
      else { apx = ax[p]; apz = az[p]; wp = k.w[L][p]; ppx = k.px[L][p]; ppz = k.pz[L][p]; }
      const T w2p = wp * wp;
      ax[a] = apx - w2p * (k.px[L][a] - ppx);
      az[a] = apz - w2p * (k.pz[L][a] - ppz);
      T rx, rz;
      rot(k.c[L][a], k.s[L][a], m.com[L][a][0], m.com[L][a][1], rx, rz);
      const T w2 = k.w[L][a] * k.w[L][a];
      fx[a] = m.mass[L][a] * (ax[a] - w2 * rx);
      fz[a] = m.mass[L][a] * (az[a] - w2 * rz + g);
      const T cx = k.px[L][a] + rx, cz = k.pz[L][a] + rz;
      n[a] = cz * fx[a] - cx * fz[a];
    }
    fx[kTarsus] += fx[kToe]; fz[kTarsus] += fz[kToe]; n[kTarsus] += n[kToe];
    fx[kKnee] += fx[kTarsus]; fz[kKnee] += fz[kTarsus]; n[kKnee] += n[kTarsus];
    fx[kThigh] += fx[kKnee] + fx[kRod]; fz[kThigh] += fz[kKnee] + fz[kRod]; n[kThigh] += n[kKnee] + n[kRod];
    Fx += fx[kThigh]; Fz += fz[kThigh]; N += n[kThigh];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++)
      bias[3 + 5 * L + a] = m.sgn[L][a] * (n[a] - (k.pz[L][a] * fx[a] - k.px[L][a] * fz[a]));
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
    const T inv = T(1) / M[k][k];
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

template <typename T>
CASSIE_HD T impedance(const T* si, T pos) {  // getimpedance [EXT], margin = 0
  if (si[0] == si[1] || si[2] <= T(kMinVal)) return T(0.5) * (si[0] + si[1]);
  T x = Num<T>::abs_(pos / si[2]);
  if (x >= T(1)) return si[1];
  if (x <= T(0)) return si[0];
  T y;
  if (si[4] == T(1)) y = x;
  else if (x <= si[3]) y = Num<T>::pow_(x, si[4]) / Num<T>::pow_(si[3], si[4] - T(1));
  else y = T(1) - Num<T>::pow_(T(1) - x, si[4]) / Num<T>::pow_(T(1) - si[3], si[4] - T(1));
  return si[0] + y * (si[1] - si[0]);
}

// point Jacobian (x row, z row) of a point P (relative to the pelvis pivot) fixed to leg link
// `a` of leg L;  a = -1: pelvis.  mj_jac [EXT] / RBDL CalcPointJacobian, planar.
template <typename T>
CASSIE_HD void point_jac(const PlanarModel<T>& m, const Kin<T>& k, int L, int a, T Px, T Pz, T Jx[8], T Jz[8]) {
  Jx[0] = T(1); Jx[1] = T(0); Jx[2] = Pz;
  Jz[0] = T(0); Jz[1] = T(1); Jz[2] = -Px;
  CASSIE_UNROLL
  for (int b = 0; b < kLegLinks; b++) {
    const bool on = a >= 0 && (a == b || b == kThigh || (a != kRod && b != kRod && b < a));
    const T sb = on ? m.sgn[L][b] : T(0);
    Jx[3 + b] = sb * (Pz - k.pz[L][b]);
    Jz[3 + b] = -sb * (Px - k.px[L][b]);
  }
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
  const T Rv = (T(1) - imp) * diag / imp;
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
  K = T(1) / (kd > T(kMinVal) ? kd : T(kMinVal));
  T bd = dmax * tc;
  Bd = T(2) / (bd > T(kMinVal) ? bd : T(kMinVal));
}

// mj_collision + mj_makeConstraint + mj_makeImpedance [EXT], canonical row order:
// connects (L, R), joint limits (dof order), contacts (pelvis sphere, then per capsule the 'to'
// end before the 'from' end).  Returns the public contact bit mask.
template <typename T>
CASSIE_HD unsigned int make_rows(const PlanarModel<T>& m, const Kin<T>& k, const T* q, const T* qd, Rows<T>& r) {
  r.n = 0;
  T K, Bd;
  // ---- connects
  kb_from_solref(m, m.eq_solref, m.eq_solimp, K, Bd);
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T ax, az, bx, bz;
    rot(k.c[L][kRod], k.s[L][kRod], m.eq_a1[L][0], m.eq_a1[L][1], ax, az);
    rot(k.c[L][kTarsus], k.s[L][kTarsus], m.eq_a2[L][0], m.eq_a2[L][1], bx, bz);
    const T P1x = k.px[L][kRod] + ax, P1z = k.pz[L][kRod] + az;
    const T P2x = k.px[L][kTarsus] + bx, P2z = k.pz[L][kTarsus] + bz;
    T J1x[8], J1z[8], J2x[8], J2z[8];
    point_jac(m, k, L, kRod, P1x, P1z, J1x, J1z);
    point_jac(m, k, L, kTarsus, P2x, P2z, J2x, J2z);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) { J1x[c] -= J2x[c]; J1z[c] -= J2z[c]; }
    const T rx = P1x - P2x, rz = P1z - P2z;
    const T imp = impedance(m.eq_solimp, Num<T>::sqrt_(rx * rx + rz * rz));
    push_row(r, J1x, L, kRowEq, m.eq_diag[L], imp, K, Bd, rx, qd);
    push_row(r, J1z, L, kRowEq, m.eq_diag[L], imp, K, Bd, rz, qd);
  }
  // ---- joint limits
  kb_from_solref(m, m.lim_solref, m.lim_solimp, K, Bd);
  CASSIE_UNROLL
  for (int j = 3; j < kNV; j++) {
    if (m.has_limit[j]) {
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
        }
      }
    }
  }
  // ---- floor contacts (plane z = 0, normal +z).  Elliptic friction rows: K = 0, pos = 0.
  unsigned int mask = 0;
  kb_from_solref(m, m.con_solref, m.con_solimp, K, Bd);
  const T height = q[1] - m.pel_ref[1] + m.pel_org[1];  // world z of the pelvis pivot
  {
    T cx, cz;
    rot(k.c0, k.s0, m.sph_c[0], m.sph_c[1], cx, cz);
    const T dist = height + cz - m.sph_r;
    if (!(dist > T(0))) {
      T Jx[8], Jz[8];
      const T Pz = T(0.5) * dist - height;
      point_jac(m, k, 0, -1, cx, Pz, Jx, Jz);
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
        const T Px = k.px[L][g] + ex;
        const T dist = height + k.pz[L][g] + ez - m.cap_r[cap];
        if (!(dist > T(0))) {
          T Jx[8], Jz[8];
          point_jac(m, k, L, g, Px, T(0.5) * dist - height, Jx, Jz);
          const T imp = impedance(m.con_solimp, dist);
          push_row(r, Jz, L, kRowNormal, m.cap_diag[cap], imp, K, Bd, dist, qd);
          push_row(r, Jx, L, kRowTangent, m.cap_diag[cap], imp, T(0), Bd, T(0), qd);
          mask |= 1u << (2 * (2 + cap) + e);
        }
      }
    }
  }
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

// One mj_step [EXT] (Cassie2d.cpp:92): forward dynamics, constraint solve, semi-implicit Euler
// with implicit joint damping.  q, qd, warm are updated in place; u is in ctrl units.
template <typename T>
CASSIE_HD void physics_step(const PlanarModel<T>& m, T q[kNV], T qd[kNV], T warm[kNV], const T u[kNU],
                            Rows<T>& r, StepStats* st) {
  Kin<T> k;
  forward_kinematics(m, q, qd, k);
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

  const unsigned int mask = make_rows(m, k, q, qd, r);
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
  // qfrc_constraint = J^T f ; qacc = qacc_smooth + M^-1 qfrc_constraint
  T fc[kNV];
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
