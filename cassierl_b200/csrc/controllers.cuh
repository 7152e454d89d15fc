// Controller side of the Cassie2d facade, batched: what the reference computes with RBDL +
// Eigen before every mj_step (CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:86-237,
// src/DynamicState.cpp:45-91, src/DynamicModel.cpp:237-367, src/HelperFunctions.h:8-29).
// Everything here runs on the CONTROLLER model (`ctrl`, the RBDL loader's view of the MJCF,
// DynamicModel.cpp:84-103), the physics step on the `phys` model -- exactly the split the
// reference has between RBDL and MuJoCo.
//
// Planar reductions: the y rows of Jeq/Jc are identically zero, so the
// 6x6 matrix Jeq M^-1 Jeq^T has the same non-zero singular values as its 4x4 x/z block and
// pseudoinverse(.,1e-3) (Cassie2d.cpp:134, OSC_RBDL.cpp:171) acts on that block only; of the
// 6-D site Jacobian only the rows [My, Fx, Fz] are ever multiplied by a non-zero wrench
// (Cassie2d.cpp:157-163).
#pragma once
#include "planar_engine.cuh"

namespace cassie {

enum StepMode { kModeTorque = 0, kModePd = 1, kModeJacobian = 2, kModeOsc = 3 };

// GetOperationalSpaceState inputs that come from the RBDL state stored at the START of the
// last Step* (Cassie2d.cpp:88,98,121,182 vs :223): world x/z position and velocity of
// body_center and of the mean of the front/rear contact sites of each foot.
template <typename T>
struct OpState {
  T body[4];   // x, z, xd, zd
  T left[4];
  T right[4];
};

// world position / velocity of a point fixed to a link (site_link: -1 pelvis, else 5L+a)
template <typename T>
CASSIE_HD void site_point(const PlanarModel<T>& m, const Kin<T>& k, const T* q, int s, T& x, T& z, T& xd, T& zd) {
  const int link = m.site_link[s];
  const T ox = m.site_off[s][0], oz = m.site_off[s][1];
  T rx = T(0), rz = T(0), px = T(0), pz = T(0), w = k.w0, vx = k.v0x, vz = k.v0z;
  rot(k.c0, k.s0, ox, oz, rx, rz);
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      if (link == 5 * L + a) {
        rot(k.c[L][a], k.s[L][a], ox, oz, rx, rz);
        px = k.px[L][a]; pz = k.pz[L][a]; w = k.w[L][a]; vx = k.vx[L][a]; vz = k.vz[L][a];
      }
    }
  }
  x = (q[0] - m.pel_ref[0] + m.pel_org[0]) + px + rx;
  z = (q[1] - m.pel_ref[1] + m.pel_org[1]) + pz + rz;
  xd = vx + w * rz;
  zd = vz - w * rx;
}

// DynamicModel::GetTargetPoints (DynamicModel.cpp:360-367) reduced to what
// GetOperationalSpaceState (Cassie2d.cpp:218-237) keeps.  Sites: 1 body_center, 2/3 left
// front/rear, 4/5 right front/rear (Cassie2d.cpp:34-36).
template <typename T>
CASSIE_HD void op_state_from_kin(const PlanarModel<T>& m, const Kin<T>& k, const T* q, OpState<T>& op) {
  site_point(m, k, q, 1, op.body[0], op.body[1], op.body[2], op.body[3]);
  T a[4], b[4];
  site_point(m, k, q, 2, a[0], a[1], a[2], a[3]);
  site_point(m, k, q, 3, b[0], b[1], b[2], b[3]);
  CASSIE_UNROLL
  for (int i = 0; i < 4; i++) op.left[i] = (a[i] + b[i]) / T(2);
  site_point(m, k, q, 4, a[0], a[1], a[2], a[3]);
  site_point(m, k, q, 5, b[0], b[1], b[2], b[3]);
  CASSIE_UNROLL
  for (int i = 0; i < 4; i++) op.right[i] = (a[i] + b[i]) / T(2);
}

// StateOperationalSpace memory order (RobotInterface.h:47-50): body_x[3] body_xd[3] left_x[3]
// left_xd[3] right_x[3] right_xd[3]; the [2] slots of left/right are never written by the
// reference (0 here); pitch / pitch rate come from the CURRENT state (Cassie2d.cpp:234-235).
template <typename T>
CASSIE_HD void op_state_array(const OpState<T>& op, const T* q, const T* qd, T o[18]) {
  o[0] = op.body[0]; o[1] = op.body[1]; o[2] = q[2];
  o[3] = op.body[2]; o[4] = op.body[3]; o[5] = qd[2];
  o[6] = op.left[0]; o[7] = op.left[1]; o[8] = T(0);
  o[9] = op.left[2]; o[10] = op.left[3]; o[11] = T(0);
  o[12] = op.right[0]; o[13] = op.right[1]; o[14] = T(0);
  o[15] = op.right[2]; o[16] = op.right[3]; o[17] = T(0);
}

// ---------------------------------------------------------------------------------------
// Jdot*qd of a link-fixed point = its acceleration at qdd = 0 (RBDL CalcPointAcceleration,
// DynamicModel.cpp:341-344).  Planar hinges are parallel, so link angular accelerations vanish
// and only centripetal terms accumulate along the chain.
template <typename T>
struct PivotAcc { T ax[2][kLegLinks], az[2][kLegLinks]; };

template <typename T>
CASSIE_HD void pivot_accelerations(const Kin<T>& k, PivotAcc<T>& pa) {
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int p = link_parent(a);
      T apx, apz, wp;
      if (p < 0) { apx = T(0); apz = T(0); wp = k.w0; }
      else { apx = pa.ax[L][p]; apz = pa.az[L][p]; wp = k.w[L][p]; }
      const T w2 = wp * wp;
      pa.ax[L][a] = apx - w2 * k.dx[L][a];
      pa.az[L][a] = apz - w2 * k.dz[L][a];
    }
  }
}

// ---------------------------------------------------------------------------------------
// symmetric 4x4 pseudo-inverse with singular-value cut-off (HelperFunctions.h:8-29 applied to
// Jeq M^-1 Jeq^T, which is symmetric PSD: singular values = eigenvalues).  Cyclic Jacobi.
template <typename T>
CASSIE_COLD void sym4_pinv_jacobi(T A[4][4], T tol, T P[4][4]) {
  T V[4][4];
  CASSIE_UNROLL
  for (int i = 0; i < 4; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < 4; j++) V[i][j] = i == j ? T(1) : T(0);
  }
  const T eps = sizeof(T) == 4 ? T(1e-7) : T(1e-16);
  for (int sweep = 0; sweep < 12; sweep++) {
    T off = T(0), dg = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < 4; i++) {
      dg += A[i][i] * A[i][i];
      CASSIE_UNROLL
      for (int j = 0; j < 4; j++)
        if (j > i) off += A[i][j] * A[i][j];
    }
    if (off <= eps * eps * dg) break;
    CASSIE_UNROLL
    for (int p = 0; p < 3; p++) {
      CASSIE_UNROLL
      for (int q = 0; q < 4; q++) {
        if (q > p) {
          const T apq = A[p][q];
          if (Num<T>::abs_(apq) > T(1e-37)) {
            const T theta = (A[q][q] - A[p][p]) / (T(2) * apq);
            const T t = (theta >= T(0) ? T(1) : T(-1)) / (Num<T>::abs_(theta) + Num<T>::sqrt_(theta * theta + T(1)));
            const T c = T(1) / Num<T>::sqrt_(t * t + T(1)), s = t * c;
            CASSIE_UNROLL
            for (int r = 0; r < 4; r++) {  // A <- A G
              const T arp = A[r][p], arq = A[r][q];
              A[r][p] = c * arp - s * arq;
              A[r][q] = s * arp + c * arq;
            }
            CASSIE_UNROLL
            for (int r = 0; r < 4; r++) {  // A <- G^T A
              const T apr = A[p][r], aqr = A[q][r];
              A[p][r] = c * apr - s * aqr;
              A[q][r] = s * apr + c * aqr;
            }
            CASSIE_UNROLL
            for (int r = 0; r < 4; r++) {
              const T vrp = V[r][p], vrq = V[r][q];
              V[r][p] = c * vrp - s * vrq;
              V[r][q] = s * vrp + c * vrq;
            }
          }
        }
      }
    }
  }
  CASSIE_UNROLL
  for (int i = 0; i < 4; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < 4; j++) P[i][j] = T(0);
  }
  CASSIE_UNROLL
  for (int e = 0; e < 4; e++) {
    const T lam = A[e][e];
    if (lam > tol) {
      const T inv = T(1) / lam;
      CASSIE_UNROLL
      for (int i = 0; i < 4; i++) {
        CASSIE_UNROLL
        for (int j = 0; j < 4; j++) P[i][j] += V[i][e] * V[j][e] * inv;
      }
    }
  }
}

// u = pinv(B, tol) * rhs for a 13x6 matrix B (Cassie2d.cpp:165, default tolerance 1e-4):
// one-sided Jacobi, B V = U S ;  u = sum_j v_j (b_j . rhs) / s_j^2 over s_j > tol.
template <typename T>
CASSIE_COLD void pinv13x6_apply_jacobi(T B[kNV][kNU], T tol, const T rhs[kNV], T u[kNU]) {
  T V[kNU][kNU];
  CASSIE_UNROLL
  for (int i = 0; i < kNU; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < kNU; j++) V[i][j] = i == j ? T(1) : T(0);
  }
  const T eps = sizeof(T) == 4 ? T(2e-7) : T(1e-15);
  for (int sweep = 0; sweep < 16; sweep++) {
    T off = T(0);
    CASSIE_UNROLL
    for (int p = 0; p < kNU - 1; p++) {
      CASSIE_UNROLL
      for (int q = 0; q < kNU; q++) {
        if (q > p) {
          T a = T(0), b = T(0), c = T(0);
          CASSIE_UNROLL
          for (int i = 0; i < kNV; i++) { a += B[i][p] * B[i][p]; b += B[i][q] * B[i][q]; c += B[i][p] * B[i][q]; }
          const T ab = Num<T>::sqrt_(a * b);
          if (Num<T>::abs_(c) > eps * ab && Num<T>::abs_(c) > T(1e-37)) {
            const T rel = Num<T>::abs_(c) / ab;
            off = rel > off ? rel : off;
            const T zeta = (b - a) / (T(2) * c);
            const T t = (zeta >= T(0) ? T(1) : T(-1)) / (Num<T>::abs_(zeta) + Num<T>::sqrt_(T(1) + zeta * zeta));
            const T cs = T(1) / Num<T>::sqrt_(T(1) + t * t), sn = cs * t;
            CASSIE_UNROLL
            for (int i = 0; i < kNV; i++) {
              const T bp = B[i][p], bq = B[i][q];
              B[i][p] = cs * bp - sn * bq;
              B[i][q] = sn * bp + cs * bq;
            }
            CASSIE_UNROLL
            for (int i = 0; i < kNU; i++) {
              const T vp = V[i][p], vq = V[i][q];
              V[i][p] = cs * vp - sn * vq;
              V[i][q] = sn * vp + cs * vq;
            }
          }
        }
      }
    }
    if (off <= eps) break;
  }
  CASSIE_UNROLL
  for (int i = 0; i < kNU; i++) u[i] = T(0);
  CASSIE_UNROLL
  for (int j = 0; j < kNU; j++) {
    T s2 = T(0), d = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) { s2 += B[i][j] * B[i][j]; d += B[i][j] * rhs[i]; }
    if (Num<T>::sqrt_(s2) > tol) {
      const T w = d / s2;
      CASSIE_UNROLL
      for (int i = 0; i < kNU; i++) u[i] += V[i][j] * w;
    }
  }
}

// Fast paths of the two pseudo-inverses.  pseudoinverse(A, tol) zeroes singular values <= tol
// (HelperFunctions.h:8-29); when ALL singular values exceed tol it is the plain inverse / least-squares
// solution.  Both routines certify that case with the bound  sigma_min >= 1 / ||A^-1||_F  and otherwise
// fall back to the (out-of-line) Jacobi SVD, so the result is the reference's in every case while the
// hot instruction stream carries a Cholesky / QR instead of ~60 Jacobi rotations.
template <typename T>
CASSIE_HD void sym4_pinv(T A[4][4], T tol, T P[4][4]) {
  // Cholesky A = L L^T, P = L^-T L^-1
  T L[4][4], Li[4][4], invd[4];
  bool ok = true;
  CASSIE_UNROLL
  for (int i = 0; i < 4; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < 4; j++) {
      if (j <= i) {
        T sacc = A[i][j];
        CASSIE_UNROLL
        for (int k = 0; k < 4; k++)
          if (k < j) sacc -= L[i][k] * L[j][k];
        if (i == j) {
          ok = ok && (sacc > T(0));
          L[i][i] = Num<T>::sqrt_(sacc > T(0) ? sacc : T(1));
          invd[i] = T(1) / L[i][i];
        } else {
          L[i][j] = sacc * invd[j];
        }
      }
    }
  }
  T fro = T(0);
  CASSIE_UNROLL
  for (int j = 0; j < 4; j++) {  // column j of L^-1
    CASSIE_UNROLL
    for (int i = 0; i < 4; i++) {
      if (i < j) Li[i][j] = T(0);
      else {
        T sacc = i == j ? T(1) : T(0);
        CASSIE_UNROLL
        for (int k = 0; k < 4; k++)
          if (k >= j && k < i) sacc -= L[i][k] * Li[k][j];
        Li[i][j] = sacc * invd[i];
      }
    }
  }
  CASSIE_UNROLL
  for (int i = 0; i < 4; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < 4; j++) {
      T sacc = T(0);
      CASSIE_UNROLL
      for (int k = 0; k < 4; k++)
        if (k >= i && k >= j) sacc += Li[k][i] * Li[k][j];
      P[i][j] = sacc;
      fro += sacc * sacc;
    }
  }
  // all eigenvalues > tol  <=>  certified by  1/||A^-1||_F > tol  (margin 2x for rounding)
  if (!ok || !(fro * (T(2) * tol) * (T(2) * tol) < T(1))) sym4_pinv_jacobi(A, tol, P);
}

template <typename T>
CASSIE_HD void pinv13x6_apply(T B[kNV][kNU], T tol, const T rhs[kNV], T u[kNU]) {
  // modified Gram-Schmidt QR, B = Q R (Q overwrites a copy of B)
  T Q[kNV][kNU], R[kNU][kNU], y[kNU], rinv[kNU];
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) {
    CASSIE_UNROLL
    for (int j = 0; j < kNU; j++) Q[i][j] = B[i][j];
  }
  bool ok = true;
  CASSIE_UNROLL
  for (int j = 0; j < kNU; j++) {
    T nn = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) nn += Q[i][j] * Q[i][j];
    ok = ok && (nn > T(1e-30));
    const T rjj = Num<T>::sqrt_(nn > T(1e-30) ? nn : T(1));
    const T inv = T(1) / rjj;
    rinv[j] = inv;
    R[j][j] = rjj;
    T d = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) { Q[i][j] *= inv; d += Q[i][j] * rhs[i]; }
    y[j] = d;
    CASSIE_UNROLL
    for (int k = 0; k < kNU; k++) {
      if (k > j) {
        T rjk = T(0);
        CASSIE_UNROLL
        for (int i = 0; i < kNV; i++) rjk += Q[i][j] * Q[i][k];
        R[j][k] = rjk;
        CASSIE_UNROLL
        for (int i = 0; i < kNV; i++) Q[i][k] -= rjk * Q[i][j];
      }
    }
  }
  // R^-1 (upper triangular) for the certificate and the solve
  T Ri[kNU][kNU], fro = T(0);
  CASSIE_UNROLL
  for (int j = kNU - 1; j >= 0; j--) {
    CASSIE_UNROLL
    for (int i = kNU - 1; i >= 0; i--) {
      if (i > j) Ri[i][j] = T(0);
      else {
        T sacc = i == j ? T(1) : T(0);
        CASSIE_UNROLL
        for (int k = 0; k < kNU; k++)
          if (k > i && k <= j) sacc -= R[i][k] * Ri[k][j];
        Ri[i][j] = sacc * rinv[i];
      }
      fro += Ri[i][j] * Ri[i][j];
    }
  }
  if (!ok || !(fro * (T(2) * tol) * (T(2) * tol) < T(1))) { pinv13x6_apply_jacobi(B, tol, rhs, u); return; }
  CASSIE_UNROLL
  for (int i = 0; i < kNU; i++) {
    T sacc = T(0);
    CASSIE_UNROLL
    for (int k = 0; k < kNU; k++)
      if (k >= i) sacc += Ri[i][k] * y[k];
    u[i] = sacc;
  }
}

// ---------------------------------------------------------------------------------------
// What DynamicState::UpdateDynamicState gathers (DynamicState.cpp:45-91) plus the loop-closure
// projector shared by StepJacobian and RunPTSC (Cassie2d.cpp:132-137 == OSC_RBDL.cpp:169-174),
// kept in factored form:  Nc x = x - T1 (JH x),  gamma = T1 JdQd,  with
//   JH = Jeq M^-1 (4x13),  T1 = Jeq^T pinv(JH Jeq^T, 1e-3) (13x4).
template <typename T>
struct CtrlDyn {
  T LD[kNV][kNV], Dinv[kNV];   // M = L^T D L of the controller model (sparse pattern)
  T bias[kNV];                 // C + G + D qd   (DynamicState.cpp:49-52)
  T Jeq[4][8];                 // rows: L x, L z, R x, R z in J8 layout
  T JH[4][kNV], T1[kNV][4], gamma[kNV];
  T sjd[4];                    // pinv(Jeq M^-1 Jeq^T) JeqdotQdot
};

template <typename T>
CASSIE_HD void apply_Nc(const CtrlDyn<T>& d, T x[kNV]) {
  T y[4];
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {
    T s = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) s += d.JH[r][i] * x[i];
    y[r] = s;
  }
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) {
    CASSIE_UNROLL
    for (int r = 0; r < 4; r++) x[i] -= d.T1[i][r] * y[r];
  }
}

template <typename T>
CASSIE_HD void ctrl_dynamics(const PlanarModel<T>& m, const Kin<T>& k, const T* qd, CtrlDyn<T>& d) {
  mass_matrix(m, k, d.LD);
  bias_forces(m, k, d.bias);
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) d.bias[i] += m.damping[i] * qd[i];
  factor(d.LD, d.Dinv);
  PivotAcc<T> pa;
  pivot_accelerations(k, pa);
  T jd[4];
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    T ax, az, bx, bz;
    rot(k.c[L][kRod], k.s[L][kRod], m.eq_a1[L][0], m.eq_a1[L][1], ax, az);
    rot(k.c[L][kTarsus], k.s[L][kTarsus], m.eq_a2[L][0], m.eq_a2[L][1], bx, bz);
    T J1x[8], J1z[8], J2x[8], J2z[8];
    point_jac(m, k, L, kRod, ax, az, J1x, J1z);
    point_jac(m, k, L, kTarsus, bx, bz, J2x, J2z);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) { d.Jeq[2 * L][c] = J1x[c] - J2x[c]; d.Jeq[2 * L + 1][c] = J1z[c] - J2z[c]; }
    const T w1 = k.w[L][kRod] * k.w[L][kRod], w2 = k.w[L][kTarsus] * k.w[L][kTarsus];
    jd[2 * L] = (pa.ax[L][kRod] - w1 * ax) - (pa.ax[L][kTarsus] - w2 * bx);
    jd[2 * L + 1] = (pa.az[L][kRod] - w1 * az) - (pa.az[L][kTarsus] - w2 * bz);
  }
  T S[4][4], P[4][4];
  CASSIE_ROLL
  for (int r = 0; r < 4; r++) {
    expand_row(d.Jeq[r], r / 2, d.JH[r]);
    solve(d.LD, d.Dinv, d.JH[r]);
  }
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) S[r][c] = dot8_dense(d.Jeq[c], c / 2, d.JH[r]);
  }
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {  // symmetrise (rounding) before the Jacobi sweeps
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++)
      if (c > r) { const T v = T(0.5) * (S[r][c] + S[c][r]); S[r][c] = v; S[c][r] = v; }
  }
  sym4_pinv(S, T(1e-3), P);
  CASSIE_UNROLL
  for (int c = 0; c < 4; c++) {
    T col[4];
    CASSIE_UNROLL
    for (int r = 0; r < 4; r++) col[r] = P[r][c];
    // T1[:,c] = sum_r Jeq[r]^T P[r][c]
    T acc[kNV];
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) acc[i] = T(0);
    CASSIE_UNROLL
    for (int r = 0; r < 4; r++) {
      T e[kNV];
      expand_row(d.Jeq[r], r / 2, e);
      CASSIE_UNROLL
      for (int i = 0; i < kNV; i++) acc[i] += e[i] * col[r];
    }
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) d.T1[i][c] = acc[i];
  }
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) {
    T s = T(0);
    CASSIE_UNROLL
    for (int r = 0; r < 4; r++) s += d.T1[i][r] * jd[r];
    d.gamma[i] = s;
  }
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {
    T s = T(0);
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) s += P[r][c] * jd[c];
    d.sjd[r] = s;
  }
}

// Cassie2d::StepJacobian control law (Cassie2d.cpp:119-165):
//   u = pinv(Nc Bt) (Nc bias + gamma - Nc Jc6^T f),  f per foot = (Fx, Fz, My).
template <typename T>
CASSIE_HD void jacobian_control(const PlanarModel<T>& m, const Kin<T>& k, const T* qd, const T f[6], T u[kNU]) {
  CtrlDyn<T> d;
  ctrl_dynamics(m, k, qd, d);
  // Jc6^T f: mean of the front/rear site Jacobians of each foot; angular row about world +y
  T x[kNV];
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) x[i] = d.bias[i];
  CASSIE_UNROLL
  for (int L = 0; L < 2; L++) {
    const T Fx = f[3 * L], Fz = f[3 * L + 1], My = f[3 * L + 2];
    T Jx[8], Jz[8], Jr[8];
    // mean point of sites (2+2L, 3+2L) on the toe link
    T rx, rz;
    const int s0 = 2 + 2 * L, s1 = 3 + 2 * L;
    rot(k.c[L][kToe], k.s[L][kToe], T(0.5) * (m.site_off[s0][0] + m.site_off[s1][0]),
        T(0.5) * (m.site_off[s0][1] + m.site_off[s1][1]), rx, rz);
    point_jac(m, k, L, kToe, rx, rz, Jx, Jz);
    Jr[0] = T(0); Jr[1] = T(0); Jr[2] = T(1);
    CASSIE_UNROLL
    for (int b = 0; b < kLegLinks; b++) Jr[3 + b] = link_anc(kToe, b) ? m.sgn[L][b] : T(0);
    T J8[8], e[kNV];
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) J8[c] = Jx[c] * Fx + Jz[c] * Fz + Jr[c] * My;
    expand_row(J8, L, e);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) x[i] -= e[i];
  }
  apply_Nc(d, x);
  CASSIE_UNROLL
  for (int i = 0; i < kNV; i++) x[i] += d.gamma[i];
  T B[kNV][kNU];
  CASSIE_ROLL
  for (int a = 0; a < kNU; a++) {
    T col[kNV];
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) col[i] = m.act_dof[a] == i ? m.act_gear[a] : T(0);
    apply_Nc(d, col);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) B[i][a] = col[i];
  }
  pinv13x6_apply(B, T(1e-4), x, u);
}

}  // namespace cassie
