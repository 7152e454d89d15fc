// Operational-space QP controller: the reference's OSC_RBDL::RunPTSC + SolveQP
// (CassieRL/cassierl src/OSC_RBDL.cpp:114-291, called from Cassie2d::StepOsc,
// src/Cassie2d/Cassie2d.cpp:179-209), rebuilt as a per-env device routine.
//
// The reference assembles a 39-variable / 45-row QP  x = [qdd(13); u(6); beta(20)]  and hands it
// to qpOASES 3.2.1 (un-vendored).  The same optimum is reached here on a reduced, box-constrained
// problem:
//   * qdd is eliminated through the 13 dynamics equalities (OSC_RBDL.cpp:180-184):
//       qdd = M^-1 (Nc Bt u + Nc Jc^T f + ce),   ce = -Nc bias - gamma;
//   * planar: the y components of the contact forces have zero Jacobian rows, only their 1e-4
//     cost remains, so they are 0 at the optimum;
//   * beta = (x-, x+, z) >= 0 with  x+- <= mu z  (OSC_RBDL.cpp:41-71, mu = 0.5,
//     RobotInterface.h:64) and cost 1e-4/2 |beta|^2 (OSC_RBDL.cpp:188-201) is equivalent to a
//     force (fx, fz) in the cone |fx| <= mu fz with cost 1e-4/2 (fx^2 + fz^2); writing the cone by
//     its two edge generators  f = l1 (mu, 1) + l2 (-mu, 1),  l >= 0  leaves only simple bounds.
// Result: min 1/2 z'Gz + g'z over z = [u(6); l(8)], u in the motor limits, l >= 0; solved by a
// block-principal-pivoting method with a masked Cholesky factorisation, in double precision in every
// build (G = 2 E'WE + reg has a condition number ~1e10; forming it in fp32 would lose the 1e-4
// regulariser that makes the optimum unique).
#pragma once
#include "controllers.cuh"

namespace cassie {

struct OscStats { int iters; int status; };  // status 0 = optimal, 1 = iteration cap, 2 = factorisation failed

constexpr int kQpN = 14;
constexpr int kQpTasks = 11;

// weights: OSC_RBDL.h:92-98 / OSC_RBDL.cpp:32-38 with all four contacts desired (Cassie2d.cpp:199)
constexpr double kOscWCom = 5.0, kOscWStance = 10.0, kOscWRest = 0.1, kOscWForce = 1e-4, kOscMu = 0.5;

// Box-constrained strictly convex QP   min 1/2 z'Gz + g'z,  lo <= z <= hi   by block principal
// pivoting (Judice & Pires): every iteration solves the KKT system of the current partition
// (free / at-lower / at-upper) with a masked Cholesky factorisation and exchanges ALL variables
// that violate primal or dual feasibility; when the number of violations stops decreasing it makes
// single exchanges (most violated first, then Murty's least-index rule, which guarantees termination).  `at_lo` / `at_hi` carry the partition
// in and out (warm start across steps, like the qpOASES hot start, OSC_RBDL.cpp:278).
CASSIE_HD constexpr int qtri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
constexpr int kQpTri = kQpN * (kQpN + 1) / 2;
constexpr int kQpGreedyIters = 60;
constexpr int kOscWsDoubles = (kQpTasks + 1) * kQpN + 2 * kQpTri;   // E, G, L

// G and the Cholesky factor are packed lower triangles (105 entries): half the thread-local lines.
// GA / VA / LA: anything indexable with [] that yields double (plain pointers in the thread-per-env engine, strided
// shared-memory views in the quad engine, quad_rt.cuh SV<double>)
template <typename GA, typename VA, typename ZA, typename LA>
CASSIE_HD void box_qp_solve_on(const GA G, const VA g, const VA lo, const VA hi, ZA z, LA L, unsigned& at_lo, unsigned& at_hi,
                               int max_iter, OscStats* st) {
  double grad[kQpN], invd[kQpN];  // invd = 1 / L_ii: one division per pivot instead of one per entry
  double gscale = 1.0;
  for (int i = 0; i < kQpN; i++) gscale = fmax(gscale, fabs(g[i]));
  const double dtol = 1e-12 * gscale;
  int it = 0, status = 1, best = kQpN + 1;
  for (; it < max_iter; it++) {
    const unsigned fixed = at_lo | at_hi;
    for (int i = 0; i < kQpN; i++) z[i] = ((at_lo >> i) & 1u) ? lo[i] : (((at_hi >> i) & 1u) ? hi[i] : 0.0);
    // rhs_F = -(g_F + G_FB z_B); masked Cholesky: pinned variables become identity rows/columns
    // KKT system of the partition on the COMPACTED free set (nf <= 14 variables): packed right-looking
    // Cholesky in tight rolled loops.  Right-looking makes the innermost updates independent of each other
    // (they pipeline), the loops stay resident in the instruction cache, and pinned variables cost nothing.
    // (A left-looking rolled version was one dependent load/DFMA chain, a fully unrolled masked version
    // 6 k straight-line instructions at ~10 cycles each: profiles/r1m.)
    int idx[kQpN], nf = 0;
    for (int i = 0; i < kQpN; i++)
      if (!((fixed >> i) & 1u)) idx[nf++] = i;
    for (int a = 0; a < nf; a++) {
      const int ia = idx[a];
      double rhs = -g[ia];
      for (int j = 0; j < kQpN; j++)
        if ((fixed >> j) & 1u) rhs -= G[qtri(ia, j)] * z[j];
      grad[a] = rhs;
      for (int b = 0; b <= a; b++) L[qtri(a, b)] = G[qtri(ia, idx[b])];
    }
    bool ok = true;
    for (int j = 0; j < nf; j++) {
      double djj = L[qtri(j, j)];
      if (!(djj > 0.0)) { ok = false; djj = 1.0; }
      const double d = sqrt(djj), inv = 1.0 / d;
      L[qtri(j, j)] = d;
      invd[j] = inv;
      for (int i = j + 1; i < nf; i++) L[qtri(i, j)] *= inv;
      for (int i = j + 1; i < nf; i++) {
        const double lij = L[qtri(i, j)];
        const int row = i * (i + 1) / 2;
        for (int k = j + 1; k <= i; k++) L[row + k] -= lij * L[qtri(k, j)];
      }
    }
    if (!ok) { status = 2; break; }
    for (int a = 0; a < nf; a++) {
      double sacc = grad[a];
      const int row = a * (a + 1) / 2;
      for (int k = 0; k < a; k++) sacc -= L[row + k] * grad[k];
      grad[a] = sacc * invd[a];
    }
    for (int a = nf - 1; a >= 0; a--) {
      double sacc = grad[a];
      for (int k = a + 1; k < nf; k++) sacc -= L[qtri(k, a)] * grad[k];
      grad[a] = sacc * invd[a];
      z[idx[a]] = grad[a];
    }
    // violations: free variables outside their bounds, pinned variables with a wrong-sign multiplier
    unsigned viol = 0u;
    int nviol = 0, last = -1, worst = -1;
    double worst_mag = -1.0;
    // cond(G) ~ 1e10: a degenerate variable (true value 0, true multiplier 0) comes out as +-1e-6 of the
    // solution scale, so feasibility is judged with a tolerance relative to that scale -- otherwise it
    // flips between "free" and "pinned" forever
    double zmax = 1.0;
    for (int i = 0; i < kQpN; i++) zmax = fmax(zmax, fabs(z[i]));
    const double ptol = 1e-8 * zmax;
    for (int i = 0; i < kQpN; i++) {
      double v, mag;  // violation, and the violation in units of its own scale
      bool bad;
      if ((fixed >> i) & 1u) {
        double s = g[i];
        for (int j = 0; j < kQpN; j++) s += G[qtri(i, j)] * z[j];
        v = ((at_lo >> i) & 1u) ? -s : s;
        bad = v > dtol;
        mag = v / gscale;
      } else {
        v = fmax(lo[i] - z[i], z[i] - hi[i]);
        bad = v > ptol;
        mag = v / zmax;
      }
      if (bad) {
        viol |= 1u << i; nviol++; last = i;
        if (mag > worst_mag) { worst_mag = mag; worst = i; }
      }
    }
    if (nviol == 0) { status = 0; it++; break; }
    // Exchange ALL violators only while that keeps shrinking the violation count; otherwise exchange ONE: the
    // most violated variable (cold starts after a new random action: 4.3 iterations on average, 15 at most, where
    // ten more block attempts followed by Murty's least-index rule took 12.5 / 126 -- and a warp waits for its
    // slowest lane), and after kQpGreedyIters iterations Murty's rule, which cannot cycle.
    if (nviol < best) best = nviol;
    else viol = 1u << (it < kQpGreedyIters ? worst : last);
    for (int i = 0; i < kQpN; i++) {
      if (!((viol >> i) & 1u)) continue;
      if ((fixed >> i) & 1u) { at_lo &= ~(1u << i); at_hi &= ~(1u << i); }
      else if (z[i] < lo[i]) at_lo |= 1u << i;
      else at_hi |= 1u << i;
    }
  }
  for (int i = 0; i < kQpN; i++) z[i] = z[i] < lo[i] ? lo[i] : (z[i] > hi[i] ? hi[i] : z[i]);
  if (st) { st->iters = it; st->status = status; }
}
CASSIE_HD void box_qp_solve(const double G[kQpTri], const double g[kQpN], const double lo[kQpN],
                            const double hi[kQpN], double z[kQpN], unsigned& at_lo, unsigned& at_hi, int max_iter,
                            OscStats* st, double* Lws = nullptr) {
#if defined(__CUDA_ARCH__) && !defined(CASSIE_NO_OSC_OVERLAY)
  box_qp_solve_on(G, g, lo, hi, z, Lws, at_lo, at_hi, max_iter, st);
#else
  double Lloc[kQpTri];
  box_qp_solve_on(G, g, lo, hi, z, Lws ? Lws : Lloc, at_lo, at_hi, max_iter, st);
#endif
}

// OSC_RBDL::RunPTSC for the planar model.  act = ControllerOsc in memory order (RobotInterface.h:23-28):
// body_xdd[2], left_xdd[2], right_xdd[2], pitch_add  (x, z pairs; Cassie2d.cpp:185-193).
template <typename T>
CASSIE_HD void osc_control(const PlanarModel<T>& m, const Kin<T>& k, const T* qd, const T act[7], T u[kNU], OscStats* st,
                            unsigned* qp_set = nullptr, double* ws = nullptr) {
#if defined(__CUDA_ARCH__) && !defined(CASSIE_NO_OSC_OVERLAY)
  CtrlDyn<T>& d = *reinterpret_cast<CtrlDyn<T>*>(ws + kOscWsDoubles);
#else
  CtrlDyn<T> d;
#endif
  ctrl_dynamics(m, k, qd, d);
  PivotAcc<T> pa;
  pivot_accelerations(k, pa);
  // ---- task rows (OSC_RBDL.cpp:123-144): site Jacobians in J8 layout, leg of each row, Jdot*qd - xdd*
  T A[kQpTasks + 1][8], e0[kQpTasks + 1];
  int aleg[kQpTasks + 1];
  double W[kQpTasks];
  {
    T rx, rz;
    rot(k.c0, k.s0, m.site_off[1][0], m.site_off[1][1], rx, rz);
    point_jac(m, k, 0, -1, rx, rz, A[0], A[1]);
    const T w2 = k.w0 * k.w0;
    e0[0] = -w2 * rx - act[0];
    e0[1] = -w2 * rz - act[1];
    aleg[0] = 0; aleg[1] = 0;
    W[0] = kOscWCom; W[1] = kOscWCom;
  }
  CASSIE_UNROLL
  for (int s = 0; s < 4; s++) {
    const int L = s / 2, site = 2 + s;
    T rx, rz;
    rot(k.c[L][kToe], k.s[L][kToe], m.site_off[site][0], m.site_off[site][1], rx, rz);
    point_jac(m, k, L, kToe, rx, rz, A[2 + 2 * s], A[3 + 2 * s]);
    const T w2 = k.w[L][kToe] * k.w[L][kToe];
    e0[2 + 2 * s] = pa.ax[L][kToe] - w2 * rx - act[2 + 2 * L];
    e0[3 + 2 * s] = pa.az[L][kToe] - w2 * rz - act[3 + 2 * L];
    aleg[2 + 2 * s] = L; aleg[3 + 2 * s] = L;
    W[2 + 2 * s] = kOscWStance; W[3 + 2 * s] = kOscWStance;
  }
  CASSIE_UNROLL
  for (int c = 0; c < 8; c++) A[10][c] = c == 2 ? T(1) : T(0);   // AddQDDIdx(2): pitch (Cassie2d.cpp:41)
  e0[10] = -act[6];
  aleg[10] = 0;
  W[10] = kOscWRest;
  // ---- E = A Mc^-1 B row by row, with Mc^-1 = M^-1 Nc (symmetric: M^-1 - M^-1 Jeq' S^+ Jeq M^-1) and
  // B = [Bt, Jc' T]:  Z_r = Mc^-1 A_r' costs one projection + one tree-sparse solve per task; the u
  // columns of E are then single entries of Z_r (Bt is a gear selector) and the contact columns are
  // dots with the site Jacobians, which ARE task rows 2..9.  P = Mc^-1 B is never formed; E (11 x 14) is
  // stored once and  G = 2 E'WE,  g = 2 E'W r0  (OSC_RBDL.cpp:186-203) are then built entry by entry.
  //   r0_r = Jdot qd_r - xdd*_r + A_r p0,   A_r p0 = -Z_r . bias - (JH A_r') . (S^+ JdQd)
  // E, G, L (and CtrlDyn above) live in caller-provided scratch `ws`: the kernels pass the storage of the physics
  // step's constraint rows, which is dead while the controller runs, so that the two phases touch the same
  // thread-local lines instead of two disjoint sets (the controller is bound by L1 / L2 hits, DESIGN.md 5:
  // E/G/L +1 %, CtrlDyn another +4 %; the task Jacobians A on top of that lost 2 % again)
  double r0v[kQpTasks + 1];
#if defined(__CUDA_ARCH__) && !defined(CASSIE_NO_OSC_OVERLAY)
  double (*const E)[kQpN] = reinterpret_cast<double (*)[kQpN]>(ws);
  double* const G = ws + (kQpTasks + 1) * kQpN;
  double* const Lws = G + kQpTri;
#else
  double Eloc[kQpTasks + 1][kQpN], Gloc[kQpTri];
  double (*const E)[kQpN] = ws ? reinterpret_cast<double (*)[kQpN]>(ws) : Eloc;
  double* const G = ws ? ws + (kQpTasks + 1) * kQpN : Gloc;
  double* const Lws = ws ? G + kQpTri : nullptr;
#endif
  CASSIE_UNROLL
  for (int c = 0; c < 8; c++) A[kQpTasks][c] = T(0);   // padding task: the loop below handles two tasks per
  e0[kQpTasks] = T(0); aleg[kQpTasks] = 0;              // iteration so that every load of JH / T1 / LD serves both
  _Pragma("unroll 2")
  for (int r = 0; r < kQpTasks + 1; r++) {
    T x[kNV], y[4];
    expand_row(A[r], aleg[r], x);
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) {
      T sacc = T(0);
      CASSIE_UNROLL
      for (int i = 0; i < kNV; i++) sacc += d.JH[c][i] * x[i];
      y[c] = sacc;
    }
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) {
      CASSIE_UNROLL
      for (int c = 0; c < 4; c++) x[i] -= d.T1[i][c] * y[c];
    }
    solve(d.LD, d.Dinv, x);  // x = Z_r
    {  // u columns: gear * Z_r[actuated dof]; dof-major so that every x[i] is touched once (x lives in spill slots)
      T v[kNU];
      CASSIE_UNROLL
      for (int a = 0; a < kNU; a++) v[a] = T(0);
      CASSIE_UNROLL
      for (int i = 3; i < kNV; i++) {
        const T xi = x[i];
        CASSIE_UNROLL
        for (int a = 0; a < kNU; a++) v[a] = m.act_dof[a] == i ? xi : v[a];
      }
      CASSIE_UNROLL
      for (int a = 0; a < kNU; a++) E[r][a] = (double)(m.act_gear[a] * v[a]);
    }
    CASSIE_UNROLL
    for (int s = 0; s < 4; s++) {
      const double ex = (double)dot8_dense(A[2 + 2 * s], s / 2, x), ez = (double)dot8_dense(A[3 + 2 * s], s / 2, x);
      E[r][kNU + 2 * s] = kOscMu * ex + ez;
      E[r][kNU + 2 * s + 1] = -kOscMu * ex + ez;
    }
    T zb = T(0), ys = T(0);
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) zb += x[i] * d.bias[i];
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) ys += y[c] * d.sjd[c];
    r0v[r] = (double)(e0[r] - zb - ys);
  }
  // G = 2 E'WE (packed lower triangle, one store per entry), g = 2 E'W r0.  Two rows of G per pass: the
  // weighted columns i, i+1 of E sit in registers, every column j <= i+1 is loaded once and feeds four
  // independent accumulation chains.
#if defined(__CUDA_ARCH__) && !defined(CASSIE_NO_OSC_OVERLAY)   // the QP's vectors too: +3.5 %
  double* const g = reinterpret_cast<double*>(reinterpret_cast<char*>(ws + kOscWsDoubles) + sizeof(CtrlDyn<T>));
  double* const lo = g + kQpN; double* const hi = lo + kQpN; double* const z = hi + kQpN;
#else
  double g[kQpN], lo[kQpN], hi[kQpN], z[kQpN];
#endif
  CASSIE_ROLL
  for (int i = 0; i < kQpN; i += 2) {
    double c0[kQpTasks], c1[kQpTasks];
    double g0 = 0.0, g1 = 0.0;
    CASSIE_UNROLL
    for (int r = 0; r < kQpTasks; r++) {
      const double w2 = 2.0 * W[r];
      c0[r] = w2 * E[r][i]; c1[r] = w2 * E[r][i + 1];
      g0 += c0[r] * r0v[r]; g1 += c1[r] * r0v[r];
    }
    g[i] = g0; g[i + 1] = g1;
    CASSIE_ROLL
    for (int j = 0; j <= i + 1; j++) {
      double s0a = 0.0, s0b = 0.0, s1a = 0.0, s1b = 0.0;
      CASSIE_UNROLL
      for (int r = 0; r < kQpTasks; r++) {
        const double ej = E[r][j];
        if (r & 1) { s0b += c0[r] * ej; s1b += c1[r] * ej; }
        else { s0a += c0[r] * ej; s1a += c1[r] * ej; }
      }
      if (j <= i) G[qtri(i, j)] = s0a + s0b;
      G[qtri(i + 1, j)] = s1a + s1b;
    }
  }
  for (int s = 0; s < 4; s++) {
    const int a = kNU + 2 * s, b = a + 1;
    G[qtri(a, a)] += kOscWForce * (kOscMu * kOscMu + 1.0);
    G[qtri(b, b)] += kOscWForce * (kOscMu * kOscMu + 1.0);
    G[qtri(b, a)] += kOscWForce * (1.0 - kOscMu * kOscMu);
  }
  for (int a = 0; a < kNU; a++) { lo[a] = (double)m.act_lo[a]; hi[a] = (double)m.act_hi[a]; }
  for (int i = kNU; i < kQpN; i++) { lo[i] = 0.0; hi[i] = 1e30; }
  unsigned at_lo = qp_set ? (*qp_set & 0x3fffu) : 0u, at_hi = qp_set ? ((*qp_set >> 14) & 0x3fu) : 0u;
  box_qp_solve(G, g, lo, hi, z, at_lo, at_hi, 300, st, Lws);
  if (qp_set) *qp_set = at_lo | (at_hi << 14);
  CASSIE_UNROLL
  for (int a = 0; a < kNU; a++) u[a] = (T)z[a];
}

}  // namespace cassie
