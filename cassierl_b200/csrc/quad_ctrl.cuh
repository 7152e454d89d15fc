// Quad engine, controller side: what the reference computes with RBDL + Eigen + qpOASES before every mj_step
// (CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:86-237, src/DynamicState.cpp:45-91, src/DynamicModel.cpp:237-367,
// src/OSC_RBDL.cpp:114-291), by the four lanes of a quad.  Same mathematics as controllers.cuh / osc_qp.cuh (the
// thread-per-env versions), in double precision in every build (profiles/r2_variants.txt: the task
// matrix E, the loop-closure projector and the QP need it; a float E alone costs 1.2e-5 on the single-step bar):
//   * controller-model kinematics, mass matrix, bias and factorisation are leg-local (quad_engine.cuh);
//   * JH = Jeq M^-1 and the twelve task rows Z_r = Mc^-1 A_r' are multi-RHS leg-split solves: half h of the quad
//     takes the right-hand sides that belong to leg h, lane (L, h) ends up with the base part and the leg-L part of
//     the result and fills the E entries of ITS leg's actuators and contact generators;
//   * G = 2 E'WE is built from shared memory two rows at a time per lane; the 14-variable box QP itself (block
//     principal pivoting, osc_qp.cuh) runs on lane 0 of the quad out of shared memory.
#pragma once
#include "quad_engine.cuh"
#include "osc_qp.cuh"

namespace cassie {
namespace quad {

typedef double TC;   // controller precision

// scratch of the controller (doubles); shares storage with PhysLayout (the two phases of a step alternate)
struct CtrlLayout {
  static constexpr int ld = 0;                       // factor of the controller model's M
  static constexpr int jeq = ld + kLdSize;           // [2 L + xz][8]
  static constexpr int jd = jeq + 32;                // JeqdotQdot [2 L + xz]
  static constexpr int jh = jd + 4;                  // Jeq M^-1: [2 L + xz][13]
  static constexpr int sp = jh + 52;                 // Jeq M^-1 Jeq' [4][4]
  static constexpr int task = sp + 16;               // site Jacobians [(2 L + site) * 2 + xz][8]
  static constexpr int e0 = task + 64;               // Jdot qd - xdd* per task row [12]
  static constexpr int r0 = e0 + 12;                 // task residual at z = 0 [12]
  static constexpr int E = r0 + 12;                  // [12][14]
  static constexpr int end = E + kQpTri + 3 + 5 * kQpN > E + 12 * kQpN ? E + kQpTri + 3 + 5 * kQpN : E + 12 * kQpN;
  // QP overlay: everything below `e0` is dead once the task loop is over
  // QP overlay: G (full symmetric 14 x 14) and g are written while E / r0 are read; the factor and the QP's vectors
  // reuse E once G is complete
  static constexpr int G = 0, g = G + kQpN * kQpN, Lw = E, qpv = Lw + kQpTri + 3;
  static_assert(g + kQpN <= e0, "G and g must not touch e0 / r0 / E while G is being built");
  // Jacobian mode: B = Nc Bt (13 x 6) and the right-hand side, gathered for lane 0
  static constexpr int jacB = E, jacRhs = jacB + kNV * kNU;
};

// world position / velocity of a point fixed to the toe link (A = kToe) or the pelvis (A = -1): controllers.cuh site_point
template <int A>
QUAD_FN void leg_site_point(const PlanarModel<TC>& m, const LegKin<TC>& k, const V8<TC>& q, TC ox, TC oz, TC out[4]) {
  TC rx, rz, px = TC(0), pz = TC(0), w = k.w0, vx = k.v0x, vz = k.v0z;
  if (A < 0) rot(k.c0, k.s0, ox, oz, rx, rz);
  else { rot(k.c[A], k.s[A], ox, oz, rx, rz); px = k.px[A]; pz = k.pz[A]; w = k.w[A]; vx = k.vx[A]; vz = k.vz[A]; }
  out[0] = (q.b[0] - m.pel_ref[0] + m.pel_org[0]) + px + rx;
  out[1] = (q.b[1] - m.pel_ref[1] + m.pel_org[1]) + pz + rz;
  out[2] = vx + w * rz;
  out[3] = vz - w * rx;
}

// Controller-model kinematics of this lane's leg + the lagged operational-space state (GetOperationalSpaceState reads
// the RBDL state of the START of the last Step*, Cassie2d.cpp:88,98,121,182 vs :223): written to St[op] as
// body[4], left[4], right[4] = (x, z, xd, zd).  Sites: 1 body_center, 2/3 left front/rear, 4/5 right (Cassie2d.cpp:34-36).
template <typename T>
QUAD_FN void quad_ctrl_kin(const PlanarModel<TC>& m, const Lane ln, SV<T> St, V8<TC>& q, V8<TC>& qd, LegKin<TC>& k, bool write_op) {
  typedef StateLayout X;
  const int L = ln.L;
  CASSIE_UNROLL
  for (int b = 0; b < 3; b++) { q.b[b] = (TC)St[X::q + b]; qd.b[b] = (TC)St[X::qd + b]; }
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) { q.l[a] = (TC)St[X::q + 3 + 5 * L + a]; qd.l[a] = (TC)St[X::qd + 3 + 5 * L + a]; }
  leg_fk_positions(m, L, q, k);
  leg_fk_velocities(m, L, qd, k);
  if (write_op) {
    TC a[4], b[4];
    leg_site_point<kToe>(m, k, q, m.site_off[2 + 2 * L][0], m.site_off[2 + 2 * L][1], a);
    leg_site_point<kToe>(m, k, q, m.site_off[3 + 2 * L][0], m.site_off[3 + 2 * L][1], b);
    if (ln.h == 0) {
      CASSIE_UNROLL
      for (int i = 0; i < 4; i++) St[X::op + 4 + 4 * L + i] = (T)((a[i] + b[i]) / TC(2));
    }
    if (ln.ql == 1) {
      leg_site_point<-1>(m, k, q, m.site_off[1][0], m.site_off[1][1], a);
      CASSIE_UNROLL
      for (int i = 0; i < 4; i++) St[X::op + i] = (T)a[i];
    }
  }
}

// StateOperationalSpace in memory order (RobotInterface.h:47-50) from the stored op-space state and the CURRENT pitch
// and pitch rate (Cassie2d.cpp:234-235); controllers.cuh op_state_array
template <typename T>
QUAD_FN void quad_op_array(SV<T> St, T o[18]) {
  typedef StateLayout X;
  o[0] = St[X::op + 0]; o[1] = St[X::op + 1]; o[2] = St[X::q + 2];
  o[3] = St[X::op + 2]; o[4] = St[X::op + 3]; o[5] = St[X::qd + 2];
  o[6] = St[X::op + 4]; o[7] = St[X::op + 5]; o[8] = T(0);
  o[9] = St[X::op + 6]; o[10] = St[X::op + 7]; o[11] = T(0);
  o[12] = St[X::op + 8]; o[13] = St[X::op + 9]; o[14] = T(0);
  o[15] = St[X::op + 10]; o[16] = St[X::op + 11]; o[17] = T(0);
}

// What DynamicState::UpdateDynamicState gathers (DynamicState.cpp:45-91) plus the loop-closure projector shared by
// StepJacobian and RunPTSC (Cassie2d.cpp:132-137 == OSC_RBDL.cpp:169-174) in factored form (controllers.cuh CtrlDyn):
//   Nc x = x - Jeq' P (JH x),  JH = Jeq M^-1,  P = pinv(JH Jeq', 1e-3).
// Leaves ld, jeq, jd, jh in C; bias (C + G + D qd), P and P jd in registers of every lane.  Also returns the pivot
// accelerations of this leg (Jdot qd of link-fixed points, DynamicModel.cpp:341-344).
struct QuadCtrlDyn {
  V8<TC> bias;
  TC P[4][4], sjd[4];
  TC pax[kLegLinks], paz[kLegLinks];
};
QUAD_FN void quad_ctrl_dynamics(const PlanarModel<TC>& m, const Lane ln, const LegKin<TC>& k, const V8<TC>& qd, SV<TC> C,
                                QuadCtrlDyn& d) {
  typedef CtrlLayout Y;
  const int L = ln.L, h = ln.h;
  {
    LegM<TC> M;
    leg_mass_matrix(m, L, k, M);
    leg_factor(M);
    if (h == 0) store_factor(C.at(Y::ld), L, M, L == 0);
  }
  leg_bias_forces(m, L, k, d.bias);
  CASSIE_UNROLL
  for (int b = 0; b < 3; b++) d.bias.b[b] += m.damping[b] * qd.b[b];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) d.bias.l[a] += m.damping[3 + 5 * L + a] * qd.l[a];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    const int p = link_parent(a);
    TC apx, apz, wp;
    if (p < 0) { apx = TC(0); apz = TC(0); wp = k.w0; }
    else { apx = d.pax[p]; apz = d.paz[p]; wp = k.w[p]; }
    const TC w2 = wp * wp;
    d.pax[a] = apx - w2 * k.dx[a];
    d.paz[a] = apz - w2 * k.dz[a];
  }
  {  // loop-closure row (L, h): x row on half 0, z row on half 1
    TC ax, az, bx, bz;
    rot(k.c[kRod], k.s[kRod], m.eq_a1[L][0], m.eq_a1[L][1], ax, az);
    rot(k.c[kTarsus], k.s[kTarsus], m.eq_a2[L][0], m.eq_a2[L][1], bx, bz);
    TC J1x[8], J1z[8], J2x[8], J2z[8];
    leg_point_jac<kRod>(m, L, k, ax, az, J1x, J1z);
    leg_point_jac<kTarsus>(m, L, k, bx, bz, J2x, J2z);
    const SV<TC> dst = C.at(Y::jeq + (2 * L + h) * 8);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) dst[c] = h ? J1z[c] - J2z[c] : J1x[c] - J2x[c];
    const TC w1 = k.w[kRod] * k.w[kRod], w2 = k.w[kTarsus] * k.w[kTarsus];
    C[Y::jd + 2 * L + h] = h ? (d.paz[kRod] - w1 * az) - (d.paz[kTarsus] - w2 * bz)
                             : (d.pax[kRod] - w1 * ax) - (d.pax[kTarsus] - w2 * bx);
  }
  wsync();
  {  // JH rows of leg h (two right-hand sides at once) and their block of S = JH Jeq'
    V8<TC> x[2];
    CASSIE_UNROLL
    for (int t = 0; t < 2; t++) {
      const SV<TC> src = C.at(Y::jeq + (2 * h + t) * 8);
      CASSIE_UNROLL
      for (int b = 0; b < 3; b++) x[t].b[b] = src[b];
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) x[t].l[a] = (L == h) ? src[3 + a] : TC(0);
    }
    quad_solve<2>(C.at(Y::ld), L, x);
    TC Jown[2][8];
    CASSIE_UNROLL
    for (int s = 0; s < 2; s++) {
      const SV<TC> src = C.at(Y::jeq + (2 * L + s) * 8);
      CASSIE_UNROLL
      for (int c = 0; c < 8; c++) Jown[s][c] = src[c];
    }
    CASSIE_UNROLL
    for (int t = 0; t < 2; t++) {
      const SV<TC> dst = C.at(Y::jh + (2 * h + t) * kNV);
      if (L == 0) {
        CASSIE_UNROLL
        for (int b = 0; b < 3; b++) dst[b] = x[t].b[b];
      }
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) dst[3 + 5 * L + a] = x[t].l[a];
      CASSIE_UNROLL
      for (int s = 0; s < 2; s++) C[Y::sp + (2 * h + t) * 4 + 2 * L + s] = dot8(Jown[s], x[t]);
    }
  }
  wsync();
  TC S[4][4];
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) S[r][c] = C[Y::sp + 4 * r + c];
  }
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++)
      if (c > r) { const TC v = TC(0.5) * (S[r][c] + S[c][r]); S[r][c] = v; S[c][r] = v; }
  }
  sym4_pinv(S, TC(1e-3), d.P);
  CASSIE_UNROLL
  for (int r = 0; r < 4; r++) {
    TC s = TC(0);
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) s += d.P[r][c] * C[Y::jd + c];
    d.sjd[r] = s;
  }
}

// x <- Nc' x in leg-split form is never needed; what both controllers need is  x <- x - Jeq' P (JH x)  for a
// leg-split x whose leg part lives on the lanes of its leg(s).  y = JH x is returned (the OSC residual uses it).
template <int NR>
QUAD_FN void quad_project(const Lane ln, SV<TC> C, const QuadCtrlDyn& d, V8<TC> x[NR], TC y[NR][4]) {
  typedef CtrlLayout Y;
  const int L = ln.L;
  TC w[NR][4];
  CASSIE_UNROLL
  for (int c = 0; c < 4; c++) {
    const SV<TC> jh = C.at(Y::jh + c * kNV);
    TC jl[kLegLinks], jb[3];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) jl[a] = jh[3 + 5 * L + a];
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) jb[b] = jh[b];
    CASSIE_UNROLL
    for (int r = 0; r < NR; r++) {
      TC pl = TC(0);
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) pl += jl[a] * x[r].l[a];
      y[r][c] = (jb[0] * x[r].b[0] + jb[1] * x[r].b[1] + jb[2] * x[r].b[2]) + sum_legs(pl);
    }
  }
  CASSIE_UNROLL
  for (int r = 0; r < NR; r++) {
    CASSIE_UNROLL
    for (int q = 0; q < 4; q++) {
      TC s = TC(0);
      CASSIE_UNROLL
      for (int c = 0; c < 4; c++) s += d.P[q][c] * y[r][c];
      w[r][q] = s;
    }
  }
  CASSIE_UNROLL
  for (int c = 0; c < 4; c++) {
    const SV<TC> je = C.at(Y::jeq + c * 8);
    TC jb[3];
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) jb[b] = je[b];
    CASSIE_UNROLL
    for (int r = 0; r < NR; r++) {
      CASSIE_UNROLL
      for (int b = 0; b < 3; b++) x[r].b[b] -= jb[b] * w[r][c];
    }
    if ((c >> 1) == L) {
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) {
        const TC ja = je[3 + a];
        CASSIE_UNROLL
        for (int r = 0; r < NR; r++) x[r].l[a] -= ja * w[r][c];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The 14-variable box QP of osc_qp.cuh (block principal pivoting, same exchange rule, same tolerances) on four lanes.
// Every pass solves the KKT system of the current partition on the masked matrix (pinned variables become identity
// rows / columns).  Lane l owns rows l, l + 4, l + 8, l + 12 ("row slots" s = 0..3) of the matrix and of every vector
// and keeps them IN REGISTERS for the whole factorisation; a pivot column reaches the other lanes through a small
// double-buffered shared-memory column (one barrier per pivot: the owners publish the unscaled column, everybody
// reads pivot + column in one batch, scales locally and updates its rows), forward substitution broadcasts one
// unknown per step by shuffle, backward substitution sums the lanes' partial dots by shuffles.  Every loop is
// unrolled through template recursion over the pivot (a `#pragma unroll` nest of this size is silently left rolled
// by the compiler, and the row arrays then live in local memory: profiles/r2f_squat_osc.txt, 25 % of the kernel's
// stall samples; rolled shared-memory loops spent 37 % of the instructions on index arithmetic and, unrolled,
// serialised on the store -> load order of the shared array: r2g, r2i).  The partition logic runs replicated.
//   G: full symmetric 14 x 14, g: gradient (read only)    V: 5 x 14 doubles of vector scratch
struct QpRegs {
  TC L0[4], L1[8], L2[12], L3[14];   // row slot s holds row 4 s + l, entries 0 .. 4 s + 3 (those right of the diagonal unused)
  TC rhs[4];                         // right-hand side / solution entries of the own rows
  TC invd[kQpN];                     // 1 / L_jj, replicated
  TC z[kQpN];                        // solution, replicated
};
template <int S> QUAD_FN TC& qp_row(QpRegs& R, int k) {
  if (S == 0) return R.L0[k < 4 ? k : 3];
  if (S == 1) return R.L1[k < 8 ? k : 7];
  if (S == 2) return R.L2[k < 12 ? k : 11];
  return R.L3[k < 14 ? k : 13];
}
template <int J, int S>
QUAD_FN void qp_pivot_slot(QpRegs& R, const TC (&ck)[kQpN], TC inv) {
  if (4 * S + 3 >= J) {
    const TC lij = qp_row<S>(R, J) * inv;   // row = pivot row: becomes sqrt(d); rows above the pivot: unused word
    qp_row<S>(R, J) = lij;
    CASSIE_UNROLL
    for (int k = J + 1; k <= 4 * S + 3 && k < kQpN; k++) qp_row<S>(R, k) -= lij * ck[k];
  }
}
template <int J>
struct QpPivot {
  static QUAD_FN void run(QpRegs& R, const Lane ln, SV<TC> col, bool& ok) {
    const int l = ln.ql;
    const SV<TC> buf = col.at((J & 1) * kQpN);
    // owners publish the unscaled column J (rows >= J)
    if (4 * 0 + 3 >= J) { if (4 * 0 + l >= J) buf[4 * 0 + l] = qp_row<0>(R, J); }
    if (4 * 1 + 3 >= J) { if (4 * 1 + l >= J) buf[4 * 1 + l] = qp_row<1>(R, J); }
    if (4 * 2 + 3 >= J) { if (4 * 2 + l >= J) buf[4 * 2 + l] = qp_row<2>(R, J); }
    if (4 * 3 + l >= J && 4 * 3 + l < kQpN) buf[4 * 3 + l] = qp_row<3>(R, J);
    wsync();
    TC djj = buf[J];
    TC ck[kQpN];
    CASSIE_UNROLL
    for (int k = J + 1; k < kQpN; k++) ck[k] = buf[k];
    if (!(djj > 0.0)) { ok = false; djj = 1.0; }
#ifdef __CUDA_ARCH__
    const TC inv = rsqrt(djj);
#else
    const TC inv = 1.0 / sqrt(djj);
#endif
    R.invd[J] = inv;
    CASSIE_UNROLL
    for (int k = J + 1; k < kQpN; k++) ck[k] *= inv;
    qp_pivot_slot<J, 0>(R, ck, inv);
    qp_pivot_slot<J, 1>(R, ck, inv);
    qp_pivot_slot<J, 2>(R, ck, inv);
    qp_pivot_slot<J, 3>(R, ck, inv);
    QpPivot<J + 1>::run(R, ln, col, ok);
  }
};
template <>
struct QpPivot<kQpN> {
  static QUAD_FN void run(QpRegs&, const Lane, SV<TC>, bool&) {}
};
// forward substitution L y = rhs, one unknown per step, broadcast by shuffle
template <int K>
struct QpForward {
  static QUAD_FN void run(QpRegs& R, const Lane ln) {
    const int l = ln.ql;
    const TC yk = shfl(R.rhs[K / 4] * R.invd[K], K % 4);
    R.rhs[K / 4] = (K % 4 == l) ? yk : R.rhs[K / 4];
    if (4 * 0 + 3 > K) R.rhs[0] -= (4 * 0 + l > K) ? qp_row<0>(R, K) * yk : TC(0);
    if (4 * 1 + 3 > K) R.rhs[1] -= (4 * 1 + l > K) ? qp_row<1>(R, K) * yk : TC(0);
    if (4 * 2 + 3 > K) R.rhs[2] -= (4 * 2 + l > K) ? qp_row<2>(R, K) * yk : TC(0);
    R.rhs[3] -= (4 * 3 + l > K && 4 * 3 + l < kQpN) ? qp_row<3>(R, K) * yk : TC(0);
    QpForward<K + 1>::run(R, ln);
  }
};
template <>
struct QpForward<kQpN> {
  static QUAD_FN void run(QpRegs&, const Lane) {}
};
// backward substitution L' z = y: z_a = (y_a - sum_{k > a} L_ka z_k) / L_aa, the sum spread over the row owners
template <int A>
struct QpBackward {
  static QUAD_FN void run(QpRegs& R, const Lane ln) {
    const int l = ln.ql;
    TC part = TC(0);
    if (4 * 0 + 3 > A) part += (4 * 0 + l > A) ? qp_row<0>(R, A) * R.rhs[0] : TC(0);
    if (4 * 1 + 3 > A) part += (4 * 1 + l > A) ? qp_row<1>(R, A) * R.rhs[1] : TC(0);
    if (4 * 2 + 3 > A) part += (4 * 2 + l > A) ? qp_row<2>(R, A) * R.rhs[2] : TC(0);
    part += (4 * 3 + l > A && 4 * 3 + l < kQpN) ? qp_row<3>(R, A) * R.rhs[3] : TC(0);
    part += shx(part, 1);
    part += shx(part, 2);
    const TC za = shfl((R.rhs[A / 4] - part) * R.invd[A], A % 4);
    R.rhs[A / 4] = (A % 4 == l) ? za : R.rhs[A / 4];
    R.z[A] = za;
    QpBackward<A - 1>::run(R, ln);
  }
};
template <>
struct QpBackward<-1> {
  static QUAD_FN void run(QpRegs&, const Lane) {}
};

QUAD_FN void quad_box_qp(const Lane ln, SV<TC> G, SV<TC> g, SV<TC> V, const PlanarModel<TC>& m, TC z[kNU],
                         unsigned& at_lo, unsigned& at_hi, int max_iter, OscStats* st) {
  constexpr int N = kQpN;
  const int l = ln.ql;
  const SV<TC> zb = V, col = V.at(N), mult = V.at(3 * N), zfin = V.at(4 * N);
  SV<TC> Grow[4];
  CASSIE_UNROLL
  for (int s = 0; s < 4; s++) {
    const int i = 4 * s + l;
    Grow[s] = G.at((i < N ? i : N - 1) * N);
  }
  const bool has3 = l < 2;
  TC lo_own[4], hi_own[4], g_own[4];
  CASSIE_UNROLL
  for (int s = 0; s < 4; s++) {
    const int i = 4 * s + l;
    lo_own[s] = i < kNU ? m.act_lo[i < kNU ? i : 0] : TC(0);
    hi_own[s] = i < kNU ? m.act_hi[i < kNU ? i : 0] : TC(1e30);
    g_own[s] = g[i < N ? i : N - 1];
  }
  TC gscale = 1.0;
  CASSIE_UNROLL
  for (int i = 0; i < N; i++) gscale = fmax(gscale, fabs(g[i]));
  const TC dtol = 1e-12 * gscale;
  int it = 0, status = 1, best = N + 1;
  // The loop is WARP UNIFORM (its collectives name all 32 lanes): a quad whose QP is finished keeps iterating on its
  // frozen partition until the slowest quad of the warp is done; its outputs are latched in zfin / status / it.
  bool fin = false;
  CASSIE_ROLL
  for (int pass = 0; pass < max_iter; pass++) {
    const unsigned fixed = at_lo | at_hi;
    CASSIE_UNROLL
    for (int s = 0; s < 4; s++) {
      const int i = 4 * s + l;
      if (s < 3 || has3) zb[i] = ((at_lo >> i) & 1u) ? lo_own[s] : (((at_hi >> i) & 1u) ? hi_own[s] : TC(0));
    }
    wsync();
    QpRegs R;
    // masked KKT matrix (own rows, in registers) and right-hand side: rhs_i = pinned ? bound : -(g_i + sum_j G_ij zB_j)
    CASSIE_UNROLL
    for (int s = 0; s < 4; s++) {
      const int i = 4 * s + l;
      const bool pin = ((fixed >> i) & 1u) != 0u || i >= N;
      TC acc = TC(0), zbi = TC(0);
      CASSIE_UNROLL
      for (int j = 0; j < N; j++) {
        const TC gij = Grow[s][j];
        const TC zbj = zb[j];
        acc += gij * zbj;
        zbi = (j == i) ? zbj : zbi;
        if (j <= 4 * s + 3) {
          const TC kij = (pin || ((fixed >> j) & 1u)) ? (i == j ? TC(1) : TC(0)) : gij;
          if (s == 0) R.L0[j < 4 ? j : 3] = kij;
          else if (s == 1) R.L1[j < 8 ? j : 7] = kij;
          else if (s == 2) R.L2[j < 12 ? j : 11] = kij;
          else R.L3[j < 14 ? j : 13] = kij;
        }
      }
      R.rhs[s] = pin ? zbi : -(g_own[s] + acc);
    }
    bool ok = true;
    QpPivot<0>::run(R, ln, col, ok);
    QpForward<0>::run(R, ln);
    QpBackward<N - 1>::run(R, ln);
    // multipliers of the own rows: s_i = g_i + sum_j G_ij z_j
    CASSIE_UNROLL
    for (int s = 0; s < 4; s++) {
      const int i = 4 * s + l;
      if (s < 3 || has3) {
        TC acc = g_own[s];
        CASSIE_UNROLL
        for (int j = 0; j < N; j++) acc += Grow[s][j] * R.z[j];
        mult[i] = acc;
      }
    }
    wsync();
    // ---- violations: free variables outside their bounds, pinned variables with a wrong-sign multiplier (osc_qp.cuh)
    TC zmax = 1.0;
    CASSIE_UNROLL
    for (int i = 0; i < N; i++) zmax = fmax(zmax, fabs(R.z[i]));
    const TC ptol = 1e-8 * zmax;
    unsigned viol = 0u, below = 0u;
    int nviol = 0, last = -1, worst = -1;
    TC worst_mag = -1.0;
    CASSIE_UNROLL
    for (int i = 0; i < N; i++) {
      const TC lo_i = i < kNU ? m.act_lo[i < kNU ? i : 0] : TC(0), hi_i = i < kNU ? m.act_hi[i < kNU ? i : 0] : TC(1e30);
      const TC zi = R.z[i];
      TC v, mag;
      bool bad;
      if ((fixed >> i) & 1u) {
        const TC si = mult[i];
        v = ((at_lo >> i) & 1u) ? -si : si;
        bad = v > dtol;
        mag = v / gscale;
      } else {
        v = fmax(lo_i - zi, zi - hi_i);
        bad = v > ptol;
        mag = v / zmax;
      }
      if (zi < lo_i) below |= 1u << i;
      if (bad) {
        viol |= 1u << i; nviol++; last = i;
        if (mag > worst_mag) { worst_mag = mag; worst = i; }
      }
    }
    if (!fin) {
      if (l == 0) {
        CASSIE_UNROLL
        for (int i = 0; i < kNU; i++) zfin[i] = R.z[i];
      }
      if (!ok) { status = 2; fin = true; }
      else if (nviol == 0) { status = 0; it++; fin = true; }
      else {
        if (nviol < best) best = nviol;
        else viol = 1u << (it < kQpGreedyIters ? worst : last);
        const unsigned rel = viol & fixed, pinv = viol & ~fixed;   // pinned violators are released, free ones pinned
        at_lo = (at_lo & ~rel) | (pinv & below);
        at_hi = (at_hi & ~rel) | (pinv & ~below);
        it++;
      }
    }
    if (!wany(!fin)) break;
  }
  wsync();
  CASSIE_UNROLL
  for (int i = 0; i < kNU; i++) {   // only the motor commands leave the QP
    const TC lo_i = m.act_lo[i], hi_i = m.act_hi[i];
    const TC zi = zfin[i];
    z[i] = zi < lo_i ? lo_i : (zi > hi_i ? hi_i : zi);
  }
  if (st) { st->iters = it; st->status = status; }
}

// ---------------------------------------------------------------------------------------------------------------
// OSC_RBDL::RunPTSC + SolveQP (OSC_RBDL.cpp:114-291) for the planar model: osc_qp.cuh osc_control on four lanes.
// act = ControllerOsc in memory order (RobotInterface.h:23-28).  u (ctrl units) is written to St[u] by lane 0.
template <typename T>
QUAD_FN void quad_osc(const PlanarModel<TC>& m, const Lane ln, const LegKin<TC>& k, const V8<TC>& qd, const TC act[7],
                      SV<T> St, SV<TC> C, unsigned& qp_set, OscStats* st) {
  typedef CtrlLayout Y;
  const int L = ln.L, h = ln.h;
  QuadCtrlDyn d;
  // site (L, h) of the toe: task rows and Jdot qd - xdd*   (OSC_RBDL.cpp:123-144)
  {
    const int site = 2 + 2 * L + h;
    TC rx, rz;
    rot(k.c[kToe], k.s[kToe], m.site_off[site][0], m.site_off[site][1], rx, rz);
    TC Jx[8], Jz[8];
    leg_point_jac<kToe>(m, L, k, rx, rz, Jx, Jz);
    const SV<TC> dx = C.at(Y::task + ((2 * L + h) * 2) * 8), dz = C.at(Y::task + ((2 * L + h) * 2 + 1) * 8);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) { dx[c] = Jx[c]; dz[c] = Jz[c]; }
    // pivot accelerations are produced by quad_ctrl_dynamics below; the toe's is recomputed here from the chain
    TC ax = TC(0), az = TC(0);
    {
      const TC w0 = k.w0 * k.w0;
      ax -= w0 * k.dx[kThigh]; az -= w0 * k.dz[kThigh];
      const TC w1 = k.w[kThigh] * k.w[kThigh];
      ax -= w1 * k.dx[kKnee]; az -= w1 * k.dz[kKnee];
      const TC w2 = k.w[kKnee] * k.w[kKnee];
      ax -= w2 * k.dx[kTarsus]; az -= w2 * k.dz[kTarsus];
      const TC w3 = k.w[kTarsus] * k.w[kTarsus];
      ax -= w3 * k.dx[kToe]; az -= w3 * k.dz[kToe];
    }
    const TC wt = k.w[kToe] * k.w[kToe];
    const int r = 2 + 2 * (2 * L + h);
    C[Y::e0 + r] = ax - wt * rx - act[2 + 2 * L];
    C[Y::e0 + r + 1] = az - wt * rz - act[3 + 2 * L];
  }
  TC blx, blz;   // lever of body_center from the pelvis pivot
  rot(k.c0, k.s0, m.site_off[1][0], m.site_off[1][1], blx, blz);
  if (ln.ql == 0) {
    const TC w2 = k.w0 * k.w0;
    C[Y::e0 + 0] = -w2 * blx - act[0];
    C[Y::e0 + 1] = -w2 * blz - act[1];
    C[Y::e0 + 10] = -act[6];
    C[Y::e0 + 11] = TC(0);
  }
  quad_ctrl_dynamics(m, ln, k, qd, C, d);   // ends with every lane past the barrier that publishes task rows and e0
  phase_sync<2>();
  // ---- task loop: half h takes the six tasks of leg h (h = 0: body x, body z, left sites; h = 1: pitch, padding,
  // right sites).  Z_r = Mc^-1 A_r' = M^-1 (A_r' - Jeq' P JH A_r'); E_r = Z_r' [Bt, Jc' T]; r0_r = Jdot qd - xdd* + A_r p0
  TC Js[2][2][8];   // own leg's two sites (x, z rows): the contact generators of this leg
  CASSIE_UNROLL
  for (int s = 0; s < 2; s++) {
    CASSIE_UNROLL
    for (int xz = 0; xz < 2; xz++) {
      const SV<TC> src = C.at(Y::task + ((2 * L + s) * 2 + xz) * 8);
      CASSIE_UNROLL
      for (int c = 0; c < 8; c++) Js[s][xz][c] = src[c];
    }
  }
  CASSIE_ROLL
  for (int t0 = 0; t0 < 6; t0 += 2) {   // two tasks per pass: independent dependency chains through the solve
    V8<TC> x[2];
    TC y[2][4];
    int rr[2];
    CASSIE_UNROLL
    for (int i = 0; i < 2; i++) {
      const int t = t0 + i;
      rr[i] = t >= 2 ? 2 + 4 * h + (t - 2) : (h ? 10 + t : t);
      if (t0 >= 2) {
        const SV<TC> src = C.at(Y::task + (4 * h + (t - 2)) * 8);
        CASSIE_UNROLL
        for (int b = 0; b < 3; b++) x[i].b[b] = src[b];
        CASSIE_UNROLL
        for (int a = 0; a < kLegLinks; a++) x[i].l[a] = (L == h) ? src[3 + a] : TC(0);
      } else {
        // body_center x / z rows (half 0), pitch row (AddQDDIdx(2), Cassie2d.cpp:41) and the padding row (half 1)
        x[i].b[0] = (h == 0 && i == 0) ? TC(1) : TC(0);
        x[i].b[1] = (h == 0 && i == 1) ? TC(1) : TC(0);
        x[i].b[2] = h == 0 ? (i == 0 ? blz : -blx) : (i == 0 ? TC(1) : TC(0));
        CASSIE_UNROLL
        for (int a = 0; a < kLegLinks; a++) x[i].l[a] = TC(0);
      }
    }
    quad_project<2>(ln, C, d, x, y);
    quad_solve<2>(C.at(Y::ld), L, x);
    CASSIE_UNROLL
    for (int i = 0; i < 2; i++) {
      const int r = rr[i];
      const SV<TC> Er = C.at(Y::E + r * kQpN);
      // actuator columns of this leg: gear * Z_r[actuated dof]
      CASSIE_UNROLL
      for (int a = 0; a < kNU; a++) {
        TC v = TC(0);
        bool mine = false;
        CASSIE_UNROLL
        for (int j = 0; j < kLegLinks; j++)
          if (m.act_dof[a] == 3 + 5 * L + j) { v = x[i].l[j]; mine = true; }
        if (mine) Er[a] = m.act_gear[a] * v;
      }
      CASSIE_UNROLL
      for (int s = 0; s < 2; s++) {
        const TC ex = dot8(Js[s][0], x[i]), ez = dot8(Js[s][1], x[i]);
        Er[kNU + 2 * (2 * L + s)] = kOscMu * ex + ez;
        Er[kNU + 2 * (2 * L + s) + 1] = -kOscMu * ex + ez;
      }
      TC pl = TC(0);
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) pl += x[i].l[a] * d.bias.l[a];
      const TC zb = (x[i].b[0] * d.bias.b[0] + x[i].b[1] * d.bias.b[1] + x[i].b[2] * d.bias.b[2]) + sum_legs(pl);
      TC ys = TC(0);
      CASSIE_UNROLL
      for (int c = 0; c < 4; c++) ys += y[i][c] * d.sjd[c];
      if (L == 0) C[Y::r0 + r] = C[Y::e0 + r] - zb - ys;
    }
  }
  wsync();
  phase_sync<2>();
  // ---- G = 2 E'WE (packed lower triangle), g = 2 E'W r0 (OSC_RBDL.cpp:186-203): lane l builds the row pairs l and 6 - l
  {
    const SV<TC> G = C.at(Y::G), g = C.at(Y::g);
    CASSIE_ROLL
    for (int pass = 0; pass < 2; pass++) {
      const int pr = pass == 0 ? ln.ql : 6 - ln.ql;
      if (pass == 1 && ln.ql == 3) break;
      const int i = 2 * pr;
      TC c0[kQpTasks], c1[kQpTasks];
      TC g0 = TC(0), g1 = TC(0);
      CASSIE_UNROLL
      for (int r = 0; r < kQpTasks; r++) {
        const TC W = r < 2 ? kOscWCom : (r < 10 ? kOscWStance : kOscWRest);
        const TC w2 = TC(2) * W;
        c0[r] = w2 * C[Y::E + r * kQpN + i]; c1[r] = w2 * C[Y::E + r * kQpN + i + 1];
        const TC rr = C[Y::r0 + r];
        g0 += c0[r] * rr; g1 += c1[r] * rr;
      }
      g[i] = g0; g[i + 1] = g1;
      CASSIE_ROLL
      for (int j = 0; j <= i + 1; j++) {
        TC s0a = TC(0), s0b = TC(0), s1a = TC(0), s1b = TC(0);
        CASSIE_UNROLL
        for (int r = 0; r < kQpTasks; r++) {
          const TC ej = C[Y::E + r * kQpN + j];
          if (r & 1) { s0b += c0[r] * ej; s1b += c1[r] * ej; }
          else { s0a += c0[r] * ej; s1a += c1[r] * ej; }
        }
        TC v0 = s0a + s0b, v1 = s1a + s1b;
        if (i >= kNU) {   // cost of the contact forces: 1e-4 |T lambda|^2 (OSC_RBDL.cpp:188-201)
          if (j == i) v0 += kOscWForce * (kOscMu * kOscMu + 1.0);
          if (j == i + 1) v1 += kOscWForce * (kOscMu * kOscMu + 1.0);
          if (j == i) v1 += kOscWForce * (1.0 - kOscMu * kOscMu);
        }
        if (j <= i) { G[i * kQpN + j] = v0; G[j * kQpN + i] = v0; }
        G[(i + 1) * kQpN + j] = v1; G[j * kQpN + i + 1] = v1;
      }
    }
  }
  wsync();
  phase_sync<2>();
  // ---- the box QP, cooperatively (quad_box_qp)
  {
    TC z[kNU];
    unsigned at_lo = qp_set & 0x3fffu, at_hi = (qp_set >> 14) & 0x3fu;
    quad_box_qp(ln, C.at(Y::G), C.at(Y::g), C.at(Y::qpv), m, z, at_lo, at_hi, 300, st);
    qp_set = at_lo | (at_hi << 14);
    if (ln.ql == 0) {
      CASSIE_UNROLL
      for (int a = 0; a < kNU; a++) St[StateLayout::u + a] = (T)z[a];
    }
  }
  wsync();
}

// ---------------------------------------------------------------------------------------------------------------
// Cassie2d::StepJacobian control law (Cassie2d.cpp:119-165): u = pinv(Nc Bt) (Nc bias + gamma - Nc Jc6' f),
// f per foot = (Fx, Fz, My).  The projections are leg-split; the 13 x 6 pseudo-inverse runs on lane 0.
template <typename T>
QUAD_FN void quad_jacobian_control(const PlanarModel<TC>& m, const Lane ln, const LegKin<TC>& k, const V8<TC>& qd,
                                   const TC f[6], SV<T> St, SV<TC> C) {
  typedef CtrlLayout Y;
  const int L = ln.L;
  QuadCtrlDyn d;
  quad_ctrl_dynamics(m, ln, k, qd, C, d);
  V8<TC> x = d.bias;
  {
    const TC Fx = f[3 * L], Fz = f[3 * L + 1], My = f[3 * L + 2];
    TC rx, rz, Jx[8], Jz[8];
    const int s0 = 2 + 2 * L, s1 = 3 + 2 * L;
    rot(k.c[kToe], k.s[kToe], TC(0.5) * (m.site_off[s0][0] + m.site_off[s1][0]), TC(0.5) * (m.site_off[s0][1] + m.site_off[s1][1]), rx, rz);
    leg_point_jac<kToe>(m, L, k, rx, rz, Jx, Jz);
    TC pb[3];
    pb[0] = Jx[0] * Fx + Jz[0] * Fz;
    pb[1] = Jx[1] * Fx + Jz[1] * Fz;
    pb[2] = Jx[2] * Fx + Jz[2] * Fz + My;
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) x.b[b] -= sum_legs(pb[b]);
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++)
      x.l[a] -= Jx[3 + a] * Fx + Jz[3 + a] * Fz + (link_anc(kToe, a) ? m.sgn[L][a] : TC(0)) * My;
  }
  TC y[1][4];
  {
    V8<TC> xs[1] = {x};
    quad_project<1>(ln, C, d, xs, y);
    x = xs[0];
  }
  // + gamma = Jeq' P JeqdotQdot
  CASSIE_UNROLL
  for (int c = 0; c < 4; c++) {
    const SV<TC> je = C.at(Y::jeq + c * 8);
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) x.b[b] += je[b] * d.sjd[c];
    if ((c >> 1) == L) {
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) x.l[a] += je[3 + a] * d.sjd[c];
    }
  }
  if (ln.h == 0) {
    if (L == 0) {
      CASSIE_UNROLL
      for (int b = 0; b < 3; b++) C[Y::jacRhs + b] = x.b[b];
    }
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) C[Y::jacRhs + 3 + 5 * L + a] = x.l[a];
  }
  CASSIE_ROLL
  for (int a = 0; a < kNU; a++) {
    V8<TC> col;
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) col.b[b] = TC(0);
    CASSIE_UNROLL
    for (int i = 0; i < kLegLinks; i++) col.l[i] = m.act_dof[a] == 3 + 5 * L + i ? m.act_gear[a] : TC(0);
    {
      V8<TC> cs[1] = {col};
      quad_project<1>(ln, C, d, cs, y);
      col = cs[0];
    }
    if (ln.h == 0) {
      if (L == 0) {
        CASSIE_UNROLL
        for (int b = 0; b < 3; b++) C[Y::jacB + b * kNU + a] = col.b[b];
      }
      CASSIE_UNROLL
      for (int i = 0; i < kLegLinks; i++) C[Y::jacB + (3 + 5 * L + i) * kNU + a] = col.l[i];
    }
  }
  wsync();
  if (ln.ql == 0) {
    TC B[kNV][kNU], rhs[kNV], u[kNU];
    for (int i = 0; i < kNV; i++) {
      rhs[i] = C[Y::jacRhs + i];
      for (int a = 0; a < kNU; a++) B[i][a] = C[Y::jacB + i * kNU + a];
    }
    pinv13x6_apply(B, TC(1e-4), rhs, u);
    for (int a = 0; a < kNU; a++) St[StateLayout::u + a] = (T)u[a];
  }
  wsync();
}

// ---------------------------------------------------------------------------------------------------------------
// One legacy Step* call (Cassie2d.cpp:86-209) of the env owned by this quad: controller (on the RBDL view of the
// model), then one physics step (on MuJoCo's view).  cassie_step.cuh controller_step on four lanes.
//   St  persistent state block (StateLayout, type T)     scratch: block used as PhysLayout (T) and CtrlLayout (double)
//   act the action of the step (uniform over the quad)   want_op: refresh the lagged op-space state (always done in
//   the Jacobian / OSC modes, whose controllers need the controller-model kinematics anyway)
template <int MODE, bool CTA = false, typename T, typename TG>
QUAD_FN void quad_controller_step(const PlanarModel<T>& mphys, const PlanarModel<TG>& mphys_g, const PlanarModel<TC>& mctrl,
                                  const Lane ln, SV<T> St, void* scratch, int ei, const T* act, bool want_op, QStepStats* st,
                                  OscStats* qst, unsigned* qp_set) {
  typedef StateLayout X;
  // `scratch` = start of the warp's scratch region, ei = env of this quad within the warp (0 on the host)
  const SV<T> Wp{reinterpret_cast<T*>(scratch) + ei};
  const SV<TC> Wc{reinterpret_cast<TC*>(scratch) + ei};
  if (MODE == kModeTorque || MODE == kModePd) {
    if (ln.ql == 0) {
      CASSIE_UNROLL
      for (int a = 0; a < kNU; a++) {
        if (MODE == kModeTorque) St[X::u + a] = act[a];
        else {   // Cassie2d.cpp:96-112: gains are in ctrl units
          const T qj = St[X::q + mphys.act_dof[a]], vj = St[X::qd + mphys.act_dof[a]];
          St[X::u + a] = T(10) * (act[a] - qj) + T(5) * (T(0) - vj);
        }
      }
    }
  }
  if (want_op || MODE >= kModeJacobian) {
    V8<TC> qc, qdc;
    LegKin<TC> kc;
    quad_ctrl_kin(mctrl, ln, St, qc, qdc, kc, true);
    if (MODE >= kModeJacobian) {
      TC ac[7];
      CASSIE_UNROLL
      for (int i = 0; i < (MODE == kModeOsc ? 7 : 6); i++) ac[i] = (TC)act[i];
      if (MODE == kModeJacobian) quad_jacobian_control(mctrl, ln, kc, qdc, ac, St, Wc);
      else quad_osc(mctrl, ln, kc, qdc, ac, St, Wc, *qp_set, qst);
    }
  }
  wsync();
  phase_sync<1>();
  quad_physics_step<CTA>(mphys, mphys_g, ln, St, Wp, st);
}

}  // namespace quad
}  // namespace cassie
