// Quad engine, physics side: one mj_step [EXT] (Cassie2d.cpp:92,115,174,206) of one env by FOUR lanes.
//
// Same arithmetic as planar_engine.cuh (the thread-per-env engine, kept for the legacy batch-of-one ABI and as the
// serial fallback of the rare-row regime); what changes is who holds what:
//   * a 13-vector lives "leg split": lane (L, h) holds the three base entries (replicated) and the five entries of
//     leg L; dots with a constraint row (non-zero on the base and ONE leg) are local to that leg's lanes;
//   * M = L^T D L is factored leg by leg (the two legs only meet in the 3x3 base block: one exchange), and
//     M^-1 x is a leg-local backward pass, a 3-value exchange, the base 3x3, and a leg-local forward pass;
//   * A = J M^-1 J^T: the half h of a quad solves for the six rows of leg h as right-hand sides; lane (L, h) then
//     dots the result with the six rows of ITS leg -- no reduction across lanes;
//   * PGS: lane l owns scalar row l and contact pair l (A rows, residual accumulators, forces in registers); a final
//     force is one shuffle and one FMA per owned row away from every other lane ("publish", planar_engine.cuh).
// Per-env scratch is a shared-memory block laid out [field][8 envs] (quad_rt.cuh).
#pragma once
#include "quad_rt.cuh"

#ifdef CASSIE_HOST_HARNESS
extern bool cassie_force_tier1;   // test hook: run the 16-row tier on states the 12-row tier could handle
extern bool cassie_force_tier2;   // test hook: run the 20-row tier on states the smaller tiers could handle
#endif

namespace cassie {
namespace quad {

// ---------------------------------------------------------------------------------------------------------------
// leg-split vector: base part (same on both legs' lanes) + this lane's leg part
template <typename R>
struct V8 {
  R b[3], l[5];
};

// kinematics of the pelvis and ONE leg (planar_engine.cuh Kin, leg-local)
template <typename R>
struct LegKin {
  R c0, s0, w0, v0x, v0z;
  R c[kLegLinks], s[kLegLinks], dx[kLegLinks], dz[kLegLinks], px[kLegLinks], pz[kLegLinks];
  R w[kLegLinks], vx[kLegLinks], vz[kLegLinks];
};

// mj_kinematics [EXT] / RBDL UpdateKinematics (DynamicModel.cpp:237-242): positions of one leg
template <typename R>
QUAD_FN void leg_fk_positions(const PlanarModel<R>& m, int L, const V8<R>& q, LegKin<R>& k) {
  const R a0 = q.b[2] - m.pel_ref[2];
  Num<R>::sincos_(a0, &k.s0, &k.c0);
  R alpha[kLegLinks];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    const int p = link_parent(a);
    R ap, cp, sp, ppx, ppz;
    if (p < 0) { ap = a0; cp = k.c0; sp = k.s0; ppx = R(0); ppz = R(0); }
    else { ap = alpha[p]; cp = k.c[p]; sp = k.s[p]; ppx = k.px[p]; ppz = k.pz[p]; }
    alpha[a] = ap + m.sgn[L][a] * q.l[a] + m.ang0[L][a];
    Num<R>::sincos_(alpha[a], &k.s[a], &k.c[a]);
    R dx, dz;
    rot(cp, sp, m.off[L][a][0], m.off[L][a][1], dx, dz);
    k.dx[a] = dx; k.dz[a] = dz;
    k.px[a] = ppx + dx; k.pz[a] = ppz + dz;
  }
}
template <typename R>
QUAD_FN void leg_fk_velocities(const PlanarModel<R>& m, int L, const V8<R>& qd, LegKin<R>& k) {
  k.w0 = qd.b[2]; k.v0x = qd.b[0]; k.v0z = qd.b[1];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    const int p = link_parent(a);
    R wp, vpx, vpz;
    if (p < 0) { wp = k.w0; vpx = k.v0x; vpz = k.v0z; }
    else { wp = k.w[p]; vpx = k.vx[p]; vpz = k.vz[p]; }
    k.w[a] = wp + m.sgn[L][a] * qd.l[a];
    k.vx[a] = vpx + wp * k.dz[a];
    k.vz[a] = vpz - wp * k.dx[a];
  }
}
template <typename R, typename RG>
QUAD_FN void cast_leg_positions(const LegKin<RG>& g, LegKin<R>& k) {
  k.c0 = (R)g.c0; k.s0 = (R)g.s0;
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    k.c[a] = (R)g.c[a]; k.s[a] = (R)g.s[a];
    k.dx[a] = (R)g.dx[a]; k.dz[a] = (R)g.dz[a];
    k.px[a] = (R)g.px[a]; k.pz[a] = (R)g.pz[a];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Mass matrix in leg form.  lb[a][b] = M[leg dof a][base dof b]; ll = the seven leg-leg couplings (knee-thigh,
// tarsus-thigh, tarsus-knee, toe-thigh, toe-knee, toe-tarsus, rod-thigh); d = leg diagonal; bb = base lower triangle
// (0,0) (1,0) (1,1) (2,0) (2,1) (2,2), complete (both legs + pelvis) and identical on every lane.
CASSIE_HD constexpr int pair_idx(int a, int b) { return a == 4 ? 6 : (a * (a - 1)) / 2 + b; }   // a > b, b ancestor of a
template <typename R>
struct LegM {
  R lb[kLegLinks][3], ll[7], d[kLegLinks], bb[6];
};

// mj_crb [EXT] / RBDL CompositeRigidBodyAlgorithm + rotor inertia (DynamicModel.cpp:267-272); planar_engine.cuh
// mass_matrix restricted to one leg, the base block summed over the legs by one exchange
template <typename R>
QUAD_FN void leg_mass_matrix(const PlanarModel<R>& m, int L, const LegKin<R>& k, LegM<R>& M) {
  R cm[kLegLinks], hx[kLegLinks], hz[kLegLinks], cI[kLegLinks];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    R rx, rz;
    rot(k.c[a], k.s[a], m.com[L][a][0], m.com[L][a][1], rx, rz);
    cm[a] = m.mass[L][a];
    hx[a] = cm[a] * rx; hz[a] = cm[a] * rz;
    cI[a] = m.inertia[L][a] + cm[a] * (rx * rx + rz * rz);
  }
  CASSIE_UNROLL
  for (int step = 0; step < 4; step++) {
    const int c = step == 0 ? kToe : (step == 1 ? kTarsus : (step == 2 ? kKnee : kRod));
    const int p = link_parent(c);
    const R dx = k.dx[c], dz = k.dz[c];
    cI[p] += cI[c] + R(2) * (dx * hx[c] + dz * hz[c]) + cm[c] * (dx * dx + dz * dz);
    hx[p] += hx[c] + cm[c] * dx;
    hz[p] += hz[c] + cm[c] * dz;
    cm[p] += cm[c];
  }
  // this leg's subtree about the pelvis pivot
  R lI, lhx, lhz, lm;
  {
    const R dx = k.dx[kThigh], dz = k.dz[kThigh];
    lI = cI[kThigh] + R(2) * (dx * hx[kThigh] + dz * hz[kThigh]) + cm[kThigh] * (dx * dx + dz * dz);
    lhx = hx[kThigh] + cm[kThigh] * dx;
    lhz = hz[kThigh] + cm[kThigh] * dz;
    lm = cm[kThigh];
  }
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    const R sa = m.sgn[L][a];
    M.lb[a][0] = sa * hz[a];
    M.lb[a][1] = -sa * hx[a];
    M.lb[a][2] = sa * (cI[a] + k.px[a] * hx[a] + k.pz[a] * hz[a]);
    M.d[a] = cI[a] + m.armature[3 + 5 * L + a];
    R lx = R(0), lz = R(0);
    int b = a;
    CASSIE_UNROLL
    for (int hop = 0; hop < 3; hop++) {
      if (link_parent(b) >= 0) {
        lx += k.dx[b]; lz += k.dz[b];
        b = link_parent(b);
        M.ll[pair_idx(a, b)] = sa * m.sgn[L][b] * (cI[a] + lx * hx[a] + lz * hz[a]);
      }
    }
  }
  R tm = m.pel_mass, thx, thz, tI;
  {
    R cx, cz;
    rot(k.c0, k.s0, m.pel_com[0], m.pel_com[1], cx, cz);
    thx = m.pel_mass * cx; thz = m.pel_mass * cz;
    tI = m.pel_inertia + m.pel_mass * (cx * cx + cz * cz);
  }
  tm += sum_legs(lm); thx += sum_legs(lhx); thz += sum_legs(lhz); tI += sum_legs(lI);
  M.bb[0] = tm + m.armature[0];
  M.bb[1] = R(0);
  M.bb[2] = tm + m.armature[1];
  M.bb[3] = thz;
  M.bb[4] = -thx;
  M.bb[5] = tI + m.armature[2];
}

// RNE with qdd = 0 (mj_rne [EXT]; RBDL NonlinearEffects, DynamicModel.cpp:320-323): leg part + base sum
template <typename R>
QUAD_FN void leg_bias_forces(const PlanarModel<R>& m, int L, const LegKin<R>& k, V8<R>& bias) {
  const R g = -m.gravity_z;
  R ax[kLegLinks], az[kLegLinks], fx[kLegLinks], fz[kLegLinks], n[kLegLinks];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    const int p = link_parent(a);
    R apx, apz, wp;
    if (p < 0) { apx = R(0); apz = R(0); wp = k.w0; }
    else { apx = ax[p]; apz = az[p]; wp = k.w[p]; }
    const R w2p = wp * wp;
    ax[a] = apx - w2p * k.dx[a];
    az[a] = apz - w2p * k.dz[a];
    R rx, rz;
    rot(k.c[a], k.s[a], m.com[L][a][0], m.com[L][a][1], rx, rz);
    const R w2 = k.w[a] * k.w[a];
    fx[a] = m.mass[L][a] * (ax[a] - w2 * rx);
    fz[a] = m.mass[L][a] * (az[a] - w2 * rz + g);
    n[a] = rz * fx[a] - rx * fz[a];
  }
  CASSIE_UNROLL
  for (int step = 0; step < 4; step++) {
    const int c = step == 0 ? kToe : (step == 1 ? kTarsus : (step == 2 ? kKnee : kRod));
    const int p = link_parent(c);
    n[p] += n[c] + (k.dz[c] * fx[c] - k.dx[c] * fz[c]);
    fx[p] += fx[c];
    fz[p] += fz[c];
  }
  const R lN = n[kThigh] + (k.dz[kThigh] * fx[kThigh] - k.dx[kThigh] * fz[kThigh]);
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) bias.l[a] = m.sgn[L][a] * n[a];
  R Fx, Fz, N;
  {
    R rx, rz;
    rot(k.c0, k.s0, m.pel_com[0], m.pel_com[1], rx, rz);
    const R w2 = k.w0 * k.w0;
    Fx = m.pel_mass * (-w2 * rx);
    Fz = m.pel_mass * (-w2 * rz + g);
    N = rz * Fx - rx * Fz;
  }
  bias.b[0] = Fx + sum_legs(fx[kThigh]);
  bias.b[1] = Fz + sum_legs(fz[kThigh]);
  bias.b[2] = N + sum_legs(lN);
}

// ---------------------------------------------------------------------------------------------------------------
// M = L^T D L in leg form (mj_factorM [EXT]; planar_engine.cuh factor).  In place: lb / ll become the unit factor,
// d / the diagonal of bb are replaced by 1 / D.  bb off-diagonals become L10, L20, L21.
template <typename R>
QUAD_FN void leg_factor(LegM<R>& M) {
  R db[6];   // this leg's eliminations acting on the base block
  CASSIE_UNROLL
  for (int i = 0; i < 6; i++) db[i] = R(0);
  CASSIE_UNROLL
  for (int k = kLegLinks - 1; k >= 0; k--) {
    const R inv = Num<R>::rcp_(M.d[k]);
    M.d[k] = inv;
    // unscaled row k over its ancestors: base 0..2, then leg ancestors
    R ub[3], lbk[3];
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) { ub[b] = M.lb[k][b]; lbk[b] = ub[b] * inv; }
    CASSIE_UNROLL
    for (int i = 0; i < kLegLinks; i++) {
      if (i < k && link_anc(k, i)) {
        const R ui = M.ll[pair_idx(k, i)];
        const R li = ui * inv;
        // (i, j) over leg ancestors j <= i of k, and the base
        M.d[i] -= li * ui;
        CASSIE_UNROLL
        for (int j = 0; j < kLegLinks; j++)
          if (j < i && link_anc(k, j)) M.ll[pair_idx(i, j)] -= li * M.ll[pair_idx(k, j)];
        CASSIE_UNROLL
        for (int b = 0; b < 3; b++) M.lb[i][b] -= li * ub[b];
      }
    }
    // base-base pairs
    db[0] -= lbk[0] * ub[0];
    db[1] -= lbk[1] * ub[0];
    db[2] -= lbk[1] * ub[1];
    db[3] -= lbk[2] * ub[0];
    db[4] -= lbk[2] * ub[1];
    db[5] -= lbk[2] * ub[2];
    // row k becomes the unit factor
    CASSIE_UNROLL
    for (int i = 0; i < kLegLinks; i++)
      if (i < k && link_anc(k, i)) M.ll[pair_idx(k, i)] *= inv;
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) M.lb[k][b] = lbk[b];
  }
  CASSIE_UNROLL
  for (int i = 0; i < 6; i++) M.bb[i] += sum_legs(db[i]);
  // base 3x3: k = 2, 1, 0
  {
    const R inv2 = Num<R>::rcp_(M.bb[5]);
    const R u20 = M.bb[3], u21 = M.bb[4];
    const R l20 = u20 * inv2, l21 = u21 * inv2;
    M.bb[2] -= l21 * u21;
    M.bb[1] -= l21 * u20;
    M.bb[0] -= l20 * u20;
    M.bb[3] = l20; M.bb[4] = l21; M.bb[5] = inv2;
    const R inv1 = Num<R>::rcp_(M.bb[2]);
    const R u10 = M.bb[1];
    const R l10 = u10 * inv1;
    M.bb[0] -= l10 * u10;
    M.bb[1] = l10; M.bb[2] = inv1;
    M.bb[0] = Num<R>::rcp_(M.bb[0]);
  }
}

// Factor as stored in shared memory: 33 entries per leg lane pair + the base.  Layout per model:
//   [0..5]   base: 1/D0, L10, 1/D1, L20, L21, 1/D2   (LegM.bb order)
//   [6 + 27 L ...] leg L: lb (15), ll (7), 1/d (5)
constexpr int kLdSize = 6 + 2 * 27;
template <typename S, typename R>
QUAD_FN void store_factor(SV<S> ld, int L, const LegM<R>& M, bool store_base) {
  if (store_base) {
    CASSIE_UNROLL
    for (int i = 0; i < 6; i++) ld[i] = (S)M.bb[i];
  }
  const SV<S> p = ld.at(6 + 27 * L);
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) p[3 * a + b] = (S)M.lb[a][b];
  }
  CASSIE_UNROLL
  for (int i = 0; i < 7; i++) p[15 + i] = (S)M.ll[i];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) p[22 + a] = (S)M.d[a];
}

// x <- M^-1 x for NR right-hand sides at once (mj_solveLD [EXT]); every entry of the factor is loaded once and used NR
// times.  x[r].b must hold the COMPLETE base part on both legs' lanes, x[r].l this lane's leg part.
template <int NR, typename R, typename S>
QUAD_FN void quad_solve(SV<S> ld, int L, V8<R> x[NR]) {
  const SV<S> p = ld.at(6 + 27 * L);
  R db[NR][3];
  CASSIE_UNROLL
  for (int r = 0; r < NR; r++) { db[r][0] = R(0); db[r][1] = R(0); db[r][2] = R(0); }
  CASSIE_UNROLL
  for (int k = kLegLinks - 1; k >= 0; k--) {
    CASSIE_UNROLL
    for (int i = 0; i < kLegLinks; i++) {
      if (i < k && link_anc(k, i)) {
        const R l = (R)p[15 + pair_idx(k, i)];
        CASSIE_UNROLL
        for (int r = 0; r < NR; r++) x[r].l[i] -= l * x[r].l[k];
      }
    }
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) {
      const R l = (R)p[3 * k + b];
      CASSIE_UNROLL
      for (int r = 0; r < NR; r++) db[r][b] -= l * x[r].l[k];
    }
  }
  const R l10 = (R)ld[1], l20 = (R)ld[3], l21 = (R)ld[4];
  const R i0 = (R)ld[0], i1 = (R)ld[2], i2 = (R)ld[5];
  CASSIE_UNROLL
  for (int r = 0; r < NR; r++) {
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) x[r].b[b] += sum_legs(db[r][b]);
    x[r].b[0] -= l20 * x[r].b[2];
    x[r].b[1] -= l21 * x[r].b[2];
    x[r].b[0] -= l10 * x[r].b[1];
    x[r].b[0] *= i0; x[r].b[1] *= i1; x[r].b[2] *= i2;
    x[r].b[1] -= l10 * x[r].b[0];
    x[r].b[2] -= l20 * x[r].b[0];
    x[r].b[2] -= l21 * x[r].b[1];
  }
  CASSIE_UNROLL
  for (int k = 0; k < kLegLinks; k++) {
    const R dk = (R)p[22 + k];
    CASSIE_UNROLL
    for (int r = 0; r < NR; r++) x[r].l[k] *= dk;
  }
  CASSIE_UNROLL
  for (int k = 0; k < kLegLinks; k++) {
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) {
      const R l = (R)p[3 * k + b];
      CASSIE_UNROLL
      for (int r = 0; r < NR; r++) x[r].l[k] -= l * x[r].b[b];
    }
    CASSIE_UNROLL
    for (int i = 0; i < kLegLinks; i++) {
      if (i < k && link_anc(k, i)) {
        const R l = (R)p[15 + pair_idx(k, i)];
        CASSIE_UNROLL
        for (int r = 0; r < NR; r++) x[r].l[k] -= l * x[r].l[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// point Jacobian (x row, z row) of a point fixed to link A of this lane's leg (A = -1: pelvis), lever (rx, rz) from
// that link's pivot; J8 = [x, z, pitch, thigh, knee, tarsus, toe, rod] (planar_engine.cuh point_jac)
template <int A, typename R>
QUAD_FN void leg_point_jac(const PlanarModel<R>& m, int L, const LegKin<R>& k, R rx, R rz, R Jx[8], R Jz[8]) {
  CASSIE_UNROLL
  for (int b = 0; b < kLegLinks; b++) { Jx[3 + b] = R(0); Jz[3 + b] = R(0); }
  R lx = rx, lz = rz;
  int b = A;
  CASSIE_UNROLL
  for (int hop = 0; hop < 4; hop++) {
    if (b >= 0) {
      Jx[3 + b] = m.sgn[L][b] * lz;
      Jz[3 + b] = -m.sgn[L][b] * lx;
      lx += k.dx[b]; lz += k.dz[b];
      b = link_parent(b);
    }
  }
  Jx[0] = R(1); Jx[1] = R(0); Jx[2] = lz;
  Jz[0] = R(0); Jz[1] = R(1); Jz[2] = -lx;
}

template <typename R>
QUAD_FN R dot8(const R J[8], const V8<R>& x) {
  R s = J[0] * x.b[0] + J[1] * x.b[1] + J[2] * x.b[2];
  CASSIE_UNROLL
  for (int b = 0; b < kLegLinks; b++) s += J[3 + b] * x.l[b];
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Constraint rows in the quad engine.
//
// Every leg has a fixed set of ROW SLOTS; lane (L, h) OWNS the slots "h" of leg L: connect row h (x / z), in tier 1
// joint-limit slot h, and contact slot h (normal + tangent).  Slots that are not in use are inert rows (J = 0, R = 1,
// b = 0: their force stays exactly 0 and every term they contribute is an exact 0), so the PGS visits the real rows in
// the canonical order of mj_makeConstraint [EXT] -- connects (left x, z, right x, z), joint limits in dof order,
// contacts (pelvis sphere, then per leg the capsules thigh, shin, tarsus, toe, 'to' end before 'from' end) -- and
// an env's result does not depend on which tier its warp runs.
//   tier 0 (common regime: standing, squatting, walking): 2 connect slots + 2 toe-contact slots per leg, 12 rows
//   tier 1: + 2 joint-limit slots per leg, contact slots take any floor contact of the leg (leg 0 also the pelvis
//           sphere), 16 rows
//   tier 2: + 4 joint-limit slots per leg (every limited joint of a leg at its stop: hip, knee, tarsus, toe -- robots
//           in flight with the legs stretched, which is where random OSC accelerations send them), 20 rows.  Lane
//           (L, h) owns the limits number h and 2 + h of leg L; the columns keep mj_makeConstraint's dof order
//           (all of leg 0, then all of leg 1).
//   anything else (3+ contacts on one leg: robots lying on the floor) -> serial fallback.
constexpr int kRowSlots = 10;                      // row slots per leg reserved in PhysLayout::rowJ (tier 2 uses all)
template <int TIER>
struct Tier {
  static constexpr int NSL = TIER + 1;            // scalar rows per lane
  static constexpr int NS = 4 * NSL;              // scalar rows
  static constexpr int NR = NS + 8;               // rows
  static constexpr int LR = NR / 2;               // row slots per leg
  static constexpr int KO = NSL + 2;              // rows owned by a lane
  // leg-local slot of owned row k of half hh: k < NSL scalar kind k, then normal, tangent of contact hh
  static QUAD_FN int slot(int k, int hh) { return k < NSL ? 2 * k + hh : 2 * NSL + 2 * hh + (k - NSL); }
  // global column (= PGS order) of the scalar row of kind k (0 connect, 1.. limit) owned by lane c = 2 Lc + hh
  static QUAD_FN constexpr int scol(int k, int c) {
    return (TIER < 2 || k == 0) ? 4 * k + c : 4 + 4 * (c >> 1) + 2 * (k - 1) + (c & 1);
  }
  // ... and back: kind and owner lane of scalar column i
  static QUAD_FN constexpr int skind(int i) { return (TIER < 2 || i < 4) ? i >> 2 : 1 + (((i - 4) & 3) >> 1); }
  static QUAD_FN constexpr int sown(int i) { return (TIER < 2 || i < 4) ? i & 3 : 2 * ((i - 4) >> 2) + ((i - 4) & 1); }
  // global column (= PGS order) of leg-local slot t of leg Lc
  static QUAD_FN constexpr int col(int Lc, int t) {
    return t < 2 * NSL ? scol(t >> 1, 2 * Lc + (t & 1)) : NS + 2 * (2 * Lc + ((t - 2 * NSL) >> 1)) + ((t - 2 * NSL) & 1);
  }
};

// persistent per-env state (type T): survives from one step to the next, loaded / stored once per launch
struct StateLayout {
  static constexpr int q = 0, qd = 13, warm = 26, u = 39, op = 45, end = 57;
};
// per-step scratch of the physics step (type T); shares its storage with the controller's scratch (quad_ctrl.cuh).
// The constraint matrix A never touches shared memory: it goes from the dots of the assembly straight into the
// registers of the lanes that own the rows.
struct PhysLayout {
  static constexpr int ld1 = 0;                         // factor of M
  static constexpr int ld2 = ld1 + kLdSize;             // factor of M + h D (mj_Euler [EXT] implicit damping)
  static constexpr int rowJ = ld2 + kLdSize;            // [leg][slot][8], kRowSlots slots per leg reserved
  static constexpr int end = rowJ + 2 * kRowSlots * 8;
};

struct QStepStats {
  int nrows, sweeps;
  unsigned contact_mask;
};

// Serial fallback (lane 0 of the quad, thread-local rows): anything outside the two tiers goes through the
// thread-per-env engine.
template <typename T, typename TG>
QUAD_NOINLINE void serial_physics_step(const PlanarModel<T>& m, const PlanarModel<TG>& mg, SV<T> S, QStepStats* st) {
  T q[kNV], qd[kNV], w[kNV], u[kNU];
  for (int i = 0; i < kNV; i++) { q[i] = S[StateLayout::q + i]; qd[i] = S[StateLayout::qd + i]; w[i] = S[StateLayout::warm + i]; }
  for (int i = 0; i < kNU; i++) u[i] = S[StateLayout::u + i];
  Rows<T> rows;
  StepStats ss;
  physics_step(m, mg, q, qd, w, u, rows, &ss);
  for (int i = 0; i < kNV; i++) { S[StateLayout::q + i] = q[i]; S[StateLayout::qd + i] = qd[i]; S[StateLayout::warm + i] = w[i]; }
  st->nrows = ss.nrows; st->sweeps = ss.sweeps; st->contact_mask = ss.contact_mask;
}

template <typename T>
QUAD_FN void load_v8(SV<T> s, int L, V8<T>& x) {
  CASSIE_UNROLL
  for (int b = 0; b < 3; b++) x.b[b] = s[b];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) x.l[a] = s[3 + 5 * L + a];
}

// PGS sweeps (mj_solPGS [EXT]).  Lane l owns scalar row(s) l (+ 4 + l) and contact pair l; NCL = contact slots per
// leg that are in use anywhere in the warp (1: only the pairs of lanes 0 and 2 are visited).  A[k][c] = owned row k,
// column c in PGS order; b, f likewise.  Same update arithmetic and order of accumulation as
// constraint_solve_fast (planar_engine.cuh).  BRANCH FREE: every lane runs the projection of every slot on its own
// rows, only the owner's result is broadcast; "own" decisions are selects (a divergent if / else costs both sides:
// profiles/r2c_pd_env.txt, 13 of 32 threads active in the publish code of the first cut).  The owner's rows restart
// their residual (acc = b + A_own f_own) through the same FMA that adds the column to everybody else's rows, with the
// accumulator input swapped for b; row i + 1 of a pair skips column i (planar_engine.cuh).
// NCL = contact slots per leg in use anywhere in the warp: with 1, only the pairs of lanes 0 and 2 are visited (the others
// are inert in every env of the warp, their contribution is an exact zero).  A run-time, warp-uniform mask that skips
// every unused slot individually was measured slower (r2n: squat_osc -9 %, pd_env -18 %): the branches break up the
// schedule of the unrolled sweep.
template <int TIER, int NCL, typename T>
QUAD_FN int quad_pgs(const PlanarModel<T>& m, const Lane ln, const T (&A)[Tier<TIER>::KO][Tier<TIER>::NR],
                     const T (&b)[Tier<TIER>::KO], const T (&jar)[Tier<TIER>::KO], const T (&Rr)[Tier<TIER>::KO],
                     T (&fout)[Tier<TIER>::KO]) {
  typedef Tier<TIER> Q;
  constexpr int NSL = Q::NSL, NS = Q::NS, NR = Q::NR, KO = Q::KO;
  const int l = ln.ql;
  // diagonal blocks of the own rows
  T Ass[NSL], App = T(1), Apt = T(0), Att = T(1);
  CASSIE_UNROLL
  for (int k = 0; k < NSL; k++) {
    Ass[k] = T(1);
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) Ass[k] = c == l ? A[k][Q::scol(k, c)] : Ass[k];
  }
  CASSIE_UNROLL
  for (int p = 0; p < 4; p++) {
    if (p == l) { App = A[NSL][NS + 2 * p]; Apt = A[NSL + 1][NS + 2 * p]; Att = A[NSL + 1][NS + 2 * p + 1]; }
  }
  // ---- warm start (mj_constraintUpdate [EXT] on jar = J qacc_warmstart - aref)
  T f[KO];
  CASSIE_UNROLL
  for (int k = 0; k < NSL; k++) {
    const T fw = -jar[k] * Num<T>::rcp_(Rr[k]);
    f[k] = (k >= 1 && !(jar[k] < T(0))) ? T(0) : fw;   // limit rows: force only when violated
  }
  {
    const T jar1 = jar[NSL], jar2 = jar[NSL + 1];
    const T mu = m.con_mu / Num<T>::sqrt_(m.impratio);
    const T N = jar1 * mu, U1 = jar2 * m.con_mu, Tn = Num<T>::abs_(U1);
    const T D0 = Num<T>::rcp_(Rr[NSL]), D1 = Num<T>::rcp_(Rr[NSL + 1]);
    T f1, f2;
    if (N >= mu * Tn || (Tn <= T(0) && N >= T(0))) { f1 = T(0); f2 = T(0); }
    else if (mu * N + Tn <= T(0) || (Tn <= T(0) && N < T(0))) { f1 = -D0 * jar1; f2 = -D1 * jar2; }
    else {
      T den = mu * mu * (T(1) + mu * mu);
      const T Dm = D0 / (den > T(kMinVal) ? den : T(kMinVal));
      f1 = -Dm * (N - mu * Tn) * mu;
      f2 = -f1 / Tn * U1 * m.con_mu;
    }
    f[NSL] = f1; f[NSL + 1] = f2;
  }
  // all forces, replicated (warm-start cost and initial residuals)
  T fall[NR];
  CASSIE_UNROLL
  for (int k = 0; k < NSL; k++) {
    CASSIE_UNROLL
    for (int c = 0; c < 4; c++) fall[Q::scol(k, c)] = shfl(f[k], c);
  }
  CASSIE_UNROLL
  for (int p = 0; p < 4; p++) { fall[NS + 2 * p] = shfl(f[NSL], p); fall[NS + 2 * p + 1] = shfl(f[NSL + 1], p); }
  {
    T part = T(0);
    CASSIE_UNROLL
    for (int k = 0; k < KO; k++) {
      T sk = T(0);
      CASSIE_UNROLL
      for (int c = 0; c < NR; c++) sk += A[k][c] * fall[c];
      part += f[k] * (T(0.5) * sk + b[k]);
    }
    T cost = part + shx(part, 1);
    cost = cost + shx(cost, 2);
    if (cost > T(0)) {
      CASSIE_UNROLL
      for (int k = 0; k < KO; k++) f[k] = T(0);
      CASSIE_UNROLL
      for (int c = 0; c < NR; c++) fall[c] = T(0);
    }
  }
  T inv[KO];
  CASSIE_UNROLL
  for (int k = 0; k < NSL; k++) inv[k] = Num<T>::rcp_(Ass[k]);
  inv[NSL] = Num<T>::rcp_(App); inv[NSL + 1] = Num<T>::rcp_(Att);
  const T scale = T(1) / (m.meaninertia * T(kNV));
  const T mu = m.con_mu, inv_mu = T(1) / mu;
  // initial residuals: acc_i = b_i + sum_{c >= i} A_ic f_c (columns c < i arrive as this sweep publishes them)
  T acc[KO];
  CASSIE_UNROLL
  for (int k = 0; k < KO; k++) {
    const int row = k < NSL ? Q::scol(k, l) : NS + 2 * l + (k - NSL);
    T a = b[k];
    CASSIE_UNROLL
    for (int c = 0; c < NR; c++) a += (c >= row) ? A[k][c] * fall[c] : T(0);
    acc[k] = a;
  }
  T rden;
  {
    const T f1 = f[NSL], f2 = f[NSL + 1];
    const T denom = f1 * (App * f1 + Apt * f2) + f2 * (Apt * f1 + Att * f2);
    rden = denom >= T(kMinVal) ? Num<T>::rcp_(denom) : T(0);
  }
  int iter = 0, done_at = 0;
  bool done = false;
  T g[KO], fprev[KO];   // forces at the sweep where this env converged; forces after the previous sweep
  CASSIE_UNROLL
  for (int k = 0; k < KO; k++) { g[k] = T(0); fprev[k] = f[k]; }
  // The convergence test of mj_solPGS runs ONE SWEEP LATE: the quad-wide sum of a sweep's cost improvement (two
  // shuffles) is issued at the end of the sweep and consumed at the end of the next one, so that its latency sits under
  // the next sweep instead of on the loop's critical path.  A converged env is detected one sweep after the fact: its
  // forces were snapshotted (fprev), the extra sweep is discarded.
  T imp_prev = T(1e30);
  while (true) {
    T so[KO], cres[KO];   // forces at the start of the sweep (= "old" of the own slots), residuals seen by the own slots
    CASSIE_UNROLL
    for (int k = 0; k < KO; k++) { so[k] = f[k]; cres[k] = T(0); }
    CASSIE_UNROLL
    for (int i = 0; i < NS; i++) {   // scalar rows: connects unbounded, joint limits f >= 0
      const int ks = Q::skind(i);
      const bool own = Q::sown(i) == l;
      T fn_own = f[ks] - acc[ks] * inv[ks];
      if (ks >= 1) fn_own = Num<T>::max_(fn_own, T(0));
      const T fn = shfl(fn_own, Q::sown(i));
      cres[ks] = own ? acc[ks] : cres[ks];
      f[ks] = own ? fn : f[ks];
      CASSIE_UNROLL
      for (int k = 0; k < KO; k++) acc[k] = ((k == ks && own) ? b[k] : acc[k]) + A[k][i] * fn;
    }
    CASSIE_UNROLL
    for (int p = 0; p < 4; p += (NCL == 1 ? 2 : 1)) {   // elliptic contact: normal + one tangent, updated as a pair by lane p
      const int i = NS + 2 * p;
      const bool own = p == l;
      const T old0 = f[NSL], old1 = f[NSL + 1];
      const T res0 = acc[NSL];
      const T res1 = acc[NSL + 1] + Apt * old0;
      const T fa = Num<T>::max_(old0 - res0 * inv[NSL], T(0));
      const T x = Num<T>::max_(-(old0 * res0 + old1 * res1) * rden, T(-1));
      const T n0 = old0 < T(kMinVal) ? fa : old0 + x * old0;
      const T bc = (res1 - Att * old1 - Apt * old0) + Apt * n0;
      T v = -bc * inv[NSL + 1];
      const T lim = mu * n0;
      if (Num<T>::kExactConeTest) {
        const T vs = v * inv_mu;
        v = (vs * vs - n0 * n0 >= T(1e-10)) ? (v > T(0) ? lim : -lim) : v;
      } else {
        v = Num<T>::min_(Num<T>::max_(v, -lim), lim);
      }
      const T n1 = n0 < T(kMinVal) ? T(0) : v;
      const T F0 = shfl(n0, p), F1 = shfl(n1, p);
      cres[NSL] = own ? res0 : cres[NSL];
      cres[NSL + 1] = own ? res1 : cres[NSL + 1];
      f[NSL] = own ? n0 : f[NSL];
      f[NSL + 1] = own ? n1 : f[NSL + 1];
      CASSIE_UNROLL
      for (int k = 0; k < NSL; k++) { acc[k] += A[k][i] * F0; acc[k] += A[k][i + 1] * F1; }
      acc[NSL] = (own ? b[NSL] : acc[NSL]) + A[NSL][i] * F0;                  // own: A = A_pp
      acc[NSL] += A[NSL][i + 1] * F1;                                         // own: A = A_pt
      acc[NSL + 1] = (own ? b[NSL + 1] : acc[NSL + 1]) + A[NSL + 1][i] * (own ? T(0) : F0);   // own: row i + 1 skips column i
      acc[NSL + 1] += A[NSL + 1][i + 1] * F1;                                 // own: A = A_tt
    }
    // verdict on the PREVIOUS sweep (its improvement has had a whole sweep to arrive)
    if (!done && iter >= 1 && imp_prev * scale < m.tolerance) {
      done = true; done_at = iter;
      CASSIE_UNROLL
      for (int k = 0; k < KO; k++) g[k] = fprev[k];
    }
    iter++;
    if (!done && iter >= m.iterations) {
      done = true; done_at = iter;
      CASSIE_UNROLL
      for (int k = 0; k < KO; k++) g[k] = f[k];
    }
    if (!wany(!done)) break;
    // once per sweep, on the own rows: ray denominator of the own pair for the next sweep, cost improvement
    {
      const T f1 = f[NSL], f2 = f[NSL + 1];
      const T denom = f1 * (App * f1 + Apt * f2) + f2 * (Apt * f1 + Att * f2);
      rden = denom >= T(kMinVal) ? Num<T>::rcp_(denom) : T(0);
    }
    T improvement = T(0);
    CASSIE_UNROLL
    for (int k = 0; k < NSL; k++) {
      const T d = f[k] - so[k];
      improvement -= d * (T(0.5) * d * Ass[k] + cres[k]);
    }
    {
      const T d0 = f[NSL] - so[NSL], d1 = f[NSL + 1] - so[NSL + 1];
      improvement -= d0 * (T(0.5) * d0 * App + Apt * d1 + cres[NSL]) + d1 * (T(0.5) * d1 * Att + cres[NSL + 1]);
    }
    T imp = improvement + shx(improvement, 1);
    imp_prev = imp + shx(imp, 2);
    CASSIE_UNROLL
    for (int k = 0; k < KO; k++) fprev[k] = f[k];
  }
  CASSIE_UNROLL
  for (int k = 0; k < KO; k++) fout[k] = g[k];
  return done_at;
}

// Row building + constraint solve of one tier: returns qfrc_constraint = J^T f (leg split) and the sweep count.
// Both halves of a leg walk all of its slots (cheap), each writes the Jacobians of the slots it owns to rowJ and
// keeps R and aref of those in registers.
template <int TIER, bool CTA, typename T>
QUAD_FN void quad_constraints(const PlanarModel<T>& m, const Lane ln, SV<T> S, const LegKin<T>& k, const V8<T>& q, const V8<T>& qd,
                              const V8<T>& qs, const V8<T>& warm, T eq_rx, T eq_rz, const T (&cdist)[4][2], T sph_dist,
                              V8<T>& fc, int* sweeps_out, int* nrows_out, unsigned* mask_out) {
  typedef PhysLayout P;
  typedef Tier<TIER> Q;
  constexpr int NSL = Q::NSL, NR = Q::NR, LR = Q::LR, KO = Q::KO;
  const int L = ln.L, h = ln.h;
  T Rown[KO], b0[KO];    // regulariser and -aref of the owned rows
  T Jown[KO][8];
  int nlim = 0, ncon = 0;
  unsigned mask = 0;
  CASSIE_UNROLL
  for (int kk = 0; kk < KO; kk++) {
    Rown[kk] = T(1); b0[kk] = T(0);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) Jown[kk][c] = T(0);
  }
  T K, Bd;
  // ---- connect row h of this leg
  kb_from_solref(m, m.eq_solref, m.eq_solimp, K, Bd);
  {
    T ax, az, bx, bz;
    rot(k.c[kRod], k.s[kRod], m.eq_a1[L][0], m.eq_a1[L][1], ax, az);
    rot(k.c[kTarsus], k.s[kTarsus], m.eq_a2[L][0], m.eq_a2[L][1], bx, bz);
    T J1x[8], J1z[8], J2x[8], J2z[8];
    leg_point_jac<kRod>(m, L, k, ax, az, J1x, J1z);
    leg_point_jac<kTarsus>(m, L, k, bx, bz, J2x, J2z);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) Jown[0][c] = h ? J1z[c] - J2z[c] : J1x[c] - J2x[c];
    const T imp = impedance(m.eq_solimp, Num<T>::sqrt_(eq_rx * eq_rx + eq_rz * eq_rz));
    const T Rv = (T(1) - imp) * m.eq_diag[L] * Num<T>::rcp_(imp);
    Rown[0] = Rv > T(kMinVal) ? Rv : T(kMinVal);
    b0[0] = -(-Bd * dot8(Jown[0], qd) - K * imp * (h ? eq_rz : eq_rx));
  }
  // ---- joint limits of this leg, dof order: the n-th violated one goes to limit slot n (owned by half n)
  if (TIER) {
    kb_from_solref(m, m.lim_solref, m.lim_solimp, K, Bd);
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      const int j = 3 + 5 * L + a;
      if (m.has_limit[j]) {
        const T dlo = q.l[a] - m.lim_lo[j], dhi = m.lim_hi[j] - q.l[a];
        CASSIE_UNROLL
        for (int side = 0; side < 2; side++) {
          const T dist = side ? dhi : dlo;
          if (dist < T(0)) {
            CASSIE_UNROLL
            for (int kl = 1; kl < NSL; kl++) {
              if (nlim == 2 * (kl - 1) + h) {
                CASSIE_UNROLL
                for (int c = 0; c < 8; c++) Jown[kl][c] = T(0);
                const T sg = side ? T(-1) : T(1);
                CASSIE_UNROLL
                for (int c = 0; c < kLegLinks; c++) Jown[kl][3 + c] = c == a ? sg : T(0);
                const T imp = impedance(m.lim_solimp, dist);
                const T Rv = (T(1) - imp) * m.lim_diag[j] * Num<T>::rcp_(imp);
                Rown[kl] = Rv > T(kMinVal) ? Rv : T(kMinVal);
                b0[kl] = -(-Bd * (sg * qd.l[a]) - K * imp * dist);
              }
            }
            nlim++;
          }
        }
      }
    }
  }
  // ---- floor contacts of this leg (plane z = 0, normal +z), canonical order; the n-th active one goes to contact slot n
  kb_from_solref(m, m.con_solref, m.con_solimp, K, Bd);
  if (TIER) {
    if (L == 0 && !(sph_dist > T(0))) {
      T cx, cz, Jx[8], Jz[8];
      rot(k.c0, k.s0, m.sph_c[0], m.sph_c[1], cx, cz);
      leg_point_jac<-1>(m, L, k, cx, cz - m.sph_r - T(0.5) * sph_dist, Jx, Jz);
      if (ncon == h) {
        const T imp = impedance(m.con_solimp, sph_dist);
        const T Rv = (T(1) - imp) * m.sph_diag * Num<T>::rcp_(imp);
        Rown[NSL] = Rown[NSL + 1] = Rv > T(kMinVal) ? Rv : T(kMinVal);
        CASSIE_UNROLL
        for (int c = 0; c < 8; c++) { Jown[NSL][c] = Jz[c]; Jown[NSL + 1][c] = Jx[c]; }
        b0[NSL] = -(-Bd * dot8(Jz, qd) - K * imp * sph_dist);
        b0[NSL + 1] = -(-Bd * dot8(Jx, qd));
      }
      mask |= 1u << 2;
      ncon++;
    }
  }
#define CASSIE_QUAD_CONTACT(G)                                                                         \
  CASSIE_UNROLL                                                                                        \
  for (int e = 0; e < 2; e++) {                                                                        \
    const int cap = 4 * L + G;                                                                         \
    const T dist = cdist[G][e];                                                                        \
    if (!(dist > T(0))) {                                                                              \
      if (ncon == h) {                                                                                 \
        T ex, ez, Jx[8], Jz[8];                                                                        \
        const T* ep = e ? m.cap_from[cap] : m.cap_to[cap];                                             \
        rot(k.c[G], k.s[G], ep[0], ep[1], ex, ez);                                                     \
        leg_point_jac<G>(m, L, k, ex, ez - m.cap_r[cap] - T(0.5) * dist, Jx, Jz);                      \
        const T imp = impedance(m.con_solimp, dist);                                                   \
        const T Rv = (T(1) - imp) * m.cap_diag[cap] * Num<T>::rcp_(imp);                               \
        Rown[NSL] = Rown[NSL + 1] = Rv > T(kMinVal) ? Rv : T(kMinVal);                                 \
        CASSIE_UNROLL                                                                                  \
        for (int c = 0; c < 8; c++) { Jown[NSL][c] = Jz[c]; Jown[NSL + 1][c] = Jx[c]; }                \
        b0[NSL] = -(-Bd * dot8(Jz, qd) - K * imp * dist);                                              \
        b0[NSL + 1] = -(-Bd * dot8(Jx, qd));                                                           \
      }                                                                                                \
      mask |= 1u << (2 * (2 + cap) + e);                                                               \
      ncon++;                                                                                          \
    }                                                                                                  \
  }
  if (TIER) {
    CASSIE_QUAD_CONTACT(kThigh)
    CASSIE_QUAD_CONTACT(kKnee)
    CASSIE_QUAD_CONTACT(kTarsus)
  }
  CASSIE_QUAD_CONTACT(kToe)
#undef CASSIE_QUAD_CONTACT
  // ---- publish the Jacobians of the owned slots
  CASSIE_UNROLL
  for (int kk = 0; kk < KO; kk++) {
    const SV<T> dst = S.at(P::rowJ + (L * kRowSlots + Q::slot(kk, h)) * 8);
    CASSIE_UNROLL
    for (int c = 0; c < 8; c++) dst[c] = Jown[kk][c];
  }
  const bool narrow = !(CTA ? cta_any(ncon > 1) : wany(ncon > 1));
  wsync();

  // ---- b = J qacc_smooth - aref, jar = J qacc_warmstart - aref of the owned rows
  T bown[KO], jar[KO];
  CASSIE_UNROLL
  for (int kk = 0; kk < KO; kk++) { bown[kk] = b0[kk] + dot8(Jown[kk], qs); jar[kk] = b0[kk] + dot8(Jown[kk], warm); }

  // ---- A = J M^-1 J^T + R.  Half h solves for the slots of leg h as right-hand sides (LR / 2 at a time: independent
  // chains, the solve is latency bound); lane (L, h) dots the results with all slots of leg L, keeps the rows it owns
  // and hands the partner's rows over by one shuffle each.
  T A[KO][NR];
  CASSIE_UNROLL
  for (int t0 = 0; t0 < LR; t0 += LR / 2) {
    constexpr int NB = LR / 2;
    V8<T> x[NB];
    CASSIE_UNROLL
    for (int r = 0; r < NB; r++) {
      const SV<T> src = S.at(P::rowJ + (h * kRowSlots + t0 + r) * 8);
      CASSIE_UNROLL
      for (int bb = 0; bb < 3; bb++) x[r].b[bb] = src[bb];
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) x[r].l[a] = (L == h) ? src[3 + a] : T(0);
    }
    quad_solve<NB>(S.at(P::ld1), L, x);
    CASSIE_UNROLL
    for (int kk = 0; kk < KO; kk++) {
      // the partner's row of the same kind, from shared memory
      const int sp = Q::slot(kk, 1 - h), so = Q::slot(kk, h);
      T Jp[8];
      const SV<T> src = S.at(P::rowJ + (L * kRowSlots + sp) * 8);
      CASSIE_UNROLL
      for (int c = 0; c < 8; c++) Jp[c] = src[c];
      CASSIE_UNROLL
      for (int r = 0; r < NB; r++) {
        const int t = t0 + r;
        T vo = dot8(Jown[kk], x[r]);
        const T vp = dot8(Jp, x[r]);
        if (L == h && so == t) vo += Rown[kk];   // the diagonal: own row = this right-hand side
        // the partner's diagonal is finished by the partner: it adds its own R below when it receives the entry
        const T got = shx(vp, 1);                // partner's dot of ITS right-hand side (leg 1 - h, slot t) with MY row
        const bool pdiag = (L == 1 - h) && (so == t);
        const T gotd = pdiag ? got + Rown[kk] : got;
        A[kk][Q::col(0, t)] = h ? gotd : vo;
        A[kk][Q::col(1, t)] = h ? vo : gotd;
      }
    }
  }

  // ---- PGS
  phase_sync<2>();
  T fown[KO];
  int sweeps;
  if (narrow) sweeps = quad_pgs<TIER, 1>(m, ln, A, bown, jar, Rown, fown);
  else sweeps = quad_pgs<TIER, 2>(m, ln, A, bown, jar, Rown, fown);

  // ---- qfrc_constraint = J^T f over the slots of this leg (own rows + the partner's)
  phase_sync<2>();
  {
    T pb[3] = {T(0), T(0), T(0)};
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) fc.l[a] = T(0);
    CASSIE_UNROLL
    for (int kk = 0; kk < KO; kk++) {
      const T fo = fown[kk], fp = shx(fown[kk], 1);
      const SV<T> src = S.at(P::rowJ + (L * kRowSlots + Q::slot(kk, 1 - h)) * 8);
      CASSIE_UNROLL
      for (int bb = 0; bb < 3; bb++) pb[bb] += Jown[kk][bb] * fo + src[bb] * fp;
      CASSIE_UNROLL
      for (int a = 0; a < kLegLinks; a++) fc.l[a] += Jown[kk][3 + a] * fo + src[3 + a] * fp;
    }
    CASSIE_UNROLL
    for (int bb = 0; bb < 3; bb++) fc.b[bb] = sum_legs(pb[bb]);
  }
  *sweeps_out = sweeps;
  const int nl_o = shx(nlim, 2), nc_o = shx(ncon, 2);
  *nrows_out = 4 + (nlim + nl_o) + 2 * (ncon + nc_o);
  *mask_out = mask | shx(mask, 2);
}

// One mj_step [EXT] (Cassie2d.cpp:92) of the env owned by this quad.  State (q, qd, warm start) and the control u
// live in the state block St and are updated in place; S = scratch (PhysLayout).  T = working type, TG = type of the
// position pass (double in the fp32 build: planar_engine.cuh physics_step explains why).
// CTA = choose the constraint tier per CTA instead of per warp (kernels whose envs spread over the tiers: env step /
// rollout with arbitrary actions); the squatting kernels, whose envs all sit in tier 0, keep the barrier-free warp vote.
template <bool CTA = false, typename T, typename TG>
QUAD_FN void quad_physics_step(const PlanarModel<T>& m, const PlanarModel<TG>& mg, const Lane ln, SV<T> St, SV<T> S,
                               QStepStats* st) {
  typedef PhysLayout P;
  typedef StateLayout X;
  const int L = ln.L, h = ln.h;
  V8<T> q, qd;
  load_v8(St.at(X::q), L, q);
  load_v8(St.at(X::qd), L, qd);
  LegKin<T> k;
  // ---- position pass in TG: angles -> sin/cos -> pivots -> constraint violations
  T eq_rx, eq_rz, cdist[4][2], sph_dist;
  {
    V8<TG> qg;
    CASSIE_UNROLL
    for (int b = 0; b < 3; b++) qg.b[b] = (TG)q.b[b];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) qg.l[a] = (TG)q.l[a];
    LegKin<TG> kg;
    leg_fk_positions(mg, L, qg, kg);
    {
      TG ax, az, bx, bz;
      rot(kg.c[kRod], kg.s[kRod], mg.eq_a1[L][0], mg.eq_a1[L][1], ax, az);
      rot(kg.c[kTarsus], kg.s[kTarsus], mg.eq_a2[L][0], mg.eq_a2[L][1], bx, bz);
      eq_rx = (T)((kg.dx[kRod] - kg.dx[kKnee] - kg.dx[kTarsus]) + (ax - bx));
      eq_rz = (T)((kg.dz[kRod] - kg.dz[kKnee] - kg.dz[kTarsus]) + (az - bz));
    }
    const TG height = qg.b[1] - mg.pel_ref[1] + mg.pel_org[1];
    {
      TG cx, cz;
      rot(kg.c0, kg.s0, mg.sph_c[0], mg.sph_c[1], cx, cz);
      sph_dist = (T)(height + cz - mg.sph_r);
    }
    CASSIE_UNROLL
    for (int g = 0; g < 4; g++) {
      const int cap = 4 * L + g;
      CASSIE_UNROLL
      for (int e = 0; e < 2; e++) {
        TG ex, ez;
        const TG* ep = e ? mg.cap_from[cap] : mg.cap_to[cap];
        rot(kg.c[g], kg.s[g], ep[0], ep[1], ex, ez);
        cdist[g][e] = (T)(height + kg.pz[g] + ez - mg.cap_r[cap]);
      }
    }
    cast_leg_positions(kg, k);
  }
  // ---- which tier?  Counted per leg; comparisons are written so that a non-finite state counts as "nothing active":
  // such an env stays on the common path (garbage in, garbage out; the env-level guard resets it) instead of dragging
  // its warp into another tier
  int nlim = 0, ncon_other = 0, ncon = 0;
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) {
    const int j = 3 + 5 * L + a;
    if (m.has_limit[j]) nlim += (q.l[a] - m.lim_lo[j] < T(0)) + (m.lim_hi[j] - q.l[a] < T(0));
  }
  CASSIE_UNROLL
  for (int g = 0; g < 3; g++) ncon_other += (cdist[g][0] <= T(0)) + (cdist[g][1] <= T(0));
  if (L == 0) ncon_other += sph_dist <= T(0);
  ncon = ncon_other + (cdist[kToe][0] <= T(0)) + (cdist[kToe][1] <= T(0));
  bool tier1 = nlim > 0 || ncon_other > 0;
  bool tier2 = nlim > 2;
  bool general = nlim > 4 || ncon > 2;
#ifdef CASSIE_HOST_HARNESS
  general = general || cassie_force_general_path;
  tier1 = tier1 || cassie_force_tier1;
  tier2 = tier2 || cassie_force_tier2;
#endif
  // The tier is chosen per WARP (the shuffles of a tier name all 32 lanes).  The two tiers agree bit for bit on the envs
  // both can handle (inert rows contribute exact zeros), and an env outside both ("general": robots lying on the floor)
  // walks through tier 1 with its warp without committing the result and is then stepped by lane 0 of its quad through
  // the thread-per-env engine -- so an env's result never depends on its warp mates.
  general = qany(general);
  const bool any_general = wany(general);
  // ... and per CTA where several warps run in lock step: seven warps in three different tiers would each pull their own
  // unrolled code through the instruction cache (OSC-action rollout, profiles/r2as_cta_tier.txt)
  const bool use_t2 = CTA ? cta_any(tier2) : wany(tier2);
  const bool use_t1 = use_t2 || (CTA ? cta_any(tier1 || general) : wany(tier1 || general));
  leg_fk_velocities(m, L, qd, k);

  // ---- mass matrix, bias, both factorisations (half 0: M, half 1: M + h D)
  V8<T> fs;   // qfrc_smooth = passive - bias + actuator
  {
    LegM<T> M;
    leg_mass_matrix(m, L, k, M);
    leg_bias_forces(m, L, k, fs);
    const T hd = h ? m.timestep : T(0);
    M.bb[0] += hd * m.damping[0]; M.bb[2] += hd * m.damping[1]; M.bb[5] += hd * m.damping[2];
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) M.d[a] += hd * m.damping[3 + 5 * L + a];
    leg_factor(M);
    store_factor(S.at(h ? P::ld2 : P::ld1), L, M, L == 0);
  }
  CASSIE_UNROLL
  for (int b = 0; b < 3; b++) fs.b[b] = -fs.b[b] - m.damping[b] * qd.b[b];
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) fs.l[a] = -fs.l[a] - m.damping[3 + 5 * L + a] * qd.l[a];
  CASSIE_UNROLL
  for (int a = 0; a < kNU; a++) {
    T c = St[X::u + a];
    c = c < m.act_lo[a] ? m.act_lo[a] : (c > m.act_hi[a] ? m.act_hi[a] : c);
    CASSIE_UNROLL
    for (int i = 0; i < kLegLinks; i++)
      if (m.act_dof[a] == 3 + 5 * L + i) fs.l[i] += m.act_gear[a] * c;
  }
  wsync();   // both factors are in shared memory
  phase_sync<2>();

  // ---- qacc_smooth
  V8<T> qs = fs;
  {
    V8<T> x[1] = {qs};
    quad_solve<1>(S.at(P::ld1), L, x);
    qs = x[0];
  }
  V8<T> warm;
  load_v8(St.at(X::warm), L, warm);

  // ---- constraint rows, A, PGS
  V8<T> fc;
  int sweeps, nrows;
  unsigned mask;
  if (use_t2) quad_constraints<2, CTA>(m, ln, S, k, q, qd, qs, warm, eq_rx, eq_rz, cdist, sph_dist, fc, &sweeps, &nrows, &mask);
  else if (use_t1) quad_constraints<1, CTA>(m, ln, S, k, q, qd, qs, warm, eq_rx, eq_rz, cdist, sph_dist, fc, &sweeps, &nrows, &mask);
  else quad_constraints<0, CTA>(m, ln, S, k, q, qd, qs, warm, eq_rx, eq_rz, cdist, sph_dist, fc, &sweeps, &nrows, &mask);

  // ---- qacc = qacc_smooth + M^-1 qfrc_constraint, mj_Euler [EXT]: (M + h D) qacc' = qfrc_smooth + qfrc_constraint
  V8<T> x2[1] = {fc};
  quad_solve<1>(S.at(P::ld1), L, x2);
  const T hh = m.timestep;
  V8<T> wnew, rhs;
  CASSIE_UNROLL
  for (int b = 0; b < 3; b++) { wnew.b[b] = qs.b[b] + x2[0].b[b]; rhs.b[b] = fs.b[b] + fc.b[b]; }
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) { wnew.l[a] = qs.l[a] + x2[0].l[a]; rhs.l[a] = fs.l[a] + fc.l[a]; }
  V8<T> x3[1] = {rhs};
  quad_solve<1>(S.at(P::ld2), L, x3);
  CASSIE_UNROLL
  for (int b = 0; b < 3; b++) { qd.b[b] += hh * x3[0].b[b]; q.b[b] += hh * qd.b[b]; }
  CASSIE_UNROLL
  for (int a = 0; a < kLegLinks; a++) { qd.l[a] += hh * x3[0].l[a]; q.l[a] += hh * qd.l[a]; }
  if (h == 0 && !general) {
    if (L == 0) {
      CASSIE_UNROLL
      for (int b = 0; b < 3; b++) { St[X::q + b] = q.b[b]; St[X::qd + b] = qd.b[b]; St[X::warm + b] = wnew.b[b]; }
    }
    CASSIE_UNROLL
    for (int a = 0; a < kLegLinks; a++) {
      St[X::q + 3 + 5 * L + a] = q.l[a]; St[X::qd + 3 + 5 * L + a] = qd.l[a]; St[X::warm + 3 + 5 * L + a] = wnew.l[a];
    }
  }
  st->nrows = nrows;
  st->sweeps = sweeps;
  st->contact_mask = mask;
  if (any_general) {
    wsync();
    if (general && ln.ql == 0) serial_physics_step(m, mg, St, st);
  }
  wsync();
}

}  // namespace quad
}  // namespace cassie
