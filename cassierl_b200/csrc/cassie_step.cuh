// One legacy Step* call of the Cassie2d facade for one env (CassieRL/cassierl
// src/Cassie2d/Cassie2d.cpp:86-209): set the controller (RBDL) state, evaluate the control law,
// take one physics step.  Also the Python-side pieces that the batched engine moves on device:
// the squatting control laws (rllab/envs/cassie2d.py:263-331, squatting.py:8-16) and the
// observation / reward / termination arithmetic (cassie_stand2d.py:86-137, cassie2d.py:97-225,
// cassie2d_structs.py:57-75).
#pragma once
#include "controllers.cuh"
#include "osc_qp.cuh"

namespace cassie {

// action sizes per mode: torque 6, pd 6, jacobian 6 (Fx,Fz,My per foot), osc 7
CASSIE_HD constexpr int action_dim(int mode) { return mode == kModeOsc ? 7 : 6; }

// `op` (may be null) receives the operational-space quantities of the state at the START of
// this step -- what GetOperationalSpaceState reports after the step.
// MODE is a template parameter so that each control mode gets its own register allocation.
// The controller model may be held in a wider type TC than the physics (T): the fp32 build runs
// the OSC controller in double (its QP is ill-conditioned, osc_qp.cuh), everything else in T.
// mg / mcg = the physics / controller model in the type of the position pass (planar_engine.cuh
// physics_step): angles -> sin/cos -> pivots always run in TG (double in the fp32 build; its sincos is
// branch-free), the result is cast to the working type.
template <int MODE, typename T, typename TG, typename TC>
CASSIE_HD void controller_step(const PlanarModel<T>& mp, const PlanarModel<TG>& mg, const PlanarModel<TC>& mc,
                               const PlanarModel<TG>& mcg, T q[kNV], T qd[kNV],
                               T warm[kNV], const T* act, Rows<T>& rows, T u[kNU], OpState<T>* op,
                               StepStats* st, OscStats* qst = nullptr, unsigned* qp_set = nullptr) {
  if (MODE == kModeTorque) {
    CASSIE_UNROLL
    for (int i = 0; i < kNU; i++) u[i] = act[i];
  } else if (MODE == kModePd) {
    CASSIE_UNROLL
    for (int a = 0; a < kNU; a++) {  // Cassie2d.cpp:96-112: gains are in ctrl units
      T qj = T(0), vj = T(0);
      CASSIE_UNROLL
      for (int i = 3; i < kNV; i++)
        if (mp.act_dof[a] == i) { qj = q[i]; vj = qd[i]; }
      u[a] = T(10) * (act[a] - qj) + T(5) * (T(0) - vj);
    }
  }
  if (op || MODE >= kModeJacobian) {
    TC qc[kNV], qdc[kNV];
    CASSIE_UNROLL
    for (int i = 0; i < kNV; i++) { qc[i] = (TC)q[i]; qdc[i] = (TC)qd[i]; }
    Kin<TC> kc;
    {
      TG qg[kNV];
      CASSIE_UNROLL
      for (int i = 0; i < kNV; i++) qg[i] = (TG)q[i];
      Kin<TG> kg;
      fk_positions(mcg, qg, kg);
      cast_kin_positions(kg, kc);
    }
    fk_velocities(mc, qdc, kc);
    if (op) {
      OpState<TC> oc;
      op_state_from_kin(mc, kc, qc, oc);
      CASSIE_UNROLL
      for (int i = 0; i < 4; i++) { op->body[i] = (T)oc.body[i]; op->left[i] = (T)oc.left[i]; op->right[i] = (T)oc.right[i]; }
    }
    if (MODE >= kModeJacobian) {
      TC ac[7], uc[kNU];
      CASSIE_UNROLL
      for (int i = 0; i < action_dim(MODE); i++) ac[i] = (TC)act[i];
      if (MODE == kModeJacobian) jacobian_control(mc, kc, qdc, ac, uc);
      else {
#if defined(__CUDA_ARCH__) && !defined(CASSIE_NO_OSC_OVERLAY)
        static_assert(sizeof(Rows<T>) >= kOscWsDoubles * sizeof(double) + sizeof(CtrlDyn<TC>) + 4 * kQpN * sizeof(double), "constraint rows too small to host the QP scratch");
        osc_control(mc, kc, qdc, ac, uc, qst, qp_set, reinterpret_cast<double*>(&rows));
        asm volatile("" ::: "memory");  // the rows are re-typed below: no reordering of the physics stores above this
#else
        osc_control(mc, kc, qdc, ac, uc, qst, qp_set);
#endif
      }
      CASSIE_UNROLL
      for (int i = 0; i < kNU; i++) u[i] = (T)uc[i];
    }
  }
  physics_step(mp, mg, q, qd, warm, u, rows, st);
}
template <typename T, typename TG, typename TC>
CASSIE_HD void controller_step_dyn(const PlanarModel<T>& mp, const PlanarModel<TG>& mg, const PlanarModel<TC>& mc,
                                   const PlanarModel<TG>& mcg, int mode, T q[kNV], T qd[kNV],
                                   T warm[kNV], const T* act, Rows<T>& rows, T u[kNU], OpState<T>* op,
                                   StepStats* st, OscStats* qst = nullptr, unsigned* qp_set = nullptr) {
  if (mode == kModeTorque) controller_step<kModeTorque>(mp, mg, mc, mcg, q, qd, warm, act, rows, u, op, st, qst, qp_set);
  else if (mode == kModePd) controller_step<kModePd>(mp, mg, mc, mcg, q, qd, warm, act, rows, u, op, st, qst, qp_set);
  else if (mode == kModeJacobian) controller_step<kModeJacobian>(mp, mg, mc, mcg, q, qd, warm, act, rows, u, op, st, qst, qp_set);
  else controller_step<kModeOsc>(mp, mg, mc, mcg, q, qd, warm, act, rows, u, op, st, qst, qp_set);
}

// standing_controller_jacobian (cassie2d.py:297-331): s = GetOperationalSpaceState array
template <typename T>
CASSIE_HD void squat_jacobian_action(const T s[18], T zt, T zdt, T f[6]) {
  const T xt = (s[6] + s[12]) / T(2);
  const T fx = T(200) * (xt - s[0]) + T(50) * (T(0) - s[3]);
  T fz = T(0.5) * T(9.806) * T(31.0) + T(200) * (zt - s[1]) + T(50) * (zdt - s[4]);
  const T my = T(100) * (T(0) - s[2]) + T(10) * (T(0) - s[5]);
  if (fz < T(0)) fz = T(0);
  f[0] = fx; f[1] = fz; f[2] = my; f[3] = fx; f[4] = fz; f[5] = my;
}
// standing_controller_osc (cassie2d.py:263-295)
template <typename T>
CASSIE_HD void squat_osc_action(const T s[18], T zt, T zdt, T a[7]) {
  a[2] = T(0); a[3] = T(100) * (T(-5e-3) - s[7]);
  a[4] = T(0); a[5] = T(100) * (T(-5e-3) - s[13]);
  const T xt = (s[6] + s[12]) / T(2);
  a[0] = T(100) * (xt - s[0]) + T(20) * (T(0) - s[3]);
  a[1] = T(100) * (zt - s[1]) + T(20) * (zdt - s[4]);
  a[6] = T(20) * (T(0) - s[2]) + T(10) * (T(0) - s[5]);
}

// operational_state_array_to_pos_invariant_array (cassie2d_structs.py:68-75): obs[0:17]
template <typename T>
CASSIE_HD void pos_invariant_obs(const T s[18], T o[17]) {
  CASSIE_UNROLL
  for (int i = 0; i < 17; i++) o[i] = s[i + 1];
  o[5] -= s[0];
  o[11] -= s[0];
}

enum TaskKind { kTaskStand = 0, kTaskImitate = 1 };

// cassie_stand2d.py:119-133.  `act`/`adim` = the policy action of this step.
template <typename T>
CASSIE_HD void stand_reward(const T s[18], const T o[17], const T* act, int adim, T& r, int& done) {
  T rr = T(0);
  const T dz = T(0.9) - s[1];
  rr -= T(2) * dz * dz;
  const T c = (o[5] + o[11]) / T(2);
  rr -= T(2) * c * c;
  rr += T(1);
  T a2 = T(0);
  for (int i = 0; i < adim; i++) a2 += act[i] * act[i];
  rr -= T(0.001) * a2;
  r = rr;
  done = s[1] < T(0.5) ? 1 : 0;
}

// cassie2d.py:196-223.  ref9 = reference (x, z, pitch, l-hip, l-knee, l-toe, r-hip, r-knee,
// r-toe) at the episode time; jsum = sum of the six actuated joint angles of `qstate` (frozen
// at the reset pose in the reference because GetGeneralState is commented out, :125).
template <typename T>
CASSIE_HD void imitate_reward(const T s[18], const T ref9[9], T jsum, T& r, int& done) {
  T j = jsum;
  CASSIE_UNROLL
  for (int i = 3; i < 9; i++) j -= ref9[i];
  j = Num<T>::exp_(-(j * j));
  T p = s[0] + s[1] - (ref9[0] + ref9[1]);
  p = Num<T>::exp_(-(p * p));
  T o = s[2] - ref9[2];
  o = Num<T>::exp_(-(o * o));
  r = T(0.5) * j + T(0.3) * p + T(0.1) * o;
  done = (s[1] < T(0.6) || s[1] > T(1.2) || r < T(0.6)) ? 1 : 0;
}

}  // namespace cassie
