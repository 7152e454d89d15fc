/* TEST INFRASTRUCTURE -- internal structs of the fp64 CPU oracle (see cassie_oracle.h). */
#ifndef CASSIE_ORACLE_INTERNAL_H_
#define CASSIE_ORACLE_INTERNAL_H_
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "cassie_oracle.h"

#define ORC_MINVAL 1e-15

struct orc_model {
  int nbody, nv, nq, ngeom, nsite, neq, nu;
  int body_parent[ORC_NB], body_jntadr[ORC_NB], body_jntnum[ORC_NB];
  double body_pos[ORC_NB][3], body_mat[ORC_NB][9], body_ipos[ORC_NB][3], body_mass[ORC_NB];
  double body_inertia[ORC_NB][9];
  int jnt_body[ORC_NV], jnt_type[ORC_NV], jnt_limited[ORC_NV], jnt_qadr[ORC_NV];
  double jnt_axis[ORC_NV][3], jnt_pos[ORC_NV][3], jnt_ref[ORC_NV], jnt_range[ORC_NV][2];
  double jnt_damping[ORC_NV], jnt_armature[ORC_NV], jnt_solref[ORC_NV][2], jnt_solimp[ORC_NV][5];
  int geom_body[ORC_NG], geom_type[ORC_NG], geom_contype[ORC_NG], geom_conaffinity[ORC_NG], geom_condim[ORC_NG];
  double geom_pos[ORC_NG][3], geom_mat[ORC_NG][9], geom_size[ORC_NG][2], geom_friction[ORC_NG][3];
  double geom_solref[ORC_NG][2], geom_solimp[ORC_NG][5], geom_margin[ORC_NG], geom_gap[ORC_NG];
  int site_body[ORC_NS];
  double site_pos[ORC_NS][3];
  int eq_body1[ORC_NEQ], eq_body2[ORC_NEQ];
  double eq_anchor1[ORC_NEQ][3], eq_anchor2[ORC_NEQ][3], eq_solref[ORC_NEQ][2], eq_solimp[ORC_NEQ][5];
  int act_jnt[ORC_NU], act_limited[ORC_NU];
  double act_gear[ORC_NU], act_range[ORC_NU][2];
  double timestep, tolerance, impratio, gravity[3];
  int iterations;
  /* compiled */
  unsigned char anc[ORC_NB][ORC_NV]; /* dof j moves body b */
  double qpos0[ORC_NQ], body_invweight0[ORC_NB][2], dof_invweight0[ORC_NV], meaninertia;
};

struct orc_kin {
  double q[ORC_NQ], qd[ORC_NV];
  double xpos[ORC_NB][3], xmat[ORC_NB][9], xipos[ORC_NB][3], Iw[ORC_NB][9];
  double S[ORC_NV][6], V[ORC_NB][6], A[ORC_NB][6];
};

typedef struct {
  double dist, pos[3], frame[9], friction[5], solref[2], solimp[5], includemargin;
  int geom1, geom2, dim, slot; /* slot: canonical (2*geom2 + capsule end) */
} orc_contact;

struct orc_data {
  int nv, nq;
  double qpos[ORC_NQ], qvel[ORC_NV], qacc_warmstart[ORC_NV], time;
  double ctrl[ORC_NU];
  orc_kin kin;
  double M[ORC_NV * ORC_NV], L[ORC_NV * ORC_NV];
  double qfrc_bias[ORC_NV], qfrc_passive[ORC_NV], qfrc_actuator[ORC_NV], qfrc_constraint[ORC_NV];
  double qacc_smooth[ORC_NV], qacc[ORC_NV];
  int ncon, nefc, ne, solver_iter;
  orc_contact contact[ORC_MAXCON];
  double efc_J[ORC_MAXEFC * ORC_NV], efc_pos[ORC_MAXEFC], efc_margin[ORC_MAXEFC];
  double efc_diagApprox[ORC_MAXEFC], efc_R[ORC_MAXEFC], efc_D[ORC_MAXEFC], efc_aref[ORC_MAXEFC];
  double efc_vel[ORC_MAXEFC], efc_b[ORC_MAXEFC], efc_force[ORC_MAXEFC], efc_KBIP[ORC_MAXEFC][4];
  double efc_AR[ORC_MAXEFC * ORC_MAXEFC];
  int efc_type[ORC_MAXEFC], efc_id[ORC_MAXEFC], efc_dim[ORC_MAXEFC];
  double min_capsule_gap; /* smallest leg-leg capsule surface distance seen (must stay > 0) */
};

/* shared helpers (cassie_oracle.c) */
void orc_mat3_mulv(double r[3], const double M[9], const double v[3]);
void orc_mat3_tmulv(double r[3], const double M[9], const double v[3]);
void orc_mat3_mul(double R[9], const double A[9], const double B[9]);
void orc_cross(double r[3], const double a[3], const double b[3]);
double orc_dot3(const double a[3], const double b[3]);
int orc_chol(int n, const double* A, double* L);
void orc_chol_solve(int n, const double* L, const double* b, double* x);
void orc_jac_world(const orc_model* m, const orc_kin* k, int body, const double P[3],
                   double* jacp, double* jacr);

#endif
