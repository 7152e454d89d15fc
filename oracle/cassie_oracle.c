/* TEST INFRASTRUCTURE -- fp64 CPU oracle, see cassie_oracle.h.  PARITY UNPINNED.
 * Part 1: model, kinematics, dynamics, constraints, PGS, Euler  (restates mj_step [EXT],
 *         call sites CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:62,81,92,115,174,206).
 * Part 2 (cassie_oracle_ctrl.c): RBDL-equivalent getters, controllers, Cassie2d facade.
 */
#include "cassie_oracle_internal.h"
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ small vector helpers */
void orc_mat3_mulv(double r[3], const double M[9], const double v[3]) {
  double a = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  double b = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  double c = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = a; r[1] = b; r[2] = c;
}
void orc_mat3_tmulv(double r[3], const double M[9], const double v[3]) {
  double a = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  double b = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  double c = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = a; r[1] = b; r[2] = c;
}
void orc_mat3_mul(double R[9], const double A[9], const double B[9]) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(R, T, sizeof(T));
}
void orc_cross(double r[3], const double a[3], const double b[3]) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
double orc_dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double a[3]) { return sqrt(orc_dot3(a, a)); }
static double normalize3(double a[3]) {
  double n = norm3(a);
  if (n < ORC_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return n; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
/* rotation matrix about unit axis by angle (Rodrigues) */
static void axis_angle(double R[9], const double ax[3], double ang) {
  double c = cos(ang), s = sin(ang), t = 1 - c;
  double x = ax[0], y = ax[1], z = ax[2];
  R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
  R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
  R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}

/* spatial: motion cross  [w;v] x [w2;v2] */
static void crm(double r[6], const double a[6], const double b[6]) {
  double t1[3], t2[3], t3[3];
  orc_cross(t1, a, b);
  orc_cross(t2, a, b + 3);
  orc_cross(t3, a + 3, b);
  r[0] = t1[0]; r[1] = t1[1]; r[2] = t1[2];
  r[3] = t2[0] + t3[0]; r[4] = t2[1] + t3[1]; r[5] = t2[2] + t3[2];
}
/* spatial: force cross  [w;v] x* [n;f] = [w x n + v x f ; w x f] */
static void crf(double r[6], const double a[6], const double b[6]) {
  double t1[3], t2[3], t3[3];
  orc_cross(t1, a, b);
  orc_cross(t2, a + 3, b + 3);
  orc_cross(t3, a, b + 3);
  r[0] = t1[0] + t2[0]; r[1] = t1[1] + t2[1]; r[2] = t1[2] + t2[2];
  r[3] = t3[0]; r[4] = t3[1]; r[5] = t3[2];
}
/* spatial inertia (mass m, com c, rotational inertia about com Ic, world axes) applied to a
 * motion vector about the world origin: returns [n_O ; p] */
static void inertia_apply(double r[6], double m, const double c[3], const double Ic[9], const double V[6]) {
  double wc[3], p[3], Iw[3], cp[3];
  orc_cross(wc, V, c);
  for (int i = 0; i < 3; i++) p[i] = m * (V[3 + i] + wc[i]);
  orc_mat3_mulv(Iw, Ic, V);
  orc_cross(cp, c, p);
  for (int i = 0; i < 3; i++) { r[i] = Iw[i] + cp[i]; r[3 + i] = p[i]; }
}

/* dense SPD Cholesky A = L L^T (lower, row-major n x n); returns 0 ok */
int orc_chol(int n, const double* A, double* L) {
  memset(L, 0, sizeof(double) * n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (s <= 0) return -1;
        L[i * n + i] = sqrt(s);
      } else
        L[i * n + j] = s / L[j * n + j];
    }
  return 0;
}
void orc_chol_solve(int n, const double* L, const double* b, double* x) {
  double y[64];
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------ model construction */
orc_model* orc_model_new(void) {
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  /* world body */
  m->nbody = 1;
  m->body_parent[0] = 0;
  m->body_mat[0][0] = m->body_mat[0][4] = m->body_mat[0][8] = 1.0;
  m->timestep = 0.002; m->iterations = 100; m->tolerance = 1e-8; m->impratio = 1.0;
  m->gravity[2] = -9.81;
  return m;
}
void orc_model_free(orc_model* m) { free(m); }

int orc_add_body(orc_model* m, int parent, const double pos[3], const double mat[9],
                 const double ipos[3], double mass, const double fi[6]) {
  int b = m->nbody++;
  m->body_parent[b] = parent;
  memcpy(m->body_pos[b], pos, 24);
  memcpy(m->body_mat[b], mat, 72);
  memcpy(m->body_ipos[b], ipos, 24);
  m->body_mass[b] = mass;
  /* MJCF fullinertia order Ixx Iyy Izz Ixy Ixz Iyz (mapping as DynamicModel.cpp:79-81) */
  double* I = m->body_inertia[b];
  I[0] = fi[0]; I[4] = fi[1]; I[8] = fi[2];
  I[1] = I[3] = fi[3]; I[2] = I[6] = fi[4]; I[5] = I[7] = fi[5];
  m->body_jntadr[b] = m->nv;
  m->body_jntnum[b] = 0;
  return b;
}
int orc_add_joint(orc_model* m, int body, int type, const double axis[3], const double pos[3],
                  double ref, int limited, const double range[2], double damping,
                  double armature, const double solref[2], const double solimp[5]) {
  if (type == ORC_JNT_FREE) {
    /* mjJNT_FREE [EXT]: dofs = world-axis translations then body-axis rotations, qpos = pos(3) + quat(4) */
    int first = m->nv;
    for (int k = 0; k < 6; k++) {
      int j = m->nv++;
      memset(m->jnt_axis[j], 0, 24);
      m->jnt_axis[j][k % 3] = 1.0;
      memset(m->jnt_pos[j], 0, 24);
      m->jnt_body[j] = body; m->jnt_type[j] = k < 3 ? ORC_JNT_FREE_T : ORC_JNT_FREE_R;
      m->jnt_qadr[j] = m->nq + (k < 3 ? k : 3);   /* rotations: address of the quaternion */
      m->jnt_ref[j] = 0; m->jnt_limited[j] = 0;
      m->jnt_damping[j] = damping; m->jnt_armature[j] = armature;
      memcpy(m->jnt_solref[j], solref, 16);
      memcpy(m->jnt_solimp[j], solimp, 40);
      if (m->body_jntnum[body] == 0) m->body_jntadr[body] = j;
      m->body_jntnum[body]++;
    }
    m->nq += 7;
    return first;
  }
  int j = m->nv++;
  m->jnt_qadr[j] = m->nq++;
  m->jnt_body[j] = body; m->jnt_type[j] = type;
  memcpy(m->jnt_axis[j], axis, 24);
  normalize3(m->jnt_axis[j]);
  memcpy(m->jnt_pos[j], pos, 24);
  m->jnt_ref[j] = ref; m->jnt_limited[j] = limited;
  m->jnt_range[j][0] = range[0]; m->jnt_range[j][1] = range[1];
  m->jnt_damping[j] = damping; m->jnt_armature[j] = armature;
  memcpy(m->jnt_solref[j], solref, 16);
  memcpy(m->jnt_solimp[j], solimp, 40);
  if (m->body_jntnum[body] == 0) m->body_jntadr[body] = j;
  m->body_jntnum[body]++;
  return j;
}
int orc_add_geom(orc_model* m, int body, int type, const double pos[3], const double mat[9],
                 const double size[2], int contype, int conaffinity, int condim,
                 const double friction[3], const double solref[2], const double solimp[5],
                 double margin, double gap) {
  int g = m->ngeom++;
  m->geom_body[g] = body; m->geom_type[g] = type;
  memcpy(m->geom_pos[g], pos, 24); memcpy(m->geom_mat[g], mat, 72);
  m->geom_size[g][0] = size[0]; m->geom_size[g][1] = size[1];
  m->geom_contype[g] = contype; m->geom_conaffinity[g] = conaffinity; m->geom_condim[g] = condim;
  memcpy(m->geom_friction[g], friction, 24);
  memcpy(m->geom_solref[g], solref, 16); memcpy(m->geom_solimp[g], solimp, 40);
  m->geom_margin[g] = margin; m->geom_gap[g] = gap;
  return g;
}
int orc_add_site(orc_model* m, int body, const double pos[3]) {
  int s = m->nsite++;
  m->site_body[s] = body; memcpy(m->site_pos[s], pos, 24);
  return s;
}
int orc_add_connect(orc_model* m, int body1, int body2, const double anchor[3],
                    const double solref[2], const double solimp[5]) {
  int e = m->neq++;
  m->eq_body1[e] = body1; m->eq_body2[e] = body2;
  memcpy(m->eq_anchor1[e], anchor, 24);
  memcpy(m->eq_solref[e], solref, 16); memcpy(m->eq_solimp[e], solimp, 40);
  return e;
}
int orc_add_motor(orc_model* m, int joint, double gear, int ctrllimited, const double range[2]) {
  int a = m->nu++;
  m->act_jnt[a] = joint; m->act_gear[a] = gear; m->act_limited[a] = ctrllimited;
  m->act_range[a][0] = range[0]; m->act_range[a][1] = range[1];
  return a;
}
void orc_set_option(orc_model* m, double timestep, int iterations, double tolerance,
                    double impratio, const double gravity[3]) {
  m->timestep = timestep; m->iterations = iterations; m->tolerance = tolerance;
  m->impratio = impratio; memcpy(m->gravity, gravity, 24);
}
int orc_nv(const orc_model* m) { return m->nv; }
int orc_nq(const orc_model* m) { return m->nq; }
int orc_nbody(const orc_model* m) { return m->nbody; }
double orc_total_mass(const orc_model* m) {
  double s = 0;
  for (int b = 1; b < m->nbody; b++) s += m->body_mass[b];
  return s;
}
void orc_get_consts(const orc_model* m, double* biw, double* diw, double* mi, double* a2) {
  for (int b = 0; b < m->nbody; b++) { biw[2 * b] = m->body_invweight0[b][0]; biw[2 * b + 1] = m->body_invweight0[b][1]; }
  for (int i = 0; i < m->nv; i++) diw[i] = m->dof_invweight0[i];
  *mi = m->meaninertia;
  for (int e = 0; e < m->neq; e++) memcpy(a2 + 3 * e, m->eq_anchor2[e], 24);
}

/* ------------------------------------------------------------------ kinematics */
orc_kin* orc_kin_new(void) { return (orc_kin*)calloc(1, sizeof(orc_kin)); }
void orc_kin_free(orc_kin* k) { free(k); }

/* Restates mj_kinematics + mj_comVel [EXT] / RBDL UpdateKinematics(q, qd, qdd=0)
 * (DynamicModel.cpp:237-242).  Everything is expressed in the world frame about the world
 * origin: S[j] = [w ; v_O] motion subspace, V[b] body spatial velocity, A[b] spatial
 * acceleration for qdd = 0 WITHOUT gravity. */
void orc_kin_update(const orc_model* m, orc_kin* k, const double* q, const double* qd) {
  memset(k->xpos[0], 0, 24);
  memset(k->xmat[0], 0, 72);
  k->xmat[0][0] = k->xmat[0][4] = k->xmat[0][8] = 1;
  memset(k->V[0], 0, 48);
  memset(k->A[0], 0, 48);
  for (int i = 0; i < m->nq; i++) k->q[i] = q[i];
  for (int i = 0; i < m->nv; i++) k->qd[i] = qd ? qd[i] : 0.0;
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b];
    double pos[3], mat[9], t[3];
    orc_mat3_mulv(t, k->xmat[p], m->body_pos[b]);
    for (int i = 0; i < 3; i++) pos[i] = k->xpos[p][i] + t[i];
    orc_mat3_mul(mat, k->xmat[p], m->body_mat[b]);
    double V[6], A[6];
    memcpy(V, k->V[p], 48);
    memcpy(A, k->A[p], 48);
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj;
      double ax[3], an[3];
      orc_mat3_mulv(ax, mat, m->jnt_axis[j]);
      orc_mat3_mulv(t, mat, m->jnt_pos[j]);
      for (int i = 0; i < 3; i++) an[i] = pos[i] + t[i];
      double* S = k->S[j];
      const int qa = m->jnt_qadr[j];
      if (m->jnt_type[j] == ORC_JNT_FREE_T) {
        /* mj_kinematics [EXT]: a free body's frame IS its qpos (the model's body pos is only qpos0) */
        int c = j - m->body_jntadr[b];
        S[0] = S[1] = S[2] = S[3] = S[4] = S[5] = 0;
        S[3 + c] = 1.0;
        pos[c] = q[qa];
        double Sd0[6];
        crm(Sd0, V, S);
        for (int i = 0; i < 6; i++) { A[i] += Sd0[i] * k->qd[j]; V[i] += S[i] * k->qd[j]; }
        continue;
      }
      if (m->jnt_type[j] == ORC_JNT_FREE_R) {
        /* the three rotational dofs at once (mj_comVel [EXT]: all three cdof_dot use the velocity BEFORE the
         * rotational part is added) */
        double w = q[qa], x = q[qa + 1], y = q[qa + 2], z = q[qa + 3];
        double nq = sqrt(w * w + x * x + y * y + z * z);
        w /= nq; x /= nq; y /= nq; z /= nq;   /* mj_normalizeQuat */
        mat[0] = 1 - 2 * (y * y + z * z); mat[1] = 2 * (x * y - w * z); mat[2] = 2 * (x * z + w * y);
        mat[3] = 2 * (x * y + w * z); mat[4] = 1 - 2 * (x * x + z * z); mat[5] = 2 * (y * z - w * x);
        mat[6] = 2 * (x * z - w * y); mat[7] = 2 * (y * z + w * x); mat[8] = 1 - 2 * (x * x + y * y);
        double V0[6];
        memcpy(V0, V, 48);
        for (int c = 0; c < 3; c++) {
          double* Sc = k->S[j + c];
          double a3[3] = {mat[c], mat[3 + c], mat[6 + c]};
          Sc[0] = a3[0]; Sc[1] = a3[1]; Sc[2] = a3[2];
          orc_cross(Sc + 3, pos, a3);
          double Sdc[6];
          crm(Sdc, V0, Sc);
          for (int i = 0; i < 6; i++) { A[i] += Sdc[i] * k->qd[j + c]; V[i] += Sc[i] * k->qd[j + c]; }
        }
        jj += 2;
        continue;
      }
      if (m->jnt_type[j] == ORC_JNT_HINGE) {
        S[0] = ax[0]; S[1] = ax[1]; S[2] = ax[2];
        orc_cross(S + 3, an, ax);
        double R[9];
        axis_angle(R, ax, q[qa] - m->jnt_ref[j]);
        double d[3] = {pos[0] - an[0], pos[1] - an[1], pos[2] - an[2]};
        orc_mat3_mulv(t, R, d);
        for (int i = 0; i < 3; i++) pos[i] = an[i] + t[i];
        orc_mat3_mul(mat, R, mat);
      } else {
        S[0] = S[1] = S[2] = 0;
        S[3] = ax[0]; S[4] = ax[1]; S[5] = ax[2];
        for (int i = 0; i < 3; i++) pos[i] += ax[i] * (q[qa] - m->jnt_ref[j]);
      }
      /* Sdot = V_before x S ; A += Sdot*qd ; V += S*qd */
      double Sd[6];
      crm(Sd, V, S);
      for (int i = 0; i < 6; i++) { A[i] += Sd[i] * k->qd[j]; V[i] += S[i] * k->qd[j]; }
    }
    memcpy(k->xpos[b], pos, 24);
    memcpy(k->xmat[b], mat, 72);
    memcpy(k->V[b], V, 48);
    memcpy(k->A[b], A, 48);
    orc_mat3_mulv(t, mat, m->body_ipos[b]);
    for (int i = 0; i < 3; i++) k->xipos[b][i] = pos[i] + t[i];
    /* world inertia about com: R I R^T */
    double RI[9], Rt[9];
    orc_mat3_mul(RI, mat, m->body_inertia[b]);
    for (int i = 0; i < 3; i++) for (int j2 = 0; j2 < 3; j2++) Rt[3 * i + j2] = mat[3 * j2 + i];
    orc_mat3_mul(k->Iw[b], RI, Rt);
  }
}

/* point Jacobian of world point P rigidly attached to `body` (mj_jac [EXT] / RBDL
 * CalcPointJacobian + CalcPointJacobian6D angular rows) */
void orc_jac_world(const orc_model* m, const orc_kin* k, int body, const double P[3],
                   double* jacp, double* jacr) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  for (int j = 0; j < nv; j++) {
    if (!m->anc[body][j]) continue;
    const double* S = k->S[j];
    double t[3];
    orc_cross(t, S, P);
    for (int i = 0; i < 3; i++) {
      if (jacp) jacp[i * nv + j] = S[3 + i] + t[i];
      if (jacr) jacr[i * nv + j] = S[i];
    }
  }
}

/* CRBA-equivalent mass matrix: M = sum_b m Jp^T Jp + Jr^T Iw Jr  (+ armature).
 * mj_crb [EXT]; RBDL CompositeRigidBodyAlgorithm + rotor inertia (DynamicModel.cpp:267-272) */
void orc_kin_mass_matrix(const orc_model* m, const orc_kin* k, double* M) {
  int nv = m->nv;
  memset(M, 0, sizeof(double) * nv * nv);
  double jp[3 * ORC_NV], jr[3 * ORC_NV];
  for (int b = 1; b < m->nbody; b++) {
    orc_jac_world(m, k, b, k->xipos[b], jp, jr);
    double mass = m->body_mass[b];
    const double* I = k->Iw[b];
    for (int i = 0; i < nv; i++) {
      if (!m->anc[b][i]) continue;
      double Ijr[3];
      double ri[3] = {jr[i], jr[nv + i], jr[2 * nv + i]};
      orc_mat3_mulv(Ijr, I, ri);
      for (int j = 0; j < nv; j++) {
        if (!m->anc[b][j]) continue;
        double s = mass * (jp[i] * jp[j] + jp[nv + i] * jp[nv + j] + jp[2 * nv + i] * jp[2 * nv + j]);
        s += Ijr[0] * jr[j] + Ijr[1] * jr[nv + j] + Ijr[2] * jr[2 * nv + j];
        M[i * nv + j] += s;
      }
    }
  }
  for (int i = 0; i < nv; i++) M[i * nv + i] += m->jnt_armature[i];
}

/* RNE with qdd = 0: bias = C(q,qd) qd + G(q).  mj_rne [EXT]; RBDL NonlinearEffects
 * (DynamicModel.cpp:320-323). */
void orc_kin_nonlinear_effects(const orc_model* m, const orc_kin* k, double* bias) {
  double F[ORC_NB][6];
  memset(F, 0, sizeof(F));
  for (int b = 1; b < m->nbody; b++) {
    double A[6], IA[6], IV[6], VIV[6];
    memcpy(A, k->A[b], 48);
    /* gravity as a fictitious base acceleration a_O = -g */
    A[3] -= m->gravity[0]; A[4] -= m->gravity[1]; A[5] -= m->gravity[2];
    inertia_apply(IA, m->body_mass[b], k->xipos[b], k->Iw[b], A);
    inertia_apply(IV, m->body_mass[b], k->xipos[b], k->Iw[b], k->V[b]);
    crf(VIV, k->V[b], IV);
    for (int i = 0; i < 6; i++) F[b][i] = IA[i] + VIV[i];
  }
  for (int b = m->nbody - 1; b >= 1; b--) {
    int p = m->body_parent[b];
    for (int i = 0; i < 6; i++) F[p][i] += F[b][i];
  }
  for (int j = 0; j < m->nv; j++) {
    const double* S = k->S[j];
    const double* f = F[m->jnt_body[j]];
    double s = 0;
    for (int i = 0; i < 6; i++) s += S[i] * f[i];
    bias[j] = s;
  }
}

void orc_kin_point_jacobian(const orc_model* m, const orc_kin* k, int body, const double pl[3],
                            double* jacp, double* jacr) {
  double P[3], t[3];
  orc_mat3_mulv(t, k->xmat[body], pl);
  for (int i = 0; i < 3; i++) P[i] = k->xpos[body][i] + t[i];
  orc_jac_world(m, k, body, P, jacp, jacr);
}

/* position, velocity and Jdot*qd of a body-fixed point (RBDL CalcBodyToBaseCoordinates,
 * CalcPointVelocity, CalcPointAcceleration with qdd=0; DynamicModel.cpp:341-344,360-367) */
void orc_kin_point_pos_vel_acc(const orc_model* m, const orc_kin* k, int body, const double pl[3],
                               double pos[3], double vel[3], double acc[3]) {
  (void)m;
  double P[3], t[3];
  orc_mat3_mulv(t, k->xmat[body], pl);
  for (int i = 0; i < 3; i++) P[i] = k->xpos[body][i] + t[i];
  const double* V = k->V[body];
  const double* A = k->A[body];
  double wP[3], aP[3], wv[3], vP[3];
  orc_cross(wP, V, P);
  for (int i = 0; i < 3; i++) vP[i] = V[3 + i] + wP[i];
  orc_cross(aP, A, P);
  orc_cross(wv, V, vP);
  for (int i = 0; i < 3; i++) {
    if (pos) pos[i] = P[i];
    if (vel) vel[i] = vP[i];
    if (acc) acc[i] = A[3 + i] + aP[i] + wv[i];
  }
}

/* ------------------------------------------------------------------ compile */
int orc_compile(orc_model* m) {
  int nv = m->nv;
  /* ancestor table */
  memset(m->anc, 0, sizeof(m->anc));
  for (int b = 1; b < m->nbody; b++) {
    int a = b;
    while (a != 0) {
      for (int jj = 0; jj < m->body_jntnum[a]; jj++) m->anc[b][m->body_jntadr[a] + jj] = 1;
      a = m->body_parent[a];
    }
  }
  for (int i = 0; i < nv; i++) {
    int qa = m->jnt_qadr[i];
    if (m->jnt_type[i] == ORC_JNT_FREE_T) m->qpos0[qa] = m->body_pos[m->jnt_body[i]][i - m->body_jntadr[m->jnt_body[i]]];
    else if (m->jnt_type[i] == ORC_JNT_FREE_R) { m->qpos0[qa] = 1.0; m->qpos0[qa + 1] = m->qpos0[qa + 2] = m->qpos0[qa + 3] = 0.0; }
    else m->qpos0[qa] = m->jnt_ref[i];
  }
  orc_kin* k = orc_kin_new();
  orc_kin_update(m, k, m->qpos0, NULL);
  /* connect: anchor2 = anchor1 expressed in body2 at qpos0 (mjCModel compile [EXT];
   * mirrored by DynamicModel.cpp:162-167) */
  for (int e = 0; e < m->neq; e++) {
    int b1 = m->eq_body1[e], b2 = m->eq_body2[e];
    double P[3], t[3];
    orc_mat3_mulv(t, k->xmat[b1], m->eq_anchor1[e]);
    for (int i = 0; i < 3; i++) P[i] = k->xpos[b1][i] + t[i] - k->xpos[b2][i];
    orc_mat3_tmulv(m->eq_anchor2[e], k->xmat[b2], P);
  }
  /* invweight0 / meaninertia (mj_setConst "set0" [EXT]) */
  double M[ORC_NV * ORC_NV], L[ORC_NV * ORC_NV];
  orc_kin_mass_matrix(m, k, M);
  if (orc_chol(nv, M, L)) { orc_kin_free(k); return -1; }
  double tr = 0;
  for (int i = 0; i < nv; i++) tr += M[i * nv + i];
  m->meaninertia = tr / nv;
  double e[ORC_NV], x[ORC_NV];
  for (int i = 0; i < nv; i++) {
    memset(e, 0, sizeof(e));
    e[i] = 1;
    orc_chol_solve(nv, L, e, x);
    m->dof_invweight0[i] = x[i];
  }
  double jp[3 * ORC_NV], jr[3 * ORC_NV];
  m->body_invweight0[0][0] = m->body_invweight0[0][1] = 0;
  for (int b = 1; b < m->nbody; b++) {
    orc_jac_world(m, k, b, k->xipos[b], jp, jr);
    double st = 0, sr = 0;
    for (int r = 0; r < 3; r++) {
      orc_chol_solve(nv, L, jp + r * nv, x);
      for (int i = 0; i < nv; i++) st += jp[r * nv + i] * x[i];
      orc_chol_solve(nv, L, jr + r * nv, x);
      for (int i = 0; i < nv; i++) sr += jr[r * nv + i] * x[i];
    }
    m->body_invweight0[b][0] = st / 3 > ORC_MINVAL ? st / 3 : ORC_MINVAL;
    m->body_invweight0[b][1] = sr / 3 > ORC_MINVAL ? sr / 3 : ORC_MINVAL;
  }
  orc_kin_free(k);
  return 0;
}

/* RBDL-style variant: DynamicModel.cpp:84-103 ("a bit of a hack") keeps the xyaxes rotation
 * only for bodies without joints or whose last joint has |ref| < 1e-3 (ref in DEGREES there);
 * otherwise identity, and the RBDL joint angle is the raw qpos (no ref subtraction).
 * Not modelled: RBDL receives the two xyaxes vectors normalised but NOT orthogonalised
 * (skew 4e-3 for the achilles rods); see DESIGN.md "known un-modelled reference quirks". */
orc_model* orc_model_rbdl_variant(const orc_model* src) {
  orc_model* m = (orc_model*)malloc(sizeof(orc_model));
  memcpy(m, src, sizeof(orc_model));
  for (int b = 1; b < m->nbody; b++) {
    int n = m->body_jntnum[b];
    if (n == 0) continue;
    int last = m->body_jntadr[b] + n - 1;
    double ref_deg = m->jnt_ref[last] * 180.0 / M_PI;
    if (m->jnt_type[last] == ORC_JNT_HINGE && fabs(ref_deg) >= 1e-3) {
      memset(m->body_mat[b], 0, 72);
      m->body_mat[b][0] = m->body_mat[b][4] = m->body_mat[b][8] = 1;
      for (int jj = 0; jj < n; jj++) m->jnt_ref[m->body_jntadr[b] + jj] = 0.0;
    }
  }
  /* connect anchor2 is computed at the ORIGINAL qpos0 = ref (DynamicModel.cpp:141-167), so
   * recompute it with the hacked frames evaluated at q = ref of the source model. */
  orc_kin* k = orc_kin_new();
  orc_kin_update(m, k, src->qpos0, NULL);
  for (int e = 0; e < m->neq; e++) {
    int b1 = m->eq_body1[e], b2 = m->eq_body2[e];
    double P[3], t[3];
    orc_mat3_mulv(t, k->xmat[b1], m->eq_anchor1[e]);
    for (int i = 0; i < 3; i++) P[i] = k->xpos[b1][i] + t[i] - k->xpos[b2][i];
    orc_mat3_tmulv(m->eq_anchor2[e], k->xmat[b2], P);
  }
  orc_kin_free(k);
  for (int i = 0; i < m->nv; i++) m->qpos0[m->jnt_qadr[i]] = m->jnt_ref[i];
  return m;
}

/* ------------------------------------------------------------------ data */
orc_data* orc_data_new(const orc_model* m) {
  orc_data* d = (orc_data*)calloc(1, sizeof(orc_data));
  d->nv = m->nv; d->nq = m->nq;
  for (int i = 0; i < m->nq; i++) d->qpos[i] = m->qpos0[i];
  d->min_capsule_gap = 1e30;
  return d;
}
void orc_data_free(orc_data* d) { free(d); }
void orc_set_state(orc_data* d, const double* qpos, const double* qvel) {
  for (int i = 0; i < d->nq; i++) d->qpos[i] = qpos[i];
  for (int i = 0; i < d->nv; i++) d->qvel[i] = qvel[i];
}
void orc_get_state(const orc_data* d, double* qpos, double* qvel) {
  for (int i = 0; i < d->nq; i++) qpos[i] = d->qpos[i];
  for (int i = 0; i < d->nv; i++) qvel[i] = d->qvel[i];
}
void orc_set_warmstart(orc_data* d, const double* w) { for (int i = 0; i < d->nv; i++) d->qacc_warmstart[i] = w[i]; }
void orc_get_warmstart(const orc_data* d, double* w) { for (int i = 0; i < d->nv; i++) w[i] = d->qacc_warmstart[i]; }
double orc_get_time(const orc_data* d) { return d->time; }
void orc_set_time(orc_data* d, double t) { d->time = t; }

/* ------------------------------------------------------------------ collision
 * mj_collision [EXT]: geom pairs filtered by (contype1 & conaffinity2) || (contype2 &
 * conaffinity1); contact kept when dist < margin (0 here).  Pair order = body-pair order =
 * geom order for this model (all plane pairs first).  Leg-leg capsule pairs are distance
 * checked only (they are 0.26 m apart in y and can never touch; min gap is recorded). */
static void make_frame(double f[9]) {
  /* mju_makeFrame [EXT]: f[0..2] normal given; f[3..5] optional y-axis hint */
  double* x = f; double* y = f + 3; double* z = f + 6;
  double n = sqrt(orc_dot3(x, x));
  for (int i = 0; i < 3; i++) x[i] /= n;
  if (sqrt(orc_dot3(y, y)) < 0.5) {
    y[0] = y[1] = y[2] = 0;
    if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  }
  double t = orc_dot3(x, y);
  for (int i = 0; i < 3; i++) y[i] -= t * x[i];
  n = sqrt(orc_dot3(y, y));
  for (int i = 0; i < 3; i++) y[i] /= n;
  orc_cross(z, x, y);
}

static int plane_sphere(const double ppos[3], const double pn[3], const double c[3], double r,
                        double margin, orc_contact* con) {
  double d[3] = {c[0] - ppos[0], c[1] - ppos[1], c[2] - ppos[2]};
  double dist = orc_dot3(d, pn) - r;
  if (dist > margin) return 0;
  con->dist = dist;
  for (int i = 0; i < 3; i++) {
    con->pos[i] = c[i] - pn[i] * (r + 0.5 * dist);
    con->frame[i] = pn[i];
    con->frame[3 + i] = 0;
  }
  make_frame(con->frame);
  return 1;
}

static double seg_seg_dist(const double a0[3], const double a1[3], const double b0[3], const double b1[3],
                           double pa[3], double pb[3]) {
  /* closest distance between two segments (Ericson, Real-Time Collision Detection 5.1.9) */
  double d1[3], d2[3], r[3];
  for (int i = 0; i < 3; i++) { d1[i] = a1[i] - a0[i]; d2[i] = b1[i] - b0[i]; r[i] = a0[i] - b0[i]; }
  double a = orc_dot3(d1, d1), e = orc_dot3(d2, d2), f = orc_dot3(d2, r);
  double s, t;
  double c = orc_dot3(d1, r), b = orc_dot3(d1, d2), den = a * e - b * b;
  s = den > 1e-14 ? (b * f - c * e) / den : 0.0;
  s = s < 0 ? 0 : (s > 1 ? 1 : s);
  t = (b * s + f) / e;
  if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
  else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
  double dd[3];
  for (int i = 0; i < 3; i++) { pa[i] = a0[i] + s * d1[i]; pb[i] = b0[i] + t * d2[i]; dd[i] = pa[i] - pb[i]; }
  return sqrt(orc_dot3(dd, dd));
}

static void collide(const orc_model* m, orc_data* d) {
  const orc_kin* k = &d->kin;
  d->ncon = 0;
  double gpos[ORC_NG][3], gmat[ORC_NG][9];
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    double t[3];
    orc_mat3_mulv(t, k->xmat[b], m->geom_pos[g]);
    for (int i = 0; i < 3; i++) gpos[g][i] = k->xpos[b][i] + t[i];
    orc_mat3_mul(gmat[g], k->xmat[b], m->geom_mat[g]);
  }
  for (int g1 = 0; g1 < m->ngeom; g1++)
    for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
      int b1 = m->geom_body[g1], b2 = m->geom_body[g2];
      if (b1 == b2) continue;
      double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
      double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
      orc_contact cons[2];
      int n = 0;
      if (m->geom_type[g1] == ORC_GEOM_PLANE) {
        double pn[3] = {gmat[g1][2], gmat[g1][5], gmat[g1][8]};
        if (m->geom_type[g2] == ORC_GEOM_SPHERE) {
          n = plane_sphere(gpos[g1], pn, gpos[g2], m->geom_size[g2][0], margin, &cons[0]);
          if (n) cons[0].slot = 2 * g2;
        } else if (m->geom_type[g2] == ORC_GEOM_CAPSULE) {
          /* mjc_PlaneCapsule [EXT]: end +axis first, then -axis; frame y aligned with axis */
          double ax[3] = {gmat[g2][2], gmat[g2][5], gmat[g2][8]};
          double hl = m->geom_size[g2][1];
          for (int e = 0; e < 2; e++) {
            double c[3];
            double sgn = e == 0 ? 1.0 : -1.0;
            for (int i = 0; i < 3; i++) c[i] = gpos[g2][i] + sgn * hl * ax[i];
            if (plane_sphere(gpos[g1], pn, c, m->geom_size[g2][0], margin, &cons[n])) {
              for (int i = 0; i < 3; i++) { cons[n].frame[i] = pn[i]; cons[n].frame[3 + i] = ax[i]; }
              make_frame(cons[n].frame);
              cons[n].slot = 2 * g2 + e;
              n++;
            }
          }
        }
      } else if (m->geom_type[g1] == ORC_GEOM_CAPSULE && m->geom_type[g2] == ORC_GEOM_CAPSULE) {
        double a0[3], a1[3], c0[3], c1[3];
        for (int i = 0; i < 3; i++) {
          a0[i] = gpos[g1][i] - m->geom_size[g1][1] * gmat[g1][3 * i + 2];
          a1[i] = gpos[g1][i] + m->geom_size[g1][1] * gmat[g1][3 * i + 2];
          c0[i] = gpos[g2][i] - m->geom_size[g2][1] * gmat[g2][3 * i + 2];
          c1[i] = gpos[g2][i] + m->geom_size[g2][1] * gmat[g2][3 * i + 2];
        }
        double pa[3], pb[3];
        double cd = seg_seg_dist(a0, a1, c0, c1, pa, pb);
        double gapd = cd - m->geom_size[g1][0] - m->geom_size[g2][0];
        if (gapd < d->min_capsule_gap) d->min_capsule_gap = gapd;
        if (gapd > margin || cd < 1e-12) continue;
        /* mjc_CapsuleCapsule [EXT], non-parallel case: one contact at the closest points of the two axes; normal
         * from geom1 to geom2, position midway between the two surfaces.  (Exactly parallel axes, where MuJoCo
         * emits two contacts, are not modelled.) */
        double nrm[3];
        for (int i = 0; i < 3; i++) nrm[i] = (pb[i] - pa[i]) / cd;
        cons[0].dist = gapd;
        for (int i = 0; i < 3; i++) {
          cons[0].pos[i] = pa[i] + nrm[i] * (m->geom_size[g1][0] + 0.5 * gapd);
          cons[0].frame[i] = nrm[i]; cons[0].frame[3 + i] = 0;
        }
        make_frame(cons[0].frame);
        cons[0].slot = 32 + (d->ncon & 31);
        n = 1;
      }
      for (int c = 0; c < n && d->ncon < ORC_MAXCON; c++) {
        orc_contact* con = &d->contact[d->ncon++];
        *con = cons[c];
        con->geom1 = g1; con->geom2 = g2;
        /* parameter mixing (mj_contactParam [EXT]): condim = max, friction = max,
         * solref/solimp: equal here, take geom with higher priority = same -> average = same */
        con->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
        double f0 = fmax(m->geom_friction[g1][0], m->geom_friction[g2][0]);
        double f1 = fmax(m->geom_friction[g1][1], m->geom_friction[g2][1]);
        double f2 = fmax(m->geom_friction[g1][2], m->geom_friction[g2][2]);
        con->friction[0] = con->friction[1] = f0; con->friction[2] = f1;
        con->friction[3] = con->friction[4] = f2;
        for (int i = 0; i < 2; i++) con->solref[i] = 0.5 * (m->geom_solref[g1][i] + m->geom_solref[g2][i]);
        for (int i = 0; i < 5; i++) con->solimp[i] = 0.5 * (m->geom_solimp[g1][i] + m->geom_solimp[g2][i]);
        con->includemargin = margin - gap;
      }
    }
}

/* ------------------------------------------------------------------ constraints
 * mj_makeConstraint [EXT]: equality rows, then joint-limit rows, then contact rows. */
static void make_constraints(const orc_model* m, orc_data* d) {
  const orc_kin* k = &d->kin;
  int nv = m->nv, n = 0;
  double jp1[3 * ORC_NV], jp2[3 * ORC_NV];
  /* connects */
  for (int e = 0; e < m->neq; e++) {
    int b1 = m->eq_body1[e], b2 = m->eq_body2[e];
    double P1[3], P2[3], t[3];
    orc_mat3_mulv(t, k->xmat[b1], m->eq_anchor1[e]);
    for (int i = 0; i < 3; i++) P1[i] = k->xpos[b1][i] + t[i];
    orc_mat3_mulv(t, k->xmat[b2], m->eq_anchor2[e]);
    for (int i = 0; i < 3; i++) P2[i] = k->xpos[b2][i] + t[i];
    orc_jac_world(m, k, b1, P1, jp1, NULL);
    orc_jac_world(m, k, b2, P2, jp2, NULL);
    double tr = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    for (int r = 0; r < 3; r++) {
      for (int j = 0; j < nv; j++) d->efc_J[n * nv + j] = jp1[r * nv + j] - jp2[r * nv + j];
      d->efc_pos[n] = P1[r] - P2[r];
      d->efc_margin[n] = 0;
      d->efc_type[n] = ORC_EFC_EQ; d->efc_id[n] = e; d->efc_dim[n] = r == 0 ? 3 : 0;
      d->efc_diagApprox[n] = tr;
      n++;
    }
  }
  d->ne = n;
  /* joint limits */
  for (int j = 0; j < nv; j++) {
    if (!m->jnt_limited[j]) continue;
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][side == -1 ? 0 : 1] - d->qpos[m->jnt_qadr[j]]);
      if (dist < 0 /* margin */) {
        memset(d->efc_J + n * nv, 0, sizeof(double) * nv);
        d->efc_J[n * nv + j] = -(double)side;
        d->efc_pos[n] = dist; d->efc_margin[n] = 0;
        d->efc_type[n] = ORC_EFC_LIMIT; d->efc_id[n] = j; d->efc_dim[n] = 1;
        d->efc_diagApprox[n] = m->dof_invweight0[j];
        n++;
      }
    }
  }
  /* contacts (elliptic, condim 3): rows normal, tangent1, tangent2 */
  for (int c = 0; c < d->ncon; c++) {
    const orc_contact* con = &d->contact[c];
    int b1 = m->geom_body[con->geom1], b2 = m->geom_body[con->geom2];
    orc_jac_world(m, k, b1, con->pos, jp1, NULL);
    orc_jac_world(m, k, b2, con->pos, jp2, NULL);
    double tr = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    for (int r = 0; r < con->dim; r++) {
      const double* ax = con->frame + 3 * r;
      for (int j = 0; j < nv; j++) {
        double s = 0;
        for (int i = 0; i < 3; i++) s += ax[i] * (jp2[i * nv + j] - jp1[i * nv + j]);
        d->efc_J[n * nv + j] = s;
      }
      d->efc_pos[n] = r == 0 ? con->dist : 0.0;
      d->efc_margin[n] = r == 0 ? con->includemargin : 0.0;
      d->efc_type[n] = ORC_EFC_CONTACT; d->efc_id[n] = c; d->efc_dim[n] = r == 0 ? con->dim : 0;
      d->efc_diagApprox[n] = tr;
      n++;
    }
  }
  d->nefc = n;
}

/* getimpedance [EXT] (5-parameter solimp; 3-parameter inputs get midpoint .5, power 2) */
static double impedance(const double* si, double pos, double margin) {
  if (si[0] == si[1] || si[2] <= ORC_MINVAL) return 0.5 * (si[0] + si[1]);
  double x = (pos - margin) / si[2];
  if (x < 0) x = -x;
  if (x >= 1) return si[1];
  if (x <= 0) return si[0];
  double y;
  if (si[4] == 1) y = x;
  else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
  else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
  return si[0] + y * (si[1] - si[0]);
}

/* mj_makeImpedance [EXT]: R, D, KBIP, aref */
static void make_impedance(const orc_model* m, orc_data* d) {
  int nv = m->nv;
  for (int i = 0; i < d->nefc; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += d->efc_J[i * nv + j] * d->qvel[j];
    d->efc_vel[i] = s;
  }
  for (int i = 0; i < d->nefc;) {
    int dim = d->efc_dim[i];
    const double *solref, *solimp;
    double pos;
    int id = d->efc_id[i];
    if (d->efc_type[i] == ORC_EFC_EQ) {
      solref = m->eq_solref[id]; solimp = m->eq_solimp[id];
      pos = sqrt(d->efc_pos[i] * d->efc_pos[i] + d->efc_pos[i + 1] * d->efc_pos[i + 1] + d->efc_pos[i + 2] * d->efc_pos[i + 2]);
    } else if (d->efc_type[i] == ORC_EFC_LIMIT) {
      solref = m->jnt_solref[id]; solimp = m->jnt_solimp[id]; pos = d->efc_pos[i];
    } else {
      solref = d->contact[id].solref; solimp = d->contact[id].solimp; pos = d->efc_pos[i];
    }
    double sr0 = solref[0], sr1 = solref[1];
    if (sr0 > 0 && sr0 < 2 * m->timestep) sr0 = 2 * m->timestep; /* refsafe */
    double imp = impedance(solimp, pos, d->efc_margin[i]);
    for (int j = 0; j < dim; j++) {
      double R = (1 - imp) * d->efc_diagApprox[i + j] / imp;
      d->efc_R[i + j] = R > ORC_MINVAL ? R : ORC_MINVAL;
      int fric = d->efc_type[i] == ORC_EFC_CONTACT && j > 0;
      double K = fric ? 0.0 : 1.0 / fmax(ORC_MINVAL, solimp[1] * solimp[1] * sr0 * sr0 * sr1 * sr1);
      double B = 2.0 / fmax(ORC_MINVAL, solimp[1] * sr0);
      d->efc_KBIP[i + j][0] = K; d->efc_KBIP[i + j][1] = B; d->efc_KBIP[i + j][2] = imp; d->efc_KBIP[i + j][3] = 0;
    }
    if (d->efc_type[i] == ORC_EFC_CONTACT && dim > 1) {
      /* friction rows regularised by impratio (elliptic) */
      const double* fr = d->contact[id].friction;
      d->efc_R[i + 1] = d->efc_R[i] / fmax(ORC_MINVAL, m->impratio);
      for (int j = 1; j < dim - 1; j++) d->efc_R[i + j + 1] = d->efc_R[i + 1] * fr[0] * fr[0] / (fr[j] * fr[j]);
    }
    i += dim;
  }
  for (int i = 0; i < d->nefc; i++) {
    d->efc_D[i] = 1.0 / d->efc_R[i];
    d->efc_aref[i] = -d->efc_KBIP[i][1] * d->efc_vel[i] - d->efc_KBIP[i][0] * d->efc_KBIP[i][2] * (d->efc_pos[i] - d->efc_margin[i]);
  }
}

/* mju_QCQP2 [EXT] */
static int qcqp2(double* res, const double* Ain, const double* bin, const double* dd, double r) {
  double b1 = bin[0] * dd[0], b2 = bin[1] * dd[1];
  double A11 = Ain[0] * dd[0] * dd[0], A22 = Ain[3] * dd[1] * dd[1], A12 = Ain[1] * dd[0] * dd[1];
  double la = 0, v1 = 0, v2 = 0;
  for (int it = 0; it < 20; it++) {
    double det = (A11 + la) * (A22 + la) - A12 * A12;
    if (det < 1e-10) { res[0] = 0; res[1] = 0; return 0; }
    double di = 1 / det;
    double P11 = (A22 + la) * di, P22 = (A11 + la) * di, P12 = -A12 * di;
    v1 = -P11 * b1 - P12 * b2; v2 = -P12 * b1 - P22 * b2;
    double val = v1 * v1 + v2 * v2 - r * r;
    if (val < 1e-10) break;
    double deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
    double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  res[0] = v1 * dd[0]; res[1] = v2 * dd[1];
  return la != 0;
}

/* mj_constraintUpdate [EXT] restricted to force computation (used by the PGS warm start) */
static void constraint_update(const orc_model* m, const orc_data* d, const double* jar, double* force) {
  for (int i = 0; i < d->nefc;) {
    int dim = d->efc_dim[i];
    if (d->efc_type[i] == ORC_EFC_EQ) {
      for (int j = 0; j < dim; j++) force[i + j] = -d->efc_D[i + j] * jar[i + j];
    } else if (d->efc_type[i] == ORC_EFC_LIMIT || dim == 1) {
      force[i] = jar[i] < 0 ? -d->efc_D[i] * jar[i] : 0.0;
    } else {
      const double* fr = d->contact[d->efc_id[i]].friction;
      double mu = fr[0] / sqrt(fmax(ORC_MINVAL, m->impratio));
      double U[6];
      U[0] = jar[i] * mu;
      for (int j = 1; j < dim; j++) U[j] = jar[i + j] * fr[j - 1];
      double N = U[0], T = 0;
      for (int j = 1; j < dim; j++) T += U[j] * U[j];
      T = sqrt(T);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
        for (int j = 0; j < dim; j++) force[i + j] = 0;
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) force[i + j] = -d->efc_D[i + j] * jar[i + j];
      } else {
        double Dm = d->efc_D[i] / fmax(ORC_MINVAL, mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        force[i] = -Dm * NmT * mu;
        for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * U[j] * fr[j - 1];
      }
    }
    i += dim;
  }
}

/* mj_solPGS [EXT] */
static void solve_pgs(const orc_model* m, orc_data* d) {
  int nv = m->nv, nefc = d->nefc;
  double* AR = d->efc_AR;
  double* f = d->efc_force;
  const double* b = d->efc_b;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  int iter = 0;
  while (iter < m->iterations) {
    double improvement = 0;
    for (int i = 0; i < nefc;) {
      int dim = d->efc_type[i] == ORC_EFC_CONTACT ? d->efc_dim[i] : 1;
      double res[6], old[6], Athis[36];
      for (int j = 0; j < dim; j++) {
        double s = b[i + j];
        for (int c = 0; c < nefc; c++) s += AR[(i + j) * nefc + c] * f[c];
        res[j] = s;
        old[j] = f[i + j];
      }
      for (int j = 0; j < dim; j++) for (int c = 0; c < dim; c++) Athis[j * dim + c] = AR[(i + j) * nefc + i + c];
      if (dim == 1) {
        f[i] -= res[0] / AR[i * nefc + i];
        if (d->efc_type[i] != ORC_EFC_EQ && f[i] < 0) f[i] = 0;
      } else {
        const double* mu = d->contact[d->efc_id[i]].friction;
        if (f[i] < ORC_MINVAL) {
          f[i] -= res[0] / AR[i * nefc + i];
          if (f[i] < 0) f[i] = 0;
          for (int j = 1; j < dim; j++) f[i + j] = 0;
        } else {
          double v[6], v1[6], denom = 0;
          for (int j = 0; j < dim; j++) v[j] = f[i + j];
          for (int j = 0; j < dim; j++) { double s = 0; for (int c = 0; c < dim; c++) s += Athis[j * dim + c] * v[c]; v1[j] = s; }
          for (int j = 0; j < dim; j++) denom += v[j] * v1[j];
          if (denom >= ORC_MINVAL) {
            double x = 0;
            for (int j = 0; j < dim; j++) x -= v[j] * res[j];
            x /= denom;
            if (f[i] + x * v[0] < 0) x = -1.0;
            for (int j = 0; j < dim; j++) f[i + j] += x * v[j];
          }
        }
        /* friction update with the normal fixed */
        double Ac[25], bc[5], v[5];
        int dm = dim - 1;
        for (int j = 0; j < dm; j++) {
          bc[j] = res[1 + j];
          for (int c = 0; c < dm; c++) { Ac[j * dm + c] = Athis[(j + 1) * dim + c + 1]; bc[j] -= Ac[j * dm + c] * old[1 + c]; }
          bc[j] += Athis[(j + 1) * dim] * (f[i] - old[0]);
        }
        if (f[i] < ORC_MINVAL) {
          for (int j = 1; j < dim; j++) f[i + j] = 0;
        } else {
          int active = 0;
          if (dim == 3) active = qcqp2(v, Ac, bc, mu, f[i]);
          else { /* dim 2 (not produced by MuJoCo; kept for completeness) */
            v[0] = -bc[0] / Ac[0];
            double lim = mu[0] * f[i];
            if (v[0] > lim) { v[0] = lim; active = 1; } else if (v[0] < -lim) { v[0] = -lim; active = 1; }
          }
          if (active) {
            double s = 0;
            for (int j = 0; j < dm; j++) s += v[j] * v[j] / (mu[j] * mu[j]);
            s = sqrt(f[i] * f[i] / fmax(ORC_MINVAL, s));
            for (int j = 0; j < dm; j++) v[j] *= s;
          }
          for (int j = 0; j < dm; j++) f[i + 1 + j] = v[j];
        }
      }
      /* costChange [EXT] */
      double delta[6], change = 0;
      for (int j = 0; j < dim; j++) delta[j] = f[i + j] - old[j];
      for (int j = 0; j < dim; j++) {
        double s = 0;
        for (int c = 0; c < dim; c++) s += Athis[j * dim + c] * delta[c];
        change += 0.5 * delta[j] * s + delta[j] * res[j];
      }
      if (change > 1e-10) {
        for (int j = 0; j < dim; j++) f[i + j] = old[j];
        change = 0;
      }
      improvement -= change;
      i += dim;
    }
    improvement *= scale;
    iter++;
    if (improvement < m->tolerance) break;
  }
  d->solver_iter = iter;
}

/* ------------------------------------------------------------------ forward / step
 * mj_forward [EXT]: fwdPosition, fwdVelocity, fwdActuation, fwdAcceleration, fwdConstraint */
void orc_forward(const orc_model* m, orc_data* d, const double* ctrl) {
  int nv = m->nv;
  orc_kin* k = &d->kin;
  orc_kin_update(m, k, d->qpos, d->qvel);
  orc_kin_mass_matrix(m, k, d->M);
  orc_chol(nv, d->M, d->L);
  collide(m, d);
  make_constraints(m, d);
  /* velocity-dependent */
  orc_kin_nonlinear_effects(m, k, d->qfrc_bias);
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->jnt_damping[i] * d->qvel[i];
  /* actuation: ctrl clamped to ctrlrange, force = gear*ctrl (mj_fwdActuation [EXT]) */
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) {
    double c = ctrl ? ctrl[a] : 0.0;
    if (m->act_limited[a]) { if (c < m->act_range[a][0]) c = m->act_range[a][0]; if (c > m->act_range[a][1]) c = m->act_range[a][1]; }
    d->ctrl[a] = c;
    d->qfrc_actuator[m->act_jnt[a]] += m->act_gear[a] * c;
  }
  double rhs[ORC_NV];
  for (int i = 0; i < nv; i++) rhs[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  orc_chol_solve(nv, d->L, rhs, d->qacc_smooth);
  /* constraint */
  int nefc = d->nefc;
  memset(d->qfrc_constraint, 0, sizeof(double) * nv);
  if (nefc == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv); d->solver_iter = 0; return; }
  make_impedance(m, d);
  /* b = J qacc_smooth - aref ; AR = J M^-1 J^T + R */
  static __thread double MinvJT[ORC_MAXEFC * ORC_NV];
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += d->efc_J[i * nv + j] * d->qacc_smooth[j];
    d->efc_b[i] = s - d->efc_aref[i];
    orc_chol_solve(nv, d->L, d->efc_J + i * nv, MinvJT + i * nv);
  }
  for (int i = 0; i < nefc; i++)
    for (int c = 0; c < nefc; c++) {
      double s = 0;
      for (int j = 0; j < nv; j++) s += d->efc_J[i * nv + j] * MinvJT[c * nv + j];
      d->efc_AR[i * nefc + c] = s + (i == c ? d->efc_R[i] : 0.0);
    }
  /* warm start (mj_fwdConstraint/warmstart [EXT]) */
  double jar[ORC_MAXEFC];
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += d->efc_J[i * nv + j] * d->qacc_warmstart[j];
    jar[i] = s - d->efc_aref[i];
  }
  constraint_update(m, d, jar, d->efc_force);
  double cost = 0;
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int c = 0; c < nefc; c++) s += d->efc_AR[i * nefc + c] * d->efc_force[c];
    cost += d->efc_force[i] * (0.5 * s + d->efc_b[i]);
  }
  if (cost > 0) memset(d->efc_force, 0, sizeof(double) * nefc);
  solve_pgs(m, d);
  /* dualFinish: qacc = qacc_smooth + M^-1 J^T f */
  for (int j = 0; j < nv; j++) {
    double s = 0;
    for (int i = 0; i < nefc; i++) s += d->efc_J[i * nv + j] * d->efc_force[i];
    d->qfrc_constraint[j] = s;
  }
  double dq[ORC_NV];
  orc_chol_solve(nv, d->L, d->qfrc_constraint, dq);
  for (int j = 0; j < nv; j++) d->qacc[j] = d->qacc_smooth[j] + dq[j];
}

/* mj_step = mj_forward + mj_Euler [EXT]: implicit-in-damping velocity update
 * (M + h D) qacc' = qfrc_smooth + qfrc_constraint ; qvel += h qacc' ; qpos += h qvel */
void orc_step(const orc_model* m, orc_data* d, const double* ctrl) {
  int nv = m->nv;
  double h = m->timestep;
  orc_forward(m, d, ctrl);
  double MM[ORC_NV * ORC_NV], LL[ORC_NV * ORC_NV], rhs[ORC_NV], qacc[ORC_NV];
  int damped = 0;
  for (int i = 0; i < nv; i++) if (m->jnt_damping[i] > 0) damped = 1;
  if (damped) {
    memcpy(MM, d->M, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) {
      MM[i * nv + i] += h * m->jnt_damping[i];
      rhs[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i] + d->qfrc_constraint[i];
    }
    orc_chol(nv, MM, LL);
    orc_chol_solve(nv, LL, rhs, qacc);
  } else
    memcpy(qacc, d->qacc, sizeof(double) * nv);
  for (int i = 0; i < nv; i++) {
    d->qvel[i] += h * qacc[i];
    d->qacc_warmstart[i] = d->qacc[i];
  }
  /* mj_integratePos [EXT]: scalar joints q += h v; a free joint's quaternion is rotated by h * (body-frame angular
   * velocity) (mju_quatIntegrate: q <- q * [cos(a/2), sin(a/2) w/|w|], a = h |w|, then normalised) */
  for (int i = 0; i < nv; i++) {
    int qa = m->jnt_qadr[i];
    if (m->jnt_type[i] == ORC_JNT_FREE_R) {
      double w[3] = {d->qvel[i], d->qvel[i + 1], d->qvel[i + 2]};
      double nw = sqrt(orc_dot3(w, w)), ang = h * nw;
      if (nw > ORC_MINVAL) {
        double sn = sin(0.5 * ang) / nw, cs = cos(0.5 * ang);
        double r[4] = {cs, sn * w[0], sn * w[1], sn * w[2]};
        double* q = d->qpos + qa;
        double o[4] = {q[0] * r[0] - q[1] * r[1] - q[2] * r[2] - q[3] * r[3],
                       q[0] * r[1] + q[1] * r[0] + q[2] * r[3] - q[3] * r[2],
                       q[0] * r[2] - q[1] * r[3] + q[2] * r[0] + q[3] * r[1],
                       q[0] * r[3] + q[1] * r[2] - q[2] * r[1] + q[3] * r[0]};
        double no = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
        for (int c = 0; c < 4; c++) q[c] = o[c] / no;
      }
      i += 2;
    } else
      d->qpos[qa] += h * d->qvel[i];
  }
  d->time += h;
}

/* ------------------------------------------------------------------ inspection */
void orc_get_M(const orc_data* d, double* M) { memcpy(M, d->M, sizeof(double) * d->nv * d->nv); }
void orc_get_vectors(const orc_data* d, double* qb, double* qp, double* qa, double* qs, double* qacc) {
  for (int i = 0; i < d->nv; i++) {
    if (qb) qb[i] = d->qfrc_bias[i];
    if (qp) qp[i] = d->qfrc_passive[i];
    if (qa) qa[i] = d->qfrc_actuator[i];
    if (qs) qs[i] = d->qacc_smooth[i];
    if (qacc) qacc[i] = d->qacc[i];
  }
}
int orc_get_nefc(const orc_data* d) { return d->nefc; }
int orc_get_ncon(const orc_data* d) { return d->ncon; }
int orc_get_solver_iter(const orc_data* d) { return d->solver_iter; }
void orc_get_efc(const orc_data* d, double* J, double* pos, double* aref, double* R, double* force, int* type, int* id) {
  for (int i = 0; i < d->nefc; i++) {
    if (J) memcpy(J + i * d->nv, d->efc_J + i * d->nv, sizeof(double) * d->nv);
    if (pos) pos[i] = d->efc_pos[i];
    if (aref) aref[i] = d->efc_aref[i];
    if (R) R[i] = d->efc_R[i];
    if (force) force[i] = d->efc_force[i];
    if (type) type[i] = d->efc_type[i];
    if (id) id[i] = d->efc_id[i];
  }
}
void orc_get_contacts(const orc_data* d, double* dist, double* pos, double* frame, int* geom) {
  for (int c = 0; c < d->ncon; c++) {
    if (dist) dist[c] = d->contact[c].dist;
    if (pos) memcpy(pos + 3 * c, d->contact[c].pos, 24);
    if (frame) memcpy(frame + 9 * c, d->contact[c].frame, 72);
    if (geom) geom[c] = d->contact[c].geom2;
  }
}
unsigned long long orc_contact_mask(const orc_model* m, const orc_data* d) {
  (void)m;
  unsigned long long mask = 0;
  for (int c = 0; c < d->ncon; c++) mask |= 1ull << d->contact[c].slot;
  return mask;
}
void orc_body_pose(const orc_data* d, int body, double xpos[3], double xmat[9]) {
  memcpy(xpos, d->kin.xpos[body], 24);
  memcpy(xmat, d->kin.xmat[body], 72);
}
void orc_site_pos(const orc_model* m, const orc_data* d, int site, double p[3]) {
  int b = m->site_body[site];
  double t[3];
  orc_mat3_mulv(t, d->kin.xmat[b], m->site_pos[site]);
  for (int i = 0; i < 3; i++) p[i] = d->kin.xpos[b][i] + t[i];
}
/* total mechanical energy at the kinematic state stored by the last orc_forward */
double orc_energy(const orc_model* m, const orc_data* d, double* kinetic, double* potential) {
  double T = 0, U = 0;
  int nv = m->nv;
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) T += 0.5 * d->qvel[i] * d->M[i * nv + j] * d->qvel[j];
  for (int b = 1; b < m->nbody; b++) U -= m->body_mass[b] * orc_dot3(m->gravity, d->kin.xipos[b]);
  if (kinetic) *kinetic = T;
  if (potential) *potential = U;
  return T + U;
}

/* ------------------------------------------------------------------ 3-D torque rollouts (see cassie_oracle.h) */
long orc_rollout_tree(const orc_model* m, int n_envs, int n_steps, int hold, const double* actions,
                      const double* qpos0, const double* qvel0, double z_done, const double* reset_qpos,
                      const double* reset_qvel, double* out_qpos, double* out_qvel, int* out_resets, int n_threads) {
  int nq = m->nq, nv = m->nv, nu = m->nu;
  int nact = hold > 0 ? (n_steps + hold - 1) / hold : 1;
  long total = 0;
  (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total) num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
  for (int e = 0; e < n_envs; e++) {
    orc_data* d = orc_data_new(m);
    orc_set_state(d, qpos0 + (size_t)e * nq, qvel0 + (size_t)e * nv);
    int resets = 0;
    for (int s = 0; s < n_steps; s++) {
      const double* u = actions ? actions + ((size_t)e * nact + (hold > 0 ? s / hold : 0)) * nu : NULL;
      orc_step(m, d, u);
      total++;
      if (z_done > 0 && hold > 0 && (s + 1) % hold == 0 && (d->qpos[2] < z_done || !(d->qpos[2] == d->qpos[2]))) {
        orc_set_state(d, reset_qpos, reset_qvel);
        memset(d->qacc_warmstart, 0, sizeof(d->qacc_warmstart));
        resets++;
      }
    }
    orc_get_state(d, out_qpos + (size_t)e * nq, out_qvel + (size_t)e * nv);
    if (out_resets) out_resets[e] = resets;
    orc_data_free(d);
  }
  return total;
}
