/* TEST INFRASTRUCTURE -- fp64 CPU oracle, part 2 (see cassie_oracle.h).  PARITY UNPINNED.
 * Restates the controller / facade side of the reference, which IS fully specified in-repo:
 *   CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:29-237, src/DynamicState.cpp:45-91,
 *   src/DynamicModel.cpp:237-367, src/OSC_RBDL.cpp:29-291, src/HelperFunctions.h:8-29.
 * Third-party pieces restated: Eigen JacobiSVD pseudo-inverse (one-sided Jacobi here) and
 * the qpOASES solve (exact optimum of the same QP by a dense primal active-set method).
 */
#include "cassie_oracle_internal.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define NQ 13
#define NUU 6
#define NCON 4
#define NEQR 6
#define NTASK 16
#define NX 39

/* ------------------------------------------------------------------ pinv via one-sided Jacobi SVD
 * HelperFunctions.h:8-29: pinv = V diag(1/s_i if s_i > tol else 0) U^T */
void orc_pinv(int rows, int cols, const double* Ain, double tol, double* Ainv, double* svout) {
  /* work on B = A (rows x cols) if rows >= cols else A^T, so that B is tall (m x n, m >= n) */
  int transposed = rows < cols;
  int m = transposed ? cols : rows, n = transposed ? rows : cols;
  double* B = (double*)malloc(sizeof(double) * m * n);
  double* V = (double*)calloc(n * n, sizeof(double));
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) B[i * n + j] = transposed ? Ain[j * cols + i] : Ain[i * cols + j];
  for (int i = 0; i < n; i++) V[i * n + i] = 1;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < m; i++) { a += B[i * n + p] * B[i * n + p]; b += B[i * n + q] * B[i * n + q]; c += B[i * n + p] * B[i * n + q]; }
        if (fabs(c) <= 1e-300 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
        off = fmax(off, fabs(c) / sqrt(a * b));
        double zeta = (b - a) / (2 * c);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
        double cs = 1 / sqrt(1 + t * t), sn = cs * t;
        for (int i = 0; i < m; i++) {
          double bp = B[i * n + p], bq = B[i * n + q];
          B[i * n + p] = cs * bp - sn * bq; B[i * n + q] = sn * bp + cs * bq;
        }
        for (int i = 0; i < n; i++) {
          double vp = V[i * n + p], vq = V[i * n + q];
          V[i * n + p] = cs * vp - sn * vq; V[i * n + q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-15) break;
  }
  /* B = U S ; pinv(Btall) = V S^-1 U^T = sum_j v_j (b_j / s_j^2)^T */
  double* Pt = (double*)calloc(n * m, sizeof(double)); /* n x m */
  for (int j = 0; j < n; j++) {
    double s2 = 0;
    for (int i = 0; i < m; i++) s2 += B[i * n + j] * B[i * n + j];
    double s = sqrt(s2);
    if (svout) svout[j] = s;
    if (s > tol)
      for (int r = 0; r < n; r++)
        for (int i = 0; i < m; i++) Pt[r * m + i] += V[r * n + j] * B[i * n + j] / s2;
  }
  /* Ainv is cols x rows */
  for (int i = 0; i < cols; i++)
    for (int j = 0; j < rows; j++) Ainv[i * rows + j] = transposed ? Pt[j * m + i] : Pt[i * m + j];
  free(B); free(V); free(Pt);
}

/* ------------------------------------------------------------------ dense linear solve */
static int gauss_solve(int n, double* A, double* b) {
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
    if (fabs(A[piv * n + c]) < 1e-300) return -1;
    if (piv != c) {
      for (int k = 0; k < n; k++) { double t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < n; r++) {
      double f = A[r * n + c] / A[c * n + c];
      if (f == 0) continue;
      for (int k = c; k < n; k++) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = b[r];
    for (int k = r + 1; k < n; k++) s -= A[r * n + k] * b[k];
    b[r] = s / A[r * n + r];
  }
  return 0;
}

/* ------------------------------------------------------------------ strictly convex QP, primal active set
 * Stands in for qpOASES::SQProblem::init/hotstart (OSC_RBDL.cpp:275-278): returns the exact
 * optimum (the reference's wall-clock cut-off of the hot start is not reproducible).
 * Constraints are normalised to  a_k' x <= b_k.  Needs a feasible start in x. */
static int qp_solve_ws(int n, int mc, const double* G, const double* g, const double* A,
                       const double* lbA, const double* ubA, const double* lb, const double* ub, double* x,
                       int* ws, int* nws);
int orc_qp_solve(int n, int mc, const double* G, const double* g, const double* A,
                 const double* lbA, const double* ubA, const double* lb, const double* ub, double* x) {
  return qp_solve_ws(n, mc, G, g, A, lbA, ubA, lb, ub, x, NULL, NULL);
}
/* ws/nws (optional): working set carried from the previous solve of a QP with the same constraints --
 * the qpOASES hot start of the reference (OSC_RBDL.cpp:278); x must then be the previous solution. */
static int qp_solve_ws(int n, int mc, const double* G, const double* g, const double* A,
                       const double* lbA, const double* ubA, const double* lb, const double* ub, double* x,
                       int* ws, int* nws) {
  const double INF = 1e30;
  int maxk = 2 * n + 2 * mc;
  double* ca = (double*)calloc((size_t)maxk * n, sizeof(double));
  double* cb = (double*)calloc(maxk, sizeof(double));
  int nk = 0;
  for (int i = 0; i < n; i++) {
    if (ub && ub[i] < INF) { ca[nk * n + i] = 1; cb[nk++] = ub[i]; }
    if (lb && lb[i] > -INF) { ca[nk * n + i] = -1; cb[nk++] = -lb[i]; }
  }
  for (int r = 0; r < mc; r++) {
    if (ubA && ubA[r] < INF) { for (int i = 0; i < n; i++) ca[nk * n + i] = A[r * n + i]; cb[nk++] = ubA[r]; }
    if (lbA && lbA[r] > -INF) { for (int i = 0; i < n; i++) ca[nk * n + i] = -A[r * n + i]; cb[nk++] = -lbA[r]; }
  }
  /* anti-stalling: the friction pyramids have degenerate vertices (at the apex beta_z = 0 nine
   * constraints meet in five dimensions) where a textbook active-set method stalls for hundreds of
   * zero-length steps; relaxing each right-hand side by a distinct amount of order 1e-11 splits them */
  for (int k2 = 0; k2 < nk; k2++) cb[k2] += 1e-11 * (1.0 + fmod(0.6180339887 * (k2 + 1), 1.0));
  int* W = (int*)calloc(nk + 1, sizeof(int));
  int nw = 0, iter, rc = -1;
  if (ws && nws) { /* keep only constraints that are still active at x */
    for (int w = 0; w < *nws && w < nk; w++) {
      int k2 = ws[w];
      double ax = 0;
      for (int j = 0; j < n; j++) ax += ca[k2 * n + j] * x[j];
      if (fabs(ax - cb[k2]) <= 1e-9 * (1.0 + fabs(cb[k2]))) W[nw++] = k2;
    }
  }
  int dim = n + nk;
  double* K = (double*)malloc(sizeof(double) * dim * dim);
  double* rhs = (double*)malloc(sizeof(double) * dim);
  for (iter = 0; iter < 2000; iter++) {
    int N = n + nw;
    for (int i = 0; i < N * N; i++) K[i] = 0;
    for (int i = 0; i < n; i++) {
      double s = g[i];
      for (int j = 0; j < n; j++) { K[i * N + j] = G[i * n + j]; s += G[i * n + j] * x[j]; }
      rhs[i] = -s;
    }
    for (int w = 0; w < nw; w++) {
      for (int j = 0; j < n; j++) { K[(n + w) * N + j] = ca[W[w] * n + j]; K[j * N + n + w] = ca[W[w] * n + j]; }
      rhs[n + w] = 0;
    }
    int singular = gauss_solve(N, K, rhs);
    for (int attempt = 0; singular && attempt < 2; attempt++) {
      /* singular KKT = linearly dependent working set (e.g. beta_x <= mu beta_z, beta_z >= 0 and
       * beta_x >= 0 all active at the apex of a friction pyramid): re-solve with a tiny dual
       * regularisation, which splits the multipliers among the dependent rows */
      const double eps = attempt == 0 ? 1e-10 : 1e-7;
      for (int i = 0; i < N * N; i++) K[i] = 0;
      for (int i = 0; i < n; i++) {
        double s = g[i];
        for (int j = 0; j < n; j++) { K[i * N + j] = G[i * n + j]; s += G[i * n + j] * x[j]; }
        rhs[i] = -s;
      }
      for (int w = 0; w < nw; w++) {
        for (int j = 0; j < n; j++) { K[(n + w) * N + j] = ca[W[w] * n + j]; K[j * N + n + w] = ca[W[w] * n + j]; }
        K[(n + w) * N + n + w] = -eps;
        rhs[n + w] = 0;
      }
      singular = gauss_solve(N, K, rhs);
    }
    if (singular) { rc = -2; break; }
    double pn = 0, xn = 1;
    for (int i = 0; i < n; i++) { pn = fmax(pn, fabs(rhs[i])); xn = fmax(xn, fabs(x[i])); }
    if (pn <= 1e-9 * xn) { /* cond(G) ~ 1e10: the Newton step of the working set carries ~1e-6 relative noise */
      /* stationary on the working set: check multipliers (lambda >= 0 for a'x <= b) */
      int worst = -1; double wv = -1e-10;
      for (int w = 0; w < nw; w++) if (rhs[n + w] < wv) { wv = rhs[n + w]; worst = w; }
      if (worst < 0) { rc = iter; break; }
      for (int w = worst; w < nw - 1; w++) W[w] = W[w + 1];
      nw--;
      continue;
    }
    double alpha = 1.0; int block = -1;
    for (int k2 = 0; k2 < nk; k2++) {
      int inW = 0;
      for (int w = 0; w < nw; w++) if (W[w] == k2) inW = 1;
      if (inW) continue;
      double ap = 0, ax = 0, an = 0;
      for (int j = 0; j < n; j++) { ap += ca[k2 * n + j] * rhs[j]; ax += ca[k2 * n + j] * x[j]; an = fmax(an, fabs(ca[k2 * n + j])); }
      if (ap > 1e-13 && ap > 1e-11 * pn * an) { /* rounding noise of a large step must not look like motion */
        double a = (cb[k2] - ax) / ap;
        if (a < 0) a = 0;
        if (a < alpha) { alpha = a; block = k2; }
      }
    }
    for (int i = 0; i < n; i++) x[i] += alpha * rhs[i];
    if (block >= 0) W[nw++] = block;
  }
  if (ws && nws) { *nws = nw; for (int w = 0; w < nw; w++) ws[w] = W[w]; }
  free(ca); free(cb); free(W); free(K); free(rhs);
  return rc;
}

/* ------------------------------------------------------------------ facade */
struct orc_cassie {
  const orc_model* phys;
  const orc_model* rbdl;
  orc_data* d;
  orc_kin kin;       /* RBDL state: set at the START of each Step* (Cassie2d.cpp:88,98,121,182) */
  double last_u[NUU];
  double osc_x[NX], osc_obj;
  int osc_iters;
  double qp_z[26]; int qp_ws[128], qp_nws, qp_started; /* QP hot-start state (survives Reset, App. D.3) */
  int contact_sites[NCON]; /* {2,3,4,5} Cassie2d.cpp:34 */
  int target_sites[5];     /* {1..5}    Cassie2d.cpp:35-36 */
};

/* Cassie2d::Cassie2d, Cassie2d.cpp:29-72 */
orc_cassie* orc_cassie_new(const orc_model* phys, const orc_model* rbdl) {
  orc_cassie* c = (orc_cassie*)calloc(1, sizeof(orc_cassie));
  c->phys = phys; c->rbdl = rbdl;
  c->d = orc_data_new(phys);
  for (int i = 0; i < NCON; i++) c->contact_sites[i] = 2 + i;
  for (int i = 0; i < 5; i++) c->target_sites[i] = 1 + i;
  /* constructor standing pose, Cassie2d.cpp:56-58 (right leg differs from the Python reset
   * pose in the 8th digit) */
  static const double qpos_init[NQ] = {0.0, 0.939, 0.0,
      0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
      0.68111815, -1.40730353, 1.62972043, -1.77611107, -0.61968402};
  double zero[NQ] = {0};
  orc_set_state(c->d, qpos_init, zero);
  orc_forward(phys, c->d, NULL);                       /* mj_forward, :62 */
  orc_kin_update(rbdl, &c->kin, c->d->qpos, c->d->qvel); /* setState, :64 */
  return c;
}
void orc_cassie_free(orc_cassie* c) { orc_data_free(c->d); free(c); }
orc_data* orc_cassie_data(orc_cassie* c) { return c->d; }
void orc_cassie_last_ctrl(const orc_cassie* c, double u[6]) { memcpy(u, c->last_u, sizeof(double) * NUU); }

/* StateGeneral <-> arrays, RobotInterface.h:98-128.  26-vector = struct memory order:
 * base_pos[3] base_vel[3] left_pos[5] left_vel[5] right_pos[5] right_vel[5] */
static void state26_to_arrays(const double s[26], double qpos[NQ], double qvel[NQ]) {
  for (int i = 0; i < 3; i++) { qpos[i] = s[i]; qvel[i] = s[3 + i]; }
  for (int i = 0; i < 5; i++) { qpos[3 + i] = s[6 + i]; qvel[3 + i] = s[11 + i]; qpos[8 + i] = s[16 + i]; qvel[8 + i] = s[21 + i]; }
}
/* Reset, Cassie2d.cpp:78-82: writes qpos/qvel and runs mj_forward; time, warm start and the
 * RBDL state are NOT touched */
void orc_cassie_reset(orc_cassie* c, const double s[26]) {
  double qpos[NQ], qvel[NQ];
  state26_to_arrays(s, qpos, qvel);
  orc_set_state(c->d, qpos, qvel);
  orc_forward(c->phys, c->d, NULL);
}
void orc_cassie_get_general_state(const orc_cassie* c, double s[26]) {
  const double* qpos = c->d->qpos; const double* qvel = c->d->qvel;
  for (int i = 0; i < 3; i++) { s[i] = qpos[i]; s[3 + i] = qvel[i]; }
  for (int i = 0; i < 5; i++) { s[6 + i] = qpos[3 + i]; s[11 + i] = qvel[3 + i]; s[16 + i] = qpos[8 + i]; s[21 + i] = qvel[8 + i]; }
}

/* GetOperationalSpaceState, Cassie2d.cpp:218-237 + DynamicModel::GetTargetPoints :360-367.
 * Uses the RBDL kinematics stored at the start of the last Step* (one-step lag); pitch and
 * pitch rate come from the CURRENT mjData (:234-235).  The [2] entries of left/right are
 * never written by the reference; ctypes zero-initialises them -> 0 here. */
void orc_cassie_get_op_state(const orc_cassie* c, double s[18]) {
  double x[15], xd[15];
  for (int i = 0; i < 5; i++) {
    int site = c->target_sites[i];
    orc_kin_point_pos_vel_acc(c->rbdl, &c->kin, c->rbdl->site_body[site], c->rbdl->site_pos[site], x + 3 * i, xd + 3 * i, NULL);
  }
  memset(s, 0, sizeof(double) * 18);
  for (int i = 0; i < 2; i++) {
    s[0 + i] = x[i * 2];
    s[3 + i] = xd[i * 2];
    s[6 + i] = (x[i * 2 + 3] + x[i * 2 + 6]) / 2.0;
    s[9 + i] = (xd[i * 2 + 3] + xd[i * 2 + 6]) / 2.0;
    s[12 + i] = (x[i * 2 + 9] + x[i * 2 + 12]) / 2.0;
    s[15 + i] = (xd[i * 2 + 9] + xd[i * 2 + 12]) / 2.0;
  }
  s[2] = c->d->qpos[2];
  s[5] = c->d->qvel[2];
}

/* DynamicState::UpdateDynamicState, DynamicState.cpp:45-91 (stiff == spring: no spring joints) */
typedef struct {
  double M[NQ * NQ], bias[NQ], Bt[NQ * NUU], Jc[12 * NQ], Jeq[NEQR * NQ], JeqdotQdot[NEQR];
} dyn_mats;

static void update_dynamic_state(const orc_cassie* c, dyn_mats* D) {
  const orc_model* m = c->rbdl;
  const orc_kin* k = &c->kin;
  orc_kin_mass_matrix(m, k, D->M);                    /* GetMassMatrix, DynamicModel.cpp:267-272 */
  orc_kin_nonlinear_effects(m, k, D->bias);           /* GetBiasForce :320-323 */
  for (int i = 0; i < NQ; i++) D->bias[i] -= -m->jnt_damping[i] * k->qd[i]; /* bias -= passive */
  memset(D->Bt, 0, sizeof(D->Bt));                    /* selector = gear, :193-211 */
  for (int a = 0; a < m->nu; a++) D->Bt[m->act_jnt[a] * NUU + a] = m->act_gear[a];
  double jp[3 * NQ];
  for (int i = 0; i < NCON; i++) {
    int s = c->contact_sites[i];
    orc_kin_point_jacobian(m, k, m->site_body[s], m->site_pos[s], jp, NULL);
    memcpy(D->Jc + 3 * i * NQ, jp, sizeof(jp));
  }
  for (int e = 0; e < m->neq; e++) {                  /* GetConstraintJacobian/Accel :274-295 */
    double j1[3 * NQ], j2[3 * NQ], a1[3], a2[3];
    orc_kin_point_jacobian(m, k, m->eq_body1[e], m->eq_anchor1[e], j1, NULL);
    orc_kin_point_jacobian(m, k, m->eq_body2[e], m->eq_anchor2[e], j2, NULL);
    for (int i = 0; i < 3 * NQ; i++) D->Jeq[3 * e * NQ + i] = j1[i] - j2[i];
    orc_kin_point_pos_vel_acc(m, k, m->eq_body1[e], m->eq_anchor1[e], NULL, NULL, a1);
    orc_kin_point_pos_vel_acc(m, k, m->eq_body2[e], m->eq_anchor2[e], NULL, NULL, a2);
    for (int i = 0; i < 3; i++) D->JeqdotQdot[3 * e + i] = a1[i] - a2[i];
  }
}
void orc_cassie_dynamic_state(const orc_cassie* c, double* M, double* bias, double* Bt, double* Jc,
                              double* Jeq, double* JeqdotQdot) {
  dyn_mats D;
  update_dynamic_state(c, &D);
  if (M) memcpy(M, D.M, sizeof(D.M));
  if (bias) memcpy(bias, D.bias, sizeof(D.bias));
  if (Bt) memcpy(Bt, D.Bt, sizeof(D.Bt));
  if (Jc) memcpy(Jc, D.Jc, sizeof(D.Jc));
  if (Jeq) memcpy(Jeq, D.Jeq, sizeof(D.Jeq));
  if (JeqdotQdot) memcpy(JeqdotQdot, D.JeqdotQdot, sizeof(D.JeqdotQdot));
}

static void set_rbdl_state(orc_cassie* c) { orc_kin_update(c->rbdl, &c->kin, c->d->qpos, c->d->qvel); }

/* Cassie2d::Step, Cassie2d.cpp:86-94 */
void orc_cassie_step_torque(orc_cassie* c, const double u[6]) {
  set_rbdl_state(c);
  memcpy(c->last_u, u, sizeof(double) * NUU);
  orc_step(c->phys, c->d, u);
}
/* Cassie2d::StepPd, Cassie2d.cpp:96-117 */
void orc_cassie_step_pd(orc_cassie* c, const double angles[6]) {
  set_rbdl_state(c);
  const double kp = 10.0, kd = 5.0;
  static const int joints[NUU] = {3, 4, 6, 8, 9, 11};
  double u[NUU];
  for (int i = 0; i < NUU; i++)
    u[i] = kp * (angles[i] - c->d->qpos[joints[i]]) + kd * (0.0 - c->d->qvel[joints[i]]);
  memcpy(c->last_u, u, sizeof(u));
  orc_step(c->phys, c->d, u);
}

static void matmul(int n, int k, int m, const double* A, const double* B, double* C) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) {
      double s = 0;
      for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * m + j];
      C[i * m + j] = s;
    }
}
static void transpose(int n, int m, const double* A, double* At) {
  for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) At[j * n + i] = A[i * m + j];
}

/* Nc = I - Jeq^T (Jeq M^-1 Jeq^T)^+ Jeq M^-1 ; gamma = Jeq^T (..)^+ JeqdotQdot
 * (Cassie2d.cpp:132-137 == OSC_RBDL.cpp:169-174).  M^-1 by Cholesky solves (Eigen inverse()). */
static void constraint_projector(const dyn_mats* D, double* Minv, double* Nc, double* gamma) {
  double L[NQ * NQ], e[NQ], col[NQ];
  orc_chol(NQ, D->M, L);
  for (int j = 0; j < NQ; j++) {
    memset(e, 0, sizeof(e)); e[j] = 1;
    orc_chol_solve(NQ, L, e, col);
    for (int i = 0; i < NQ; i++) Minv[i * NQ + j] = col[i];
  }
  double JH[NEQR * NQ], JeqT[NQ * NEQR], JHJ[NEQR * NEQR], JHJp[NEQR * NEQR], T1[NQ * NEQR], T2[NQ * NQ];
  matmul(NEQR, NQ, NQ, D->Jeq, Minv, JH);
  transpose(NEQR, NQ, D->Jeq, JeqT);
  matmul(NEQR, NQ, NEQR, JH, JeqT, JHJ);
  orc_pinv(NEQR, NEQR, JHJ, 1e-3, JHJp, NULL);
  matmul(NQ, NEQR, NEQR, JeqT, JHJp, T1);
  matmul(NQ, NEQR, NQ, T1, JH, T2);
  for (int i = 0; i < NQ; i++) for (int j = 0; j < NQ; j++) Nc[i * NQ + j] = (i == j ? 1.0 : 0.0) - T2[i * NQ + j];
  matmul(NQ, NEQR, 1, T1, D->JeqdotQdot, gamma);
}

/* Cassie2d::StepJacobian, Cassie2d.cpp:119-177 */
void orc_cassie_step_jacobian(orc_cassie* c, const double fin[6]) {
  set_rbdl_state(c);
  dyn_mats D;
  update_dynamic_state(c, &D);
  double Minv[NQ * NQ], Nc[NQ * NQ], gamma[NQ];
  constraint_projector(&D, Minv, Nc, gamma);
  /* Jc6: per foot, mean of the two 6-D site Jacobians [angular; linear] (:139-154) */
  double Jc6[12 * NQ];
  memset(Jc6, 0, sizeof(Jc6));
  const orc_model* m = c->rbdl;
  for (int i = 0; i < 4; i++) {
    int s = c->contact_sites[i];
    double jp[3 * NQ], jr[3 * NQ];
    orc_kin_point_jacobian(m, &c->kin, m->site_body[s], m->site_pos[s], jp, jr);
    int off = (i < 2 ? 0 : 6) * NQ;
    for (int k = 0; k < 3 * NQ; k++) { Jc6[off + k] += jr[k] / 2.0; Jc6[off + 3 * NQ + k] += jp[k] / 2.0; }
  }
  double f[12] = {0};
  f[1] = fin[2]; f[3] = fin[0]; f[5] = fin[1];   /* left: My, Fx, Fz  (:157-160) */
  f[7] = fin[5]; f[9] = fin[3]; f[11] = fin[4];  /* right */
  double rhs[NQ], Jf[NQ], tmp[NQ], NcBt[NQ * NUU], P[NUU * NQ], u[NUU];
  for (int j = 0; j < NQ; j++) { double s = 0; for (int r = 0; r < 12; r++) s += Jc6[r * NQ + j] * f[r]; Jf[j] = s; }
  for (int i = 0; i < NQ; i++) tmp[i] = D.bias[i] - Jf[i];
  matmul(NQ, NQ, 1, Nc, tmp, rhs);
  for (int i = 0; i < NQ; i++) rhs[i] += gamma[i];
  matmul(NQ, NQ, NUU, Nc, D.Bt, NcBt);
  orc_pinv(NQ, NUU, NcBt, 1e-4, P, NULL);        /* default tolerance, :165 */
  matmul(NUU, NQ, 1, P, rhs, u);
  memcpy(c->last_u, u, sizeof(u));
  orc_step(c->phys, c->d, u);
}

/* Cassie2d::StepOsc (Cassie2d.cpp:179-209) + OSC_RBDL::RunPTSC/SolveQP (OSC_RBDL.cpp:114-291) */
void orc_cassie_step_osc(orc_cassie* c, const double act[7]) {
  set_rbdl_state(c);
  dyn_mats D;
  update_dynamic_state(c, &D);
  const orc_model* m = c->rbdl;
  /* xdd (16): Cassie2d.cpp:185-193 */
  double xdd[NTASK] = {0};
  for (int i = 0; i < 2; i++) {
    xdd[i * 2] = act[i];
    xdd[3 + i * 2] = xdd[6 + i * 2] = act[2 + i];
    xdd[9 + i * 2] = xdd[12 + i * 2] = act[4 + i];
  }
  xdd[15] = act[6];
  /* A, AdotQdot: OSC_RBDL.cpp:123-144 */
  double A[NTASK * NQ] = {0}, AdQd[NTASK] = {0};
  for (int i = 0; i < 5; i++) {
    int s = c->target_sites[i];
    double jp[3 * NQ], acc[3];
    orc_kin_point_jacobian(m, &c->kin, m->site_body[s], m->site_pos[s], jp, NULL);
    orc_kin_point_pos_vel_acc(m, &c->kin, m->site_body[s], m->site_pos[s], NULL, NULL, acc);
    memcpy(A + 3 * i * NQ, jp, sizeof(jp));
    for (int r = 0; r < 3; r++) AdQd[3 * i + r] = acc[r];
  }
  A[15 * NQ + 2] = 1.0; /* AddQDDIdx(2), Cassie2d.cpp:41 */
  /* W: OSC_RBDL.cpp:32-38, OSC_RBDL.h:92-95 (all contacts desired -> stance weights) */
  double W[NTASK];
  for (int i = 0; i < 3; i++) W[i] = 5.0;
  for (int i = 3; i < 15; i++) W[i] = 10.0;
  W[15] = 0.1;
  double Minv[NQ * NQ], Nc[NQ * NQ], gamma[NQ];
  constraint_projector(&D, Minv, Nc, gamma);
  /* V (12 x 20): OSC_RBDL.cpp:41-50 ; beta per contact = (x-, x+, y-, y+, z) */
  double V[12 * 20] = {0};
  for (int i = 0; i < NCON; i++) {
    for (int j = 0; j < 2; j++) { V[(i * 3 + j) * 20 + i * 5 + j * 2] = -1.0; V[(i * 3 + j) * 20 + i * 5 + j * 2 + 1] = 1.0; }
    V[(i * 3 + 2) * 20 + i * 5 + 4] = 1.0;
  }
  /* CE = [M, -Nc Bt, -Nc Jc^T V], ce = -Nc bias - gamma  (:180-184) */
  double NcBt[NQ * NUU], JcT[NQ * 12], JcTV[NQ * 20], NcJV[NQ * 20], ce[NQ], Ncb[NQ];
  matmul(NQ, NQ, NUU, Nc, D.Bt, NcBt);
  transpose(12, NQ, D.Jc, JcT);
  matmul(NQ, 12, 20, JcT, V, JcTV);
  matmul(NQ, NQ, 20, Nc, JcTV, NcJV);
  matmul(NQ, NQ, 1, Nc, D.bias, Ncb);
  for (int i = 0; i < NQ; i++) ce[i] = -Ncb[i] - gamma[i];
  /* eliminate qdd = Minv (NcBt u + NcJV beta + ce) =: P z + p0, z = [u; beta] (26) */
  enum { NZ = 26 };
  double Bz[NQ * NZ], P[NQ * NZ], p0[NQ];
  for (int i = 0; i < NQ; i++) {
    for (int j = 0; j < NUU; j++) Bz[i * NZ + j] = NcBt[i * NUU + j];
    for (int j = 0; j < 20; j++) Bz[i * NZ + NUU + j] = NcJV[i * 20 + j];
  }
  matmul(NQ, NQ, NZ, Minv, Bz, P);
  matmul(NQ, NQ, 1, Minv, ce, p0);
  /* H11 = 2 A^T W A ; g1 = 2 A^T W (AdQd - xdd)  (:186,203) */
  double H11[NQ * NQ], g1[NQ];
  for (int i = 0; i < NQ; i++) {
    for (int j = 0; j < NQ; j++) { double s = 0; for (int r = 0; r < NTASK; r++) s += A[r * NQ + i] * W[r] * A[r * NQ + j]; H11[i * NQ + j] = 2 * s; }
    double s = 0;
    for (int r = 0; r < NTASK; r++) s += A[r * NQ + i] * W[r] * (AdQd[r] - xdd[r]);
    g1[i] = 2 * s;
  }
  double HP[NQ * NZ], G[NZ * NZ], gz[NZ], Hp0[NQ];
  matmul(NQ, NQ, NZ, H11, P, HP);
  matmul(NQ, NQ, 1, H11, p0, Hp0);
  for (int i = 0; i < NZ; i++) {
    for (int j = 0; j < NZ; j++) { double s = 0; for (int r = 0; r < NQ; r++) s += P[r * NZ + i] * HP[r * NZ + j]; G[i * NZ + j] = s; }
    double s = 0;
    for (int r = 0; r < NQ; r++) s += P[r * NZ + i] * (Hp0[r] + g1[r]);
    gz[i] = s;
  }
  for (int i = NUU; i < NZ; i++) G[i * NZ + i] += 1e-4; /* m_dWeight_Fx/Fz, :188-201 */
  /* friction pyramid rows C (32 x 26) <= 0, mu = 0.5 (:53-71, RobotInterface.h:64) */
  double C[32 * NZ] = {0}, ubA[32], lbA[32], lb[NZ], ub[NZ];
  for (int i = 0; i < NCON; i++)
    for (int j = 0; j < 2; j++) {
      int l = 0;
      for (int k = 0; k < 4; k++) {
        int row = i * 8 + 4 * j + k;
        if (k % 2) { C[row * NZ + NUU + i * 5 + 2 * j + l] = 1.0; l++; }
        else C[row * NZ + NUU + i * 5 + 2 * j + l] = -1.0;
        C[row * NZ + NUU + (i + 1) * 5 - 1] = -0.5;
      }
    }
  for (int r = 0; r < 32; r++) { ubA[r] = 0; lbA[r] = -1e300; }
  for (int a = 0; a < NUU; a++) { lb[a] = m->act_range[a][0]; ub[a] = m->act_range[a][1]; }
  for (int i = NUU; i < NZ; i++) { lb[i] = 0; ub[i] = 1e300; }
  /* first call: strictly feasible start ("init"); afterwards hot start from the previous solution and
   * working set (OSC_RBDL.cpp:275-278) -- the constraints never change, so it stays feasible */
  double z[NZ] = {0};
  if (!c->qp_started) {
    for (int i = 0; i < NCON; i++) { for (int j = 0; j < 4; j++) z[NUU + i * 5 + j] = 0.1; z[NUU + i * 5 + 4] = 1.0; }
    c->qp_nws = 0;
  } else {
    memcpy(z, c->qp_z, sizeof(z));
  }
  c->osc_iters = qp_solve_ws(NZ, 32, G, gz, C, lbA, ubA, lb, ub, z, c->qp_ws, &c->qp_nws);
  if (c->osc_iters < 0) { /* cap or failure: restart cold once */
    memset(z, 0, sizeof(z));
    for (int i = 0; i < NCON; i++) { for (int j = 0; j < 4; j++) z[NUU + i * 5 + j] = 0.1; z[NUU + i * 5 + 4] = 1.0; }
    c->qp_nws = 0;
    int it2 = qp_solve_ws(NZ, 32, G, gz, C, lbA, ubA, lb, ub, z, c->qp_ws, &c->qp_nws);
    if (it2 >= 0) c->osc_iters = it2;
  }
  memcpy(c->qp_z, z, sizeof(z));
  c->qp_started = 1;
  /* recover x = [qdd; u; beta] and the objective 1/2 x'Hx + g'x */
  double qdd[NQ];
  for (int i = 0; i < NQ; i++) { double s = p0[i]; for (int j = 0; j < NZ; j++) s += P[i * NZ + j] * z[j]; qdd[i] = s; }
  for (int i = 0; i < NQ; i++) c->osc_x[i] = qdd[i];
  for (int i = 0; i < NZ; i++) c->osc_x[NQ + i] = z[i];
  double obj = 0;
  for (int i = 0; i < NQ; i++) { double s = 0; for (int j = 0; j < NQ; j++) s += H11[i * NQ + j] * qdd[j]; obj += qdd[i] * (0.5 * s + g1[i]); }
  for (int i = NUU; i < NZ; i++) obj += 0.5 * 1e-4 * z[i] * z[i];
  c->osc_obj = obj;
  double u[NUU];
  for (int i = 0; i < NUU; i++) u[i] = z[i];
  memcpy(c->last_u, u, sizeof(u));
  orc_step(c->phys, c->d, u);
}
void orc_cassie_osc_last(const orc_cassie* c, double x[39], double* obj, int* iters) {
  memcpy(x, c->osc_x, sizeof(c->osc_x));
  if (obj) *obj = c->osc_obj;
  if (iters) *iters = c->osc_iters;
}

/* ------------------------------------------------------------------ squatting control laws
 * cassie2d.py:263-331 (python) restated so that the CPU baseline has no Python in the loop */
static void squat_jacobian_action(const orc_cassie* c, double zt, double zdt, double f[6]) {
  double s[18];
  orc_cassie_get_op_state(c, s);
  double xt = (s[6] + s[12]) / 2.0;
  double fx = 200.0 * (xt - s[0]) + 50.0 * (0.0 - s[3]);
  double fz = 0.5 * 9.806 * 31.0 + 200.0 * (zt - s[1]) + 50.0 * (zdt - s[4]);
  double my = 100.0 * (0.0 - s[2]) + 10.0 * (0.0 - s[5]);
  if (fz < 0.0) fz = 0.0;
  f[0] = fx; f[1] = fz; f[2] = my; f[3] = fx; f[4] = fz; f[5] = my;
}
static void squat_osc_action(const orc_cassie* c, double zt, double zdt, double a[7]) {
  double s[18];
  orc_cassie_get_op_state(c, s);
  a[2] = 0.0; a[3] = 100.0 * (-5e-3 - s[7]);
  a[4] = 0.0; a[5] = 100.0 * (-5e-3 - s[13]);
  double xt = (s[6] + s[12]) / 2.0;
  a[0] = 100.0 * (xt - s[0]) + 20.0 * (0.0 - s[3]);
  a[1] = 100.0 * (zt - s[1]) + 20.0 * (zdt - s[4]);
  a[6] = 20.0 * (0.0 - s[2]) + 10.0 * (0.0 - s[5]);
}

long orc_rollout(const orc_model* phys, const orc_model* rbdl, int n_envs, int n_steps, int mode,
                 int hold, const double* actions, int adim, const double* phase,
                 const double* init26, double* out_state, int n_threads) {
  long total = 0;
  if (hold < 1) hold = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int e = 0; e < n_envs; e++) {
    orc_cassie* c = orc_cassie_new(phys, rbdl);
    if (init26) orc_cassie_reset(c, init26);
    const double w = 0.5 * 3.1415; /* squatting.py:8-9 */
    double t = 0.0;
    int nact = (n_steps + hold - 1) / hold;
    for (int k = 0; k < n_steps; k++) {
      if (mode == 0) orc_cassie_step_torque(c, actions + ((size_t)e * nact + k / hold) * adim);
      else if (mode == 1) orc_cassie_step_pd(c, actions + ((size_t)e * nact + k / hold) * adim);
      else {
        double ph = phase ? phase[e] : 0.0;
        double zt = 0.7 + 0.25 * sin(w * t + ph), zdt = 0.25 * cos(w * t + ph);
        if (mode == 2) { double f[6]; squat_jacobian_action(c, zt, zdt, f); orc_cassie_step_jacobian(c, f); }
        else { double a[7]; squat_osc_action(c, zt, zdt, a); orc_cassie_step_osc(c, a); }
        t = t + 0.0005;
      }
      total++;
    }
    if (out_state) orc_cassie_get_general_state(c, out_state + 26 * (size_t)e);
    orc_cassie_free(c);
  }
  return total;
}

/* Persistent pool of facades for bench.py --impl reference: the envs (mjData, RBDL state, QP hot start,
 * squat clock) stay alive from one bench "step" to the next, like the reference process that keeps its
 * Cassie2d object for the whole squatting.py loop (squatting.py:6-16). */
struct orc_pool { int n; orc_cassie** c; double* t; };
orc_pool* orc_pool_new(const orc_model* phys, const orc_model* rbdl, int n_envs) {
  orc_pool* p = (orc_pool*)calloc(1, sizeof(orc_pool));
  p->n = n_envs;
  p->c = (orc_cassie**)calloc((size_t)n_envs, sizeof(orc_cassie*));
  p->t = (double*)calloc((size_t)n_envs, sizeof(double));
  for (int e = 0; e < n_envs; e++) p->c[e] = orc_cassie_new(phys, rbdl);
  return p;
}
void orc_pool_free(orc_pool* p) {
  if (!p) return;
  for (int e = 0; e < p->n; e++) orc_cassie_free(p->c[e]);
  free(p->c); free(p->t); free(p);
}
/* n_steps more simulator steps of every env.  mode 0/1: actions [n_envs][nact][adim] held `hold` steps
 * (indexed from step 0 of THIS call); mode 2/3: the squatting laws with per-env phase. */
long orc_pool_run(orc_pool* p, int n_steps, int mode, int hold, const double* actions, int adim,
                  const double* phase, double* out_state, int n_threads) {
  long total = 0;
  if (hold < 1) hold = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int e = 0; e < p->n; e++) {
    orc_cassie* c = p->c[e];
    const double w = 0.5 * 3.1415; /* squatting.py:8-9 */
    double t = p->t[e];
    int nact = (n_steps + hold - 1) / hold;
    for (int k = 0; k < n_steps; k++) {
      if (mode == 0) orc_cassie_step_torque(c, actions + ((size_t)e * nact + k / hold) * adim);
      else if (mode == 1) orc_cassie_step_pd(c, actions + ((size_t)e * nact + k / hold) * adim);
      else {
        double ph = phase ? phase[e] : 0.0;
        double zt = 0.7 + 0.25 * sin(w * t + ph), zdt = 0.25 * cos(w * t + ph);
        if (mode == 2) { double f[6]; squat_jacobian_action(c, zt, zdt, f); orc_cassie_step_jacobian(c, f); }
        else { double a[7]; squat_osc_action(c, zt, zdt, a); orc_cassie_step_osc(c, a); }
        t = t + 0.0005;
      }
      total++;
    }
    p->t[e] = t;
    if (out_state) orc_cassie_get_general_state(c, out_state + 26 * (size_t)e);
  }
  return total;
}
