/* TEST INFRASTRUCTURE -- fp64 CPU oracle for the cassie2d step path.  PARITY UNPINNED.
 *
 * This is a from-scratch CPU restatement of the arithmetic behind the reference's
 * libcassie2d.so (CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:15-237), written as a GENERAL
 * 3-D rigid-body pipeline (21 bodies, full 3x3 inertias, 3-row connects, 3-row elliptic
 * contacts) so that it is independent of the product's planar CUDA engine.
 *
 * The arithmetic the reference delegates to un-vendored third parties is restated from
 * their published algorithms (none of them is present in /root/reference or this image):
 *   - MuJoCo 1.50 (mjpro150, closed binary; call sites Cassie2d.cpp:46-62,81,92,115,174,206):
 *     mj_forward / mj_step pipeline restated after the open-source MuJoCo >= 2.1 semantics
 *     (SURVEY.md Appendix B).  1.50 may differ in details (3-parameter solimp sigmoid, PGS
 *     termination); these are listed in DESIGN.md "version hazards".
 *   - RBDL (unpinned; call sites DynamicModel.cpp:241,270,280-282,291-292,322,333,338,343,
 *     364-365): UpdateKinematics, CRBA, NonlinearEffects, CalcPointJacobian(6D),
 *     CalcPointAcceleration, CalcPointVelocity, CalcBodyToBaseCoordinates.
 *   - Eigen JacobiSVD pseudo-inverse (HelperFunctions.h:8-29) and qpOASES 3.2.1
 *     (OSC_RBDL.cpp:238-291).
 * "PARITY UNPINNED": the reference ships no golden simulator outputs and cannot be built
 * here, so this oracle is validated by physical identities and by the reference's few
 * fixtures only (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call this library.  The product (cassierl_b200/) never does.
 */
#ifndef CASSIE_ORACLE_H_
#define CASSIE_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NB 24
#define ORC_NV 24
#define ORC_NQ (ORC_NV + 4)
#define ORC_NG 16
#define ORC_NS 12
#define ORC_NEQ 4
#define ORC_NU 12
#define ORC_MAXCON 32
#define ORC_MAXEFC 128

enum { ORC_GEOM_PLANE = 0, ORC_GEOM_SPHERE = 2, ORC_GEOM_CAPSULE = 3 };
/* ORC_JNT_FREE (cassie3d_stiff.xml:60) expands into six dofs: three world-axis translations (FREE_T) and three
 * rotations about the body's own axes (FREE_R); its seven qpos entries are position + unit quaternion (w x y z) */
enum { ORC_JNT_FREE = 0, ORC_JNT_SLIDE = 2, ORC_JNT_HINGE = 3, ORC_JNT_FREE_T = 10, ORC_JNT_FREE_R = 11 };
enum { ORC_EFC_EQ = 0, ORC_EFC_LIMIT = 1, ORC_EFC_CONTACT = 2 };

typedef struct orc_model orc_model;
typedef struct orc_data orc_data;
typedef struct orc_cassie orc_cassie;

/* ---- model construction (filled by oracle/mjcf_reader.py through ctypes) ---- */
orc_model* orc_model_new(void);
void orc_model_free(orc_model* m);
int orc_add_body(orc_model* m, int parent, const double pos[3], const double mat[9],
                 const double ipos[3], double mass, const double fullinertia[6]);
int orc_add_joint(orc_model* m, int body, int type, const double axis[3], const double pos[3],
                  double ref, int limited, const double range[2], double damping,
                  double armature, const double solref[2], const double solimp[5]);
int orc_add_geom(orc_model* m, int body, int type, const double pos[3], const double mat[9],
                 const double size[2], int contype, int conaffinity, int condim,
                 const double friction[3], const double solref[2], const double solimp[5],
                 double margin, double gap);
int orc_add_site(orc_model* m, int body, const double pos[3]);
int orc_add_connect(orc_model* m, int body1, int body2, const double anchor[3],
                    const double solref[2], const double solimp[5]);
int orc_add_motor(orc_model* m, int joint, double gear, int ctrllimited, const double range[2]);
void orc_set_option(orc_model* m, double timestep, int iterations, double tolerance,
                    double impratio, const double gravity[3]);
/* qpos0 kinematics, connect anchor2, invweight0, meaninertia.  returns 0 on success */
int orc_compile(orc_model* m);
/* model -> RBDL-style controller model (DynamicModel.cpp:84-103: bodies whose last joint has
 * |ref|>=1e-3 lose their xyaxes rotation and the joint angle is used without ref) */
orc_model* orc_model_rbdl_variant(const orc_model* m);

int orc_nv(const orc_model* m);
int orc_nq(const orc_model* m);
int orc_nbody(const orc_model* m);
double orc_total_mass(const orc_model* m);
void orc_get_consts(const orc_model* m, double* body_invweight0 /*nbody*2*/,
                    double* dof_invweight0 /*nv*/, double* meaninertia, double* eq_anchor2 /*neq*3*/);

/* ---- simulation data ---- */
orc_data* orc_data_new(const orc_model* m);
void orc_data_free(orc_data* d);
void orc_set_state(orc_data* d, const double* qpos, const double* qvel);
void orc_get_state(const orc_data* d, double* qpos, double* qvel);
void orc_set_warmstart(orc_data* d, const double* qacc_ws);
void orc_get_warmstart(const orc_data* d, double* qacc_ws);
double orc_get_time(const orc_data* d);
void orc_set_time(orc_data* d, double t);
/* mj_forward / mj_step equivalents; ctrl has nu entries (clamped to ctrlrange inside) */
void orc_forward(const orc_model* m, orc_data* d, const double* ctrl);
void orc_step(const orc_model* m, orc_data* d, const double* ctrl);

/* inspection (all row-major) */
void orc_get_M(const orc_data* d, double* M /*nv*nv*/);
void orc_get_vectors(const orc_data* d, double* qfrc_bias, double* qfrc_passive,
                     double* qfrc_actuator, double* qacc_smooth, double* qacc);
int orc_get_nefc(const orc_data* d);
int orc_get_ncon(const orc_data* d);
int orc_get_solver_iter(const orc_data* d);
void orc_get_efc(const orc_data* d, double* J /*nefc*nv*/, double* pos, double* aref,
                 double* R, double* force, int* type, int* id);
void orc_get_contacts(const orc_data* d, double* dist, double* pos /*ncon*3*/,
                      double* frame /*ncon*9*/, int* geom /*ncon*/);
/* contact bitmask in canonical slot order: geom index g, capsule end e (0 = 'to' end first)
 * -> bit (2*g+e); see DESIGN.md */
unsigned long long orc_contact_mask(const orc_model* m, const orc_data* d);
void orc_body_pose(const orc_data* d, int body, double xpos[3], double xmat[9]);
void orc_site_pos(const orc_model* m, const orc_data* d, int site, double p[3]);
double orc_energy(const orc_model* m, const orc_data* d, double* kinetic, double* potential);

/* ---- RBDL-equivalent kinematics/dynamics on an arbitrary (q, qd) ---- */
typedef struct orc_kin orc_kin;
orc_kin* orc_kin_new(void);
void orc_kin_free(orc_kin* k);
void orc_kin_update(const orc_model* m, orc_kin* k, const double* q, const double* qd);
void orc_kin_mass_matrix(const orc_model* m, const orc_kin* k, double* M);
void orc_kin_nonlinear_effects(const orc_model* m, const orc_kin* k, double* bias);
void orc_kin_point_jacobian(const orc_model* m, const orc_kin* k, int body, const double plocal[3],
                            double* jacp /*3*nv*/, double* jacr /*3*nv*/);
void orc_kin_point_pos_vel_acc(const orc_model* m, const orc_kin* k, int body,
                               const double plocal[3], double pos[3], double vel[3],
                               double acc[3] /* Jdot*qd */);

/* ---- small dense helpers exposed for tests ---- */
void orc_pinv(int rows, int cols, const double* A, double tol, double* Ainv /*cols*rows*/,
              double* sv /*min(rows,cols)*/);
/* strictly convex QP: min 1/2 x'Gx + g'x  s.t. lb<=x<=ub, lbA <= A x <= ubA
 * (dense primal active set; returns #iterations, <0 on failure) */
int orc_qp_solve(int n, int mc, const double* G, const double* g, const double* A,
                 const double* lbA, const double* ubA, const double* lb, const double* ub,
                 double* x);

/* ---- the Cassie2d facade (Cassie2d.cpp:29-237) ---- */
orc_cassie* orc_cassie_new(const orc_model* phys, const orc_model* rbdl);
void orc_cassie_free(orc_cassie* c);
void orc_cassie_reset(orc_cassie* c, const double state[26]);          /* Reset, :78-82 */
void orc_cassie_step_torque(orc_cassie* c, const double u[6]);          /* Step, :86-94 */
void orc_cassie_step_pd(orc_cassie* c, const double angles[6]);         /* StepPd, :96-117 */
void orc_cassie_step_jacobian(orc_cassie* c, const double f[6]);        /* StepJacobian, :119-177 */
void orc_cassie_step_osc(orc_cassie* c, const double a[7]);             /* StepOsc, :179-209 */
void orc_cassie_get_general_state(const orc_cassie* c, double s[26]);   /* :213-216 */
void orc_cassie_get_op_state(const orc_cassie* c, double s[18]);        /* :218-237 */
orc_data* orc_cassie_data(orc_cassie* c);
void orc_cassie_last_ctrl(const orc_cassie* c, double u[6]);
/* DynamicState::UpdateDynamicState outputs (DynamicState.cpp:45-91) at the stored RBDL state */
void orc_cassie_dynamic_state(const orc_cassie* c, double* M /*13x13*/, double* bias /*13*/,
                              double* Bt /*13x6*/, double* Jc /*12x13*/, double* Jeq /*6x13*/,
                              double* JeqdotQdot /*6*/);
/* OSC QP pieces of the last StepOsc (for tests): x (39), objective value */
void orc_cassie_osc_last(const orc_cassie* c, double x[39], double* obj, int* iters);

/* ---- bulk rollouts for the CPU baseline (OpenMP over envs) ----
 * mode: 0 torque (u given per env per step, held `hold` steps), 1 pd, 2 squat-jacobian, 3 squat-osc
 * actions: [n_envs][n_steps/hold][adim] (modes 0,1) ; phase: [n_envs] (modes 2,3)
 * out_state: [n_envs][26] final qpos,qvel.  returns total sim steps done */
long orc_rollout(const orc_model* phys, const orc_model* rbdl, int n_envs, int n_steps, int mode,
                 int hold, const double* actions, int adim, const double* phase,
                 const double* init_state26, double* out_state, int n_threads);
/* 3-D torque rollouts (BASELINE configs[3], cassie3d_stiff.xml: no reference library exists for 3-D):
 * every env starts at (qpos0[nq], qvel0[nv]) (per env), actions [n_envs][n_steps/hold][nu] held `hold` steps; when
 * z_done > 0 an env whose pelvis height qpos[2] is below z_done after a held-action block is reset to
 * (reset_qpos, reset_qvel) with a zero warm start (the rule borrowed from cassie_stand2d.py:131-133).
 * out_qpos [n_envs][nq], out_qvel [n_envs][nv], out_resets [n_envs] (may be NULL).  returns total sim steps */
long orc_rollout_tree(const orc_model* m, int n_envs, int n_steps, int hold, const double* actions,
                      const double* qpos0, const double* qvel0, double z_done, const double* reset_qpos,
                      const double* reset_qvel, double* out_qpos, double* out_qvel, int* out_resets, int n_threads);

/* persistent env pool (bench.py --impl reference): envs, QP hot start and squat clocks survive across calls */
typedef struct orc_pool orc_pool;
orc_pool* orc_pool_new(const orc_model* phys, const orc_model* rbdl, int n_envs);
void orc_pool_free(orc_pool* p);
long orc_pool_run(orc_pool* p, int n_steps, int mode, int hold, const double* actions, int adim,
                  const double* phase, double* out_state, int n_threads);


#ifdef __cplusplus
}
#endif
#endif
