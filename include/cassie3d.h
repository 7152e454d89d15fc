/* C-ABI of the batched 3-D Cassie step path (libcassie2d.so, same shared library as include/cassie2d.h).
 *
 * The reference ships NO library for its 3-D model: model/cassie3d_stiff.xml:1-192 is only ever loaded by MuJoCo's own
 * viewer ("No reference library exists for 3-D"; 8(f) row 2: "also a 3-D RobotInterface
 * (none exists)").  These entry points are therefore the batch analogue of what src/Cassie2d/Cassie2d.cpp:15-27 exports
 * for the planar model, restricted to what BASELINE.json configs[3] needs: torque actions (the ten motors of
 * cassie3d_stiff.xml:180-191), mj_step semantics (cassie3d_stiff.xml:5), done when the pelvis height drops below a
 * threshold (rule borrowed from rllab/envs/cassie_stand2d.py:131-133), auto-reset to a standing pose.
 *
 * Conventions: plain pointers and sizes only; device pointers unless a name ends in Host; `real` = float (precision 32)
 * or double (precision 64); qpos [n][nq] and qvel [n][nv] row-major per env in MuJoCo's order (free joint: position,
 * quaternion w x y z; then the hinges in file order); every call is asynchronous on `stream` and returns 0 or -1 with
 * Cassie3dGetLastError(); nothing aborts.  There is no CPU fallback: Create fails without a CUDA device.
 */
#ifndef CASSIE3D_H_
#define CASSIE3D_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Cassie3dBatch Cassie3dBatch;

const char* Cassie3dGetLastError(void);
/* xml_path NULL: the packaged physics-only rendition of cassie3d_stiff.xml (cassierl_b200/model/). */
Cassie3dBatch* Cassie3dBatchCreate(const char* xml_path, int n_envs, int device, int precision);
void Cassie3dBatchDestroy(Cassie3dBatch* h);
/* out[0..5] = nq, nv, nu, constraint-row capacity, contact capacity (per env and step; what exceeds them is dropped in
 * MuJoCo's pair order and counted in the stats), bytes of shared memory per env of the first pass (the step runs a small
 * capacity first and finishes the rare env that needs more in a full-capacity continuation launch: same results) */
int Cassie3dBatchSizes(Cassie3dBatch* h, int32_t* out);
/* lanes per env of the step kernel: 8, 16 or 32 (default: env CASSIE3D_LANES, else 32) */
int Cassie3dBatchSetLanes(Cassie3dBatch* h, int lanes);
/* the state done envs are reset to, and the state SetAll writes (host doubles, qpos[nq], qvel[nv]); the default is
 * the standing pose of Cassie2d.cpp:56-58 carried over to the 3-D joints (abduction = yaw = 0) */
int Cassie3dBatchSetResetState(Cassie3dBatch* h, const double* qpos, const double* qvel);
int Cassie3dBatchGetResetState(Cassie3dBatch* h, double* qpos, double* qvel);
/* every env (mask NULL) or the envs with mask[e] != 0 <- the reset state; warm start cleared */
int Cassie3dBatchResetAll(Cassie3dBatch* h, const uint8_t* mask, void* stream);
/* whole-batch state I/O, real [n][nq] / [n][nv]; SetState clears the warm start */
int Cassie3dBatchSetState(Cassie3dBatch* h, const void* qpos, const void* qvel, void* stream);
int Cassie3dBatchGetState(Cassie3dBatch* h, void* qpos, void* qvel, void* stream);
int Cassie3dBatchGetWarmStart(Cassie3dBatch* h, void* qacc_warmstart, void* stream);
int Cassie3dBatchSetWarmStart(Cassie3dBatch* h, const void* qacc_warmstart, void* stream);
/* n_substeps x mj_step with the controls `action` (real [n][nu], clamped to ctrlrange; NULL = zero) held.  z_done > 0:
 * done[e] = 1 when qpos[2] < z_done after the last substep, 2 when the state is not finite; with auto_reset such envs
 * restart from the reset state inside the same launch.  done may be NULL. */
int Cassie3dBatchStep(Cassie3dBatch* h, const void* action, int n_substeps, double z_done, int auto_reset, uint8_t* done,
                      void* stream);
/* the same with HOST buffers (pinned or pageable): action in, qpos / qvel / done out; copies and the synchronisation
 * are inside the call -- the end-to-end path of bench3d */
int Cassie3dBatchStepHost(Cassie3dBatch* h, const void* action, int n_substeps, double z_done, int auto_reset,
                          void* qpos_out, void* qvel_out, uint8_t* done_out);
/* int32 [n][4]: constraint rows, contacts, PGS sweeps of the last step; contacts dropped for capacity so far */
int Cassie3dBatchGetStats(Cassie3dBatch* h, int32_t* stats, void* stream);
/* int32 [n]: auto-resets so far */
int Cassie3dBatchGetResets(Cassie3dBatch* h, int32_t* resets, void* stream);

#ifdef __cplusplus
}
#endif
#endif
