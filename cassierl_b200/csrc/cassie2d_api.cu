// C-ABI of libcassie2d.so (include/cassie2d.h): the reference's ten legacy symbols
// (CassieRL/cassierl src/Cassie2d/Cassie2d.cpp:15-27) as a batch of one, plus the batch entry
// points.  Host glue only -- all arithmetic is in the CUDA kernels (env_kernels.cuh).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <dlfcn.h>

#include "../../include/cassie2d.h"
#include "batch_state.h"
#include "rollout_args.h"
#include "mjcf_flatten.h"

namespace cassie {
static std::atomic<long long> g_launches{0};
long long kernel_launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }
}  // namespace cassie

using namespace cassie;

static thread_local std::string g_last_error;
static int fail(const std::string& msg) {
  g_last_error = msg;
  return -1;
}
#define CU_OK(expr)                                                                              \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

// Constructor pose (Cassie2d.cpp:56-58) and Python reset pose (cassie2d.py:79-85) as StateGeneral
static const double kCtorState26[26] = {0.0, 0.939, 0.0, 0, 0, 0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                                        0, 0, 0, 0, 0, 0.68111815, -1.40730353, 1.62972043, -1.77611107, -0.61968402,
                                        0, 0, 0, 0, 0};
static const double kPyResetState26[26] = {0.0, 0.939, 0.0, 0, 0, 0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                                           0, 0, 0, 0, 0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                                           0, 0, 0, 0, 0};

static std::string default_xml_path() {
  if (const char* e = getenv("CASSIE2D_XML")) return e;
  Dl_info info;
  if (dladdr((void*)&default_xml_path, &info) && info.dli_fname) {
    std::string p = info.dli_fname;  // <pkg>/lib/libcassie2d.so -> <pkg>/model/cassie2d_stiff.xml
    size_t s = p.rfind('/');
    if (s != std::string::npos) {
      p = p.substr(0, s);
      size_t s2 = p.rfind('/');
      if (s2 != std::string::npos) return p.substr(0, s2) + "/model/cassie2d_stiff.xml";
    }
  }
  return "cassie2d_stiff.xml";
}

struct CassieBatch {
  int n = 0, device = 0, precision = 32;
  FlatModels models;
  ModelPair<float> mp32;
  ModelPair<double> mp64;
  BatchView<float> v32;
  BatchView<double> v64;
  std::vector<void*> allocs;
  void* scratch_state = nullptr;   // real [26]: single reset state staging (device)
  // host-variant staging
  void* d_action = nullptr;        // real [n][7]
  void* d_obs = nullptr;           // real [n][26]
  void* d_reward = nullptr;        // real [n]
  uint8_t* d_done = nullptr;       // [n]
  void* d_state26 = nullptr;       // real [n][26]
  void* d_phase = nullptr;         // real [n]
  double* d_traj = nullptr;
  double* d_traj_qvel = nullptr;   // [rows][13] velocities of the reference trajectory (random-phase reset), may be null
  double* d_traj_time = nullptr;   // [rows]
  double* d_sample_time = nullptr; // [n] staging: env clock of the sampled row
  void* d_state18 = nullptr;       // real [n][18] staging of the sampled reset's observation
  cudaStream_t own_stream = nullptr;
  size_t real_size() const { return precision == 64 ? 8 : 4; }
};

template <typename T>
static int alloc_view(CassieBatch* h, BatchView<T>& v) {
  const size_t n = (size_t)h->n;
  auto A = [&](void** p, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) { h->allocs.push_back(*p); e = cudaMemset(*p, 0, bytes); }
    return e;
  };
  v.n = h->n;
  CU_OK(A((void**)&v.qpos, sizeof(T) * 13 * n));
  CU_OK(A((void**)&v.qvel, sizeof(T) * 13 * n));
  CU_OK(A((void**)&v.warm, sizeof(T) * 13 * n));
  CU_OK(A((void**)&v.op, sizeof(T) * 12 * n));
  CU_OK(A((void**)&v.clock, sizeof(double) * n));
  CU_OK(A((void**)&v.jsum0, sizeof(T) * n));
  CU_OK(A((void**)&v.stats, sizeof(int32_t) * 4 * n));
  CU_OK(A((void**)&v.qp_set, sizeof(uint32_t) * n));
  CU_OK(A((void**)&v.ep_len, sizeof(int32_t) * n));
  CU_OK(A((void**)&v.policy_step, sizeof(int32_t) * n));
  v.traj = nullptr; v.traj_rows = 0; v.traj_tmax = 1.0;
  CU_OK(A(&h->scratch_state, sizeof(T) * 26));
  CU_OK(A(&h->d_action, sizeof(T) * 7 * n));
  CU_OK(A(&h->d_obs, sizeof(T) * 26 * n));
  CU_OK(A(&h->d_reward, sizeof(T) * n));
  CU_OK(A((void**)&h->d_done, n));
  CU_OK(A(&h->d_state26, sizeof(T) * 26 * n));
  CU_OK(A(&h->d_phase, sizeof(T) * n));
  return 0;
}

template <typename T> static ModelPair<T>& MP(CassieBatch* h);
template <> ModelPair<float>& MP<float>(CassieBatch* h) { return h->mp32; }
template <> ModelPair<double>& MP<double>(CassieBatch* h) { return h->mp64; }
template <typename T> static BatchView<T>& BV(CassieBatch* h);
template <> BatchView<float>& BV<float>(CassieBatch* h) { return h->v32; }
template <> BatchView<double>& BV<double>(CassieBatch* h) { return h->v64; }

// uploads one StateGeneral (host doubles) into the device scratch in the handle's precision
template <typename T>
static int upload_state(CassieBatch* h, const double* s26, cudaStream_t st) {
  T tmp[26];
  for (int i = 0; i < 26; i++) tmp[i] = (T)s26[i];
  CU_OK(cudaMemcpyAsync(h->scratch_state, tmp, sizeof(tmp), cudaMemcpyHostToDevice, st));
  CU_OK(cudaStreamSynchronize(st));  // tmp is on the stack
  return 0;
}

#define DISPATCH(h, CALL)                          \
  do {                                             \
    if ((h)->precision == 64) { using R = double; CALL; } \
    else { using R = float; CALL; }                \
  } while (0)

static int set_device(const CassieBatch* h) {
  CU_OK(cudaSetDevice(h->device));
  return 0;
}

// ---------------------------------------------------------------------------------------
// Random-phase reset (Cassie3dTraj.sample(), rllab/envs/cassie2d_trajectory.py:26-28: i = randrange(len(time)),
// returns time[i], qpos[i], qvel[i]) on the device: Philox4x32-10 keyed by (seed, global env id, draw) picks the row.
namespace {
__host__ __device__ inline uint32_t philox_first(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {c0, c1, c2, 0x53414d50u};   // "SAMP": a stream of its own, apart from the action noise
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c[0];
}
// StateGeneral order: base pos[3], base vel[3], left pos[5], left vel[5], right pos[5], right vel[5]
template <typename T>
__global__ void k_sample_state26(int n, const double* __restrict__ tq, const double* __restrict__ tv, const double* __restrict__ tt,
                                 int rows, double tmax, uint64_t seed, uint32_t env0, uint32_t draw, const uint8_t* __restrict__ mask,
                                 T* __restrict__ state26, double* __restrict__ time_out, int32_t* __restrict__ index_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || (mask && !mask[e])) return;
  const int i = (int)(philox_first(env0 + (uint32_t)e, draw, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)) % (uint32_t)rows);
  const double* q = tq + (size_t)i * 13;
  T* s = state26 + (size_t)e * 26;
  for (int k = 0; k < 3; k++) { s[k] = (T)q[k]; s[3 + k] = tv ? (T)tv[(size_t)i * 13 + k] : T(0); }
  for (int L = 0; L < 2; L++)
    for (int k = 0; k < 5; k++) {
      s[6 + 10 * L + k] = (T)q[3 + 5 * L + k];
      s[11 + 10 * L + k] = tv ? (T)tv[(size_t)i * 13 + 3 + 5 * L + k] : T(0);
    }
  time_out[e] = tt ? tt[i] : tmax * (double)i / (double)rows;
  if (index_out) index_out[e] = i;
}
// episode bookkeeping of the sampled reset + the observation reset() returns (cassie2d.py:78-95: the 17 pos-invariant
// slots, reference slots 17..25 zero; cassie2d_structs.py:68-75)
template <typename T>
__global__ void k_sample_finish(BatchView<T> v, const double* __restrict__ time_in, const uint8_t* __restrict__ mask, int task,
                                const T* __restrict__ s18, T* __restrict__ obs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= v.n || (mask && !mask[e])) return;
  const size_t n = (size_t)v.n;
  v.clock[e] = time_in[e];
  v.ep_len[e] = 0;
  v.qp_set[e] = 0u;
  for (int i = 0; i < 13; i++) v.warm[(size_t)i * n + e] = T(0);
  v.jsum0[e] = v.qpos[3 * n + e] + v.qpos[4 * n + e] + v.qpos[6 * n + e] + v.qpos[8 * n + e] + v.qpos[9 * n + e] + v.qpos[11 * n + e];
  if (obs) {
    const T* s = s18 + (size_t)e * 18;
    const int od = task == 1 ? 26 : 17;
    T* o = obs + (size_t)e * od;
    for (int i = 0; i < 17; i++) o[i] = s[i + 1];
    o[5] -= s[0];
    o[11] -= s[0];
    for (int i = 17; i < od; i++) o[i] = T(0);
  }
}
}  // namespace

template <typename R>
static int sampled_reset(CassieBatch* h, int task, unsigned long long seed, unsigned int env0, unsigned int draw,
                         const uint8_t* mask_dev, int32_t* index_out_dev, void* obs_dev, cudaStream_t st) {
  const int n = h->n, grid = (n + 127) / 128;
  k_sample_state26<R><<<grid, 128, 0, st>>>(n, h->d_traj, h->d_traj_qvel, h->d_traj_time, BV<R>(h).traj_rows, BV<R>(h).traj_tmax,
                                            (uint64_t)seed, env0, draw, mask_dev, (R*)h->d_state26, h->d_sample_time, index_out_dev);
  count_launch();
  CU_OK(cudaGetLastError());
  CU_OK(Launch<R>::reset(BV<R>(h), (const R*)h->d_state26, 1, mask_dev, st));
  CU_OK(Launch<R>::refresh_op(MP<R>(h), BV<R>(h), mask_dev, st));
  if (obs_dev) CU_OK(Launch<R>::get_op(BV<R>(h), (R*)h->d_state18, st));
  k_sample_finish<R><<<grid, 128, 0, st>>>(BV<R>(h), h->d_sample_time, mask_dev, task, (const R*)h->d_state18, (R*)obs_dev);
  count_launch();
  CU_OK(cudaGetLastError());
  return 0;
}

extern "C" {

const char* CassieGetLastError(void) { return g_last_error.c_str(); }
long long CassieKernelLaunchCount(void) { return kernel_launch_count(); }

CassieBatch* Cassie2dBatchInit(int n_envs, int device, const char* xml_path, int precision) {
  g_last_error.clear();
  if (n_envs <= 0) { fail("n_envs must be positive"); return nullptr; }
  if (precision != 32 && precision != 64) { fail("precision must be 32 or 64"); return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail(std::string("no CUDA device (libcassie2d has no CPU path): ") + cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= ndev) { fail("bad device index"); return nullptr; }
  CassieBatch* h = new CassieBatch();
  h->n = n_envs; h->device = device; h->precision = precision;
  std::string err;
  const std::string path = xml_path && *xml_path ? std::string(xml_path) : default_xml_path();
  if (!flatten_mjcf_file(path, &h->models, &err)) {
    fail("model '" + path + "': " + err);
    delete h;
    return nullptr;
  }
  h->mp32.phys = cast_model<float>(h->models.phys);
  h->mp32.ctrl = cast_model<float>(h->models.ctrl);
  h->mp64.phys = h->models.phys;
  h->mp64.ctrl = h->models.ctrl;
  h->mp32.phys_d = h->models.phys;
  h->mp64.phys_d = h->models.phys;
  h->mp32.ctrl_d = h->models.ctrl;
  h->mp64.ctrl_d = h->models.ctrl;
  if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice failed"); delete h; return nullptr; }
  int rc = precision == 64 ? alloc_view<double>(h, h->v64) : alloc_view<float>(h, h->v32);
  if (rc == 0 && cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail("stream create failed");
  if (rc != 0) { Cassie2dBatchDestroy(h); return nullptr; }
  // constructor: standing pose, mj_forward, setState (Cassie2d.cpp:56-64)
  if (Cassie2dBatchReset(h, nullptr, kCtorState26, nullptr) != 0) { Cassie2dBatchDestroy(h); return nullptr; }
  cudaError_t ce;
  DISPATCH(h, ce = Launch<R>::refresh_op(MP<R>(h), BV<R>(h), nullptr, nullptr));
  if (ce != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    fail(std::string("init launch failed: ") + cudaGetErrorString(ce != cudaSuccess ? ce : cudaGetLastError()));
    Cassie2dBatchDestroy(h);
    return nullptr;
  }
  return h;
}

void Cassie2dBatchDestroy(CassieBatch* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  if (h->d_traj) cudaFree(h->d_traj);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}
int Cassie2dBatchNumEnvs(const CassieBatch* h) { return h ? h->n : -1; }
int Cassie2dBatchPrecision(const CassieBatch* h) { return h ? h->precision : -1; }
int Cassie2dBatchDevice(const CassieBatch* h) { return h ? h->device : -1; }
int Cassie2dBatchRealSize(const CassieBatch* h) { return h ? (int)h->real_size() : -1; }

int Cassie2dBatchReset(CassieBatch* h, const uint8_t* mask_dev, const double* state26_host, void* stream) {
  if (!h) return fail("null handle");
  if (set_device(h)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const double* s = state26_host ? state26_host : kPyResetState26;
  DISPATCH(h, {
    if (upload_state<R>(h, s, st)) return -1;
    CU_OK(Launch<R>::reset(BV<R>(h), (const R*)h->scratch_state, 0, mask_dev, st));
  });
  return 0;
}

int Cassie2dBatchSetState(CassieBatch* h, const void* state26_dev, void* stream) {
  if (!h || !state26_dev) return fail("null argument");
  if (set_device(h)) return -1;
  DISPATCH(h, CU_OK(Launch<R>::reset(BV<R>(h), (const R*)state26_dev, 1, nullptr, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchGetGeneralState(CassieBatch* h, void* state26_dev, void* stream) {
  if (!h || !state26_dev) return fail("null argument");
  if (set_device(h)) return -1;
  DISPATCH(h, CU_OK(Launch<R>::get_general(BV<R>(h), (R*)state26_dev, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchGetOperationalSpaceState(CassieBatch* h, void* state18_dev, void* stream) {
  if (!h || !state18_dev) return fail("null argument");
  if (set_device(h)) return -1;
  DISPATCH(h, CU_OK(Launch<R>::get_op(BV<R>(h), (R*)state18_dev, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchStep(CassieBatch* h, int mode, const void* action_dev, int n_substeps, uint32_t* contact_mask_dev,
                      void* stream) {
  if (!h || !action_dev) return fail("null argument");
  if (mode < 0 || mode > 3) return fail("bad mode");
  if (n_substeps < 0) return fail("n_substeps < 0");
  if (set_device(h)) return -1;
  StepArgs a{mode, n_substeps, action_dev, contact_mask_dev};
  DISPATCH(h, CU_OK(Launch<R>::step(MP<R>(h), BV<R>(h), a, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchEnvStep(CassieBatch* h, int task, int mode, const void* action_dev, int n_substeps, int flags,
                         void* obs_dev, void* reward_dev, uint8_t* done_dev, void* stream) {
  if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev) return fail("null argument");
  if (task != 0 && task != 1) return fail("bad task");
  if (mode != 0 && mode != 1 && mode != 3) return fail("bad mode (the Python envs offer Torque, PD, OSC)");
  if (n_substeps < 0) return fail("n_substeps < 0");
  if (task == 1 && !h->d_traj) return fail("imitation task needs Cassie2dBatchSetTrajectory first");
  if (set_device(h)) return -1;
  EnvStepArgs a{task, mode, n_substeps, flags, action_dev, obs_dev, reward_dev, done_dev};
  DISPATCH(h, CU_OK(Launch<R>::env_step(MP<R>(h), BV<R>(h), a, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchEnvReset(CassieBatch* h, int task, int flags, void* obs_dev, void* stream) {
  if (!h) return fail("null handle");
  if (task == 1 && !h->d_traj) return fail("imitation task needs Cassie2dBatchSetTrajectory first");
  if (set_device(h)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(h, {
    if (upload_state<R>(h, kPyResetState26, st)) return -1;
    CU_OK(Launch<R>::env_reset(MP<R>(h), BV<R>(h), task, flags, (const R*)h->scratch_state, (R*)obs_dev, st));
  });
  return 0;
}

int Cassie2dBatchSetTrajectory(CassieBatch* h, const double* qpos_rows_host, int n_rows, double t_max) {
  if (!h || !qpos_rows_host || n_rows <= 0 || !(t_max > 0)) return fail("bad trajectory");
  if (set_device(h)) return -1;
  CU_OK(cudaDeviceSynchronize());
  if (h->d_traj) cudaFree(h->d_traj);
  CU_OK(cudaMalloc((void**)&h->d_traj, sizeof(double) * 13 * (size_t)n_rows));
  CU_OK(cudaMemcpy(h->d_traj, qpos_rows_host, sizeof(double) * 13 * (size_t)n_rows, cudaMemcpyHostToDevice));
  h->v32.traj = h->v64.traj = h->d_traj;
  h->v32.traj_rows = h->v64.traj_rows = n_rows;
  h->v32.traj_tmax = h->v64.traj_tmax = t_max;
  return 0;
}

int Cassie2dBatchSetTrajectoryDetail(CassieBatch* h, const double* qvel_rows_host, const double* time_rows_host, int n_rows) {
  if (!h) return fail("null handle");
  if (!h->d_traj || n_rows != h->v32.traj_rows) return fail("SetTrajectoryDetail: call Cassie2dBatchSetTrajectory first, with the same row count");
  if (set_device(h)) return -1;
  CU_OK(cudaDeviceSynchronize());
  if (h->d_traj_qvel) { cudaFree(h->d_traj_qvel); h->d_traj_qvel = nullptr; }
  if (h->d_traj_time) { cudaFree(h->d_traj_time); h->d_traj_time = nullptr; }
  if (qvel_rows_host) {
    CU_OK(cudaMalloc((void**)&h->d_traj_qvel, sizeof(double) * 13 * (size_t)n_rows));
    CU_OK(cudaMemcpy(h->d_traj_qvel, qvel_rows_host, sizeof(double) * 13 * (size_t)n_rows, cudaMemcpyHostToDevice));
  }
  if (time_rows_host) {
    CU_OK(cudaMalloc((void**)&h->d_traj_time, sizeof(double) * (size_t)n_rows));
    CU_OK(cudaMemcpy(h->d_traj_time, time_rows_host, sizeof(double) * (size_t)n_rows, cudaMemcpyHostToDevice));
  }
  return 0;
}

int Cassie2dBatchEnvResetSampled(CassieBatch* h, int task, unsigned long long seed, unsigned int first_global_env,
                                 unsigned int draw, const uint8_t* mask_dev, int32_t* index_out_dev, void* obs_dev, void* stream) {
  if (!h) return fail("null handle");
  if (task != 0 && task != 1) return fail("bad task");
  if (!h->d_traj) return fail("random-phase reset needs Cassie2dBatchSetTrajectory first");
  if (set_device(h)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = h->n;
  if (!h->d_sample_time) {
    CU_OK(cudaMalloc((void**)&h->d_sample_time, sizeof(double) * (size_t)n));
    h->allocs.push_back(h->d_sample_time);
    CU_OK(cudaMalloc(&h->d_state18, h->real_size() * 18 * (size_t)n));
    h->allocs.push_back(h->d_state18);
  }
  int rc = 0;
  DISPATCH(h, rc = sampled_reset<R>(h, task, seed, first_global_env, draw, mask_dev, index_out_dev, obs_dev, st));
  return rc;
}

// action-space boxes of the reference envs (cassie_stand2d.py:255-268 == cassie2d.py:353-368)
static void action_box(int mode, double lo[7], double hi[7]) {
  const double d2r = 3.14159265358979323846 / 180.0;
  for (int i = 0; i < 7; i++) { lo[i] = 0; hi[i] = 0; }
  if (mode == 3) {
    const double l[7] = {-2e1, -2e1, -2e1, 0, -2e1, 0, -2e1};
    for (int i = 0; i < 7; i++) { lo[i] = l[i]; hi[i] = 2e1; }
  } else if (mode == 0) {
    const double t[6] = {12.0, 12.0, 0.9, 12.0, 12.0, 0.9};
    for (int i = 0; i < 6; i++) { lo[i] = -t[i]; hi[i] = t[i]; }
  } else {
    const double h6[6] = {80.0, -37.0, -30.0, 80.0, -37.0, -30.0}, l6[6] = {-50.0, -164.0, -140.0, -50.0, -164.0, -140.0};
    for (int i = 0; i < 6; i++) { lo[i] = l6[i] * d2r; hi[i] = h6[i] * d2r; }
  }
}

int Cassie2dBatchRollout(CassieBatch* h, int task, int mode, const void* params_dev, int n_params, int n_policy_steps,
                         int n_substeps, int max_path_length, int flags, int normalize, unsigned long long seed,
                         unsigned int first_global_env, void* obs_dev, void* action_dev, void* mean_dev, void* reward_dev,
                         uint8_t* done_dev, void* stream) {
  if (!h || !params_dev || !obs_dev || !action_dev || !mean_dev || !reward_dev || !done_dev) return fail("null argument");
  if (task != 0 && task != 1) return fail("bad task");
  if (mode != 0 && mode != 1 && mode != 3) return fail("bad mode (the Python envs offer Torque, PD, OSC)");
  if (n_policy_steps < 0 || n_substeps < 0 || max_path_length <= 0) return fail("bad step counts");
  if (task == 1 && !h->d_traj) return fail("imitation task needs Cassie2dBatchSetTrajectory first");
  const int odim = task == 1 ? 26 : 17, adim = mode == 3 ? 7 : 6;
  const int want = odim * 32 + 32 + 32 * 32 + 32 + 32 * adim + adim + adim;
  if (n_params != want) return fail("policy parameter vector has " + std::to_string(n_params) + " entries, expected " + std::to_string(want));
  if (set_device(h)) return -1;
  RolloutArgs a;
  a.task = task; a.mode = mode; a.n_substeps = n_substeps; a.T_steps = n_policy_steps; a.max_path_length = max_path_length;
  a.flags = flags; a.normalize = normalize; a.seed = seed; a.env0 = first_global_env;
  a.params = params_dev; a.obs = obs_dev; a.act = action_dev; a.mean = mean_dev; a.rew = reward_dev; a.done = done_dev;
  action_box(mode, a.act_lo, a.act_hi);
  DISPATCH(h, CU_OK(launch_rollout<R>(MP<R>(h), BV<R>(h), a, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchDiscountedReturns(CassieBatch* h, const void* reward_dev, const uint8_t* done_dev, const void* tail_dev,
                                   double gamma, int n_policy_steps, void* returns_dev, void* stream) {
  if (!h || !reward_dev || !done_dev || !returns_dev) return fail("null argument");
  if (n_policy_steps < 0) return fail("bad step count");
  if (set_device(h)) return -1;
  DISPATCH(h, CU_OK(launch_discounted_returns<R>(reward_dev, done_dev, tail_dev, gamma, n_policy_steps, h->n, returns_dev,
                                                 (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchBaselineMoments(CassieBatch* h, int task, const void* obs_dev, const void* returns_dev, const uint8_t* done_dev,
                                 const int32_t* path_start_dev, int n_policy_steps, int32_t* path_index_dev,
                                 double* moments_dev, void* stream) {
  if (!h || !obs_dev || !returns_dev || !done_dev || !path_index_dev || !moments_dev) return fail("null argument");
  if (task != 0 && task != 1) return fail("bad task");
  if (set_device(h)) return -1;
  BaselineArgs a{};
  a.obs = obs_dev; a.ret = returns_dev; a.done = done_dev; a.start = path_start_dev; a.idx = path_index_dev; a.moments = moments_dev;
  a.odim = task == 1 ? 26 : 17; a.T_steps = n_policy_steps; a.n = h->n;
  DISPATCH(h, CU_OK(launch_baseline_moments<R>(a, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchAdvantages(CassieBatch* h, int task, const void* obs_dev, const void* reward_dev, const uint8_t* done_dev,
                            const int32_t* path_index_dev, const void* coeffs_dev, double gamma, double gae_lambda,
                            int n_policy_steps, void* advantages_dev, void* values_dev, void* stream) {
  if (!h || !obs_dev || !reward_dev || !done_dev || !path_index_dev || !coeffs_dev || !advantages_dev) return fail("null argument");
  if (task != 0 && task != 1) return fail("bad task");
  if (set_device(h)) return -1;
  BaselineArgs a{};
  a.obs = obs_dev; a.rew = reward_dev; a.done = done_dev; a.idx = const_cast<int32_t*>(path_index_dev); a.coeffs = coeffs_dev;
  a.adv = advantages_dev; a.value = values_dev; a.odim = task == 1 ? 26 : 17; a.T_steps = n_policy_steps; a.n = h->n;
  a.gamma = gamma; a.lambda = gae_lambda;
  DISPATCH(h, CU_OK(launch_advantages<R>(a, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchSquat(CassieBatch* h, int mode, int n_steps, const void* phase_dev, uint32_t* contact_mask_dev,
                       void* stream) {
  if (!h) return fail("null handle");
  if (mode != 2 && mode != 3) return fail("squat mode must be JACOBIAN or OSC");
  if (n_steps < 0) return fail("n_steps < 0");
  if (set_device(h)) return -1;
  SquatArgs a{mode, n_steps, phase_dev, contact_mask_dev};
  DISPATCH(h, CU_OK(Launch<R>::squat(MP<R>(h), BV<R>(h), a, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchGetStats(CassieBatch* h, int32_t* stats_dev, void* stream) {
  if (!h || !stats_dev) return fail("null argument");
  if (set_device(h)) return -1;
  const int32_t* src = h->precision == 64 ? h->v64.stats : h->v32.stats;
  CU_OK(cudaMemcpyAsync(stats_dev, src, sizeof(int32_t) * 4 * (size_t)h->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int Cassie2dBatchGetEpisodeLengths(CassieBatch* h, int32_t* ep_len_dev, void* stream) {
  if (!h || !ep_len_dev) return fail("null argument");
  if (set_device(h)) return -1;
  const int32_t* src = h->precision == 64 ? h->v64.ep_len : h->v32.ep_len;
  CU_OK(cudaMemcpyAsync(ep_len_dev, src, sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int Cassie2dBatchSetWarmStart(CassieBatch* h, const void* qacc_dev, void* stream) {
  if (!h || !qacc_dev) return fail("null argument");
  if (set_device(h)) return -1;
  DISPATCH(h, CU_OK(Launch<R>::warm_io(BV<R>(h), (R*)qacc_dev, 1, (cudaStream_t)stream)));
  return 0;
}
int Cassie2dBatchGetWarmStart(CassieBatch* h, void* qacc_dev, void* stream) {
  if (!h || !qacc_dev) return fail("null argument");
  if (set_device(h)) return -1;
  DISPATCH(h, CU_OK(Launch<R>::warm_io(BV<R>(h), (R*)qacc_dev, 0, (cudaStream_t)stream)));
  return 0;
}

int Cassie2dBatchSync(CassieBatch* h) {
  if (!h) return fail("null handle");
  if (set_device(h)) return -1;
  CU_OK(cudaDeviceSynchronize());
  return 0;
}

// ---- host-buffer variants: H2D + launch + D2H + sync on the handle's own stream
int Cassie2dBatchStepHost(CassieBatch* h, int mode, const void* action_host, int n_substeps, void* state26_host) {
  if (!h || !action_host) return fail("null argument");
  if (mode < 0 || mode > 3) return fail("bad mode");
  if (set_device(h)) return -1;
  const size_t rs = h->real_size(), adim = mode == 3 ? 7 : 6;
  cudaStream_t st = h->own_stream;
  CU_OK(cudaMemcpyAsync(h->d_action, action_host, rs * adim * h->n, cudaMemcpyHostToDevice, st));
  if (Cassie2dBatchStep(h, mode, h->d_action, n_substeps, nullptr, st)) return -1;
  if (state26_host) {
    if (Cassie2dBatchGetGeneralState(h, h->d_state26, st)) return -1;
    CU_OK(cudaMemcpyAsync(state26_host, h->d_state26, rs * 26 * h->n, cudaMemcpyDeviceToHost, st));
  }
  CU_OK(cudaStreamSynchronize(st));
  return 0;
}

int Cassie2dBatchEnvStepHost(CassieBatch* h, int task, int mode, const void* action_host, int n_substeps, int flags,
                             void* obs_host, void* reward_host, uint8_t* done_host) {
  if (!h || !action_host || !obs_host || !reward_host || !done_host) return fail("null argument");
  if (mode != 0 && mode != 1 && mode != 3) return fail("bad mode (the Python envs offer Torque, PD, OSC)");
  if (set_device(h)) return -1;
  const size_t rs = h->real_size(), adim = mode == 3 ? 7 : 6, odim = task == 1 ? 26 : 17;
  cudaStream_t st = h->own_stream;
  CU_OK(cudaMemcpyAsync(h->d_action, action_host, rs * adim * h->n, cudaMemcpyHostToDevice, st));
  if (Cassie2dBatchEnvStep(h, task, mode, h->d_action, n_substeps, flags, h->d_obs, h->d_reward, h->d_done, st)) return -1;
  CU_OK(cudaMemcpyAsync(obs_host, h->d_obs, rs * odim * h->n, cudaMemcpyDeviceToHost, st));
  CU_OK(cudaMemcpyAsync(reward_host, h->d_reward, rs * h->n, cudaMemcpyDeviceToHost, st));
  CU_OK(cudaMemcpyAsync(done_host, h->d_done, (size_t)h->n, cudaMemcpyDeviceToHost, st));
  CU_OK(cudaStreamSynchronize(st));
  return 0;
}

int Cassie2dBatchSquatHost(CassieBatch* h, int mode, int n_steps, const void* phase_host, void* state26_host) {
  if (!h) return fail("null handle");
  if (set_device(h)) return -1;
  const size_t rs = h->real_size();
  cudaStream_t st = h->own_stream;
  if (phase_host) CU_OK(cudaMemcpyAsync(h->d_phase, phase_host, rs * h->n, cudaMemcpyHostToDevice, st));
  if (Cassie2dBatchSquat(h, mode, n_steps, phase_host ? h->d_phase : nullptr, nullptr, st)) return -1;
  if (state26_host) {
    if (Cassie2dBatchGetGeneralState(h, h->d_state26, st)) return -1;
    CU_OK(cudaMemcpyAsync(state26_host, h->d_state26, rs * 26 * h->n, cudaMemcpyDeviceToHost, st));
  }
  CU_OK(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------ FP32 peak probe
}  // extern "C"

__global__ void __launch_bounds__(256) k_fp32_probe(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.9999f, c = 1e-4f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" {

double CassieMeasureFp32Peak(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice failed"); return -1.0; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail("no device properties"); return -1.0; }
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float* out = nullptr;
  if (cudaMalloc((void**)&out, sizeof(float) * blocks * threads) != cudaSuccess) { fail("cudaMalloc failed"); return -1.0; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_fp32_probe<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

// ------------------------------------------------------------------ legacy one-env ABI
// Cassie2d.cpp:15-27.  The handle is a batch of one in fp64 (the reference computes in double).
struct Cassie2d {
  CassieBatch* b;
  bool display;
  double* pinned;  // 26 doubles staging
};

static void legacy_die(const char* what) {
  // the reference aborts through mju_error_s on fatal errors (Cassie2d.cpp:49-52)
  fprintf(stderr, "libcassie2d: %s: %s\n", what, g_last_error.c_str());
  abort();
}

Cassie2d* Cassie2dInit(void) {
  int dev = 0;
  if (const char* e = getenv("CASSIE2D_DEVICE")) dev = atoi(e);
  int prec = 64;
  if (const char* e = getenv("CASSIE2D_LEGACY_PRECISION")) prec = atoi(e);
  CassieBatch* b = Cassie2dBatchInit(1, dev, nullptr, prec);
  if (!b) legacy_die("Cassie2dInit");
  Cassie2d* c = new Cassie2d();
  c->b = b;
  c->display = false;
  if (cudaMallocHost((void**)&c->pinned, sizeof(double) * 32) != cudaSuccess) legacy_die("Cassie2dInit(pinned)");
  return c;
}

static void legacy_step(Cassie2d* c, int mode, const double* act, int adim) {
  CassieBatch* b = c->b;
  if (b->precision == 64) {
    memcpy(c->pinned, act, sizeof(double) * adim);
  } else {
    float* f = (float*)c->pinned;
    for (int i = 0; i < adim; i++) f[i] = (float)act[i];
  }
  if (Cassie2dBatchStepHost(b, mode, c->pinned, 1, nullptr)) legacy_die("Step");
}

void Reset(Cassie2d* c, StateGeneral* state) {
  if (Cassie2dBatchReset(c->b, nullptr, (const double*)state, c->b->own_stream) || cudaStreamSynchronize(c->b->own_stream) != cudaSuccess)
    legacy_die("Reset");
}
void StepOsc(Cassie2d* c, ControllerOsc* a) { legacy_step(c, CASSIE_MODE_OSC, (const double*)a, 7); }
void StepTorque(Cassie2d* c, ControllerTorque* a) { legacy_step(c, CASSIE_MODE_TORQUE, a->torques, 6); }
void StepJacobian(Cassie2d* c, ControllerForce* a) { legacy_step(c, CASSIE_MODE_JACOBIAN, (const double*)a, 6); }
void StepPd(Cassie2d* c, ControllerPd* a) { legacy_step(c, CASSIE_MODE_PD, a->angles, 6); }

static void legacy_read(Cassie2d* c, bool op, double* out, int count) {
  CassieBatch* b = c->b;
  cudaStream_t st = b->own_stream;
  int rc = op ? Cassie2dBatchGetOperationalSpaceState(b, b->d_state26, st) : Cassie2dBatchGetGeneralState(b, b->d_state26, st);
  if (rc || cudaMemcpyAsync(c->pinned, b->d_state26, b->real_size() * count, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    legacy_die("GetState");
  if (b->precision == 64) memcpy(out, c->pinned, sizeof(double) * count);
  else for (int i = 0; i < count; i++) out[i] = (double)((float*)c->pinned)[i];
}
void GetGeneralState(Cassie2d* c, StateGeneral* s) { legacy_read(c, false, (double*)s, 26); }
void GetOperationalSpaceState(Cassie2d* c, StateOperationalSpace* s) {
  // the reference leaves left_x[2], left_xd[2], right_x[2], right_xd[2] untouched
  // (Cassie2d.cpp:226-235); keep whatever the caller had there
  double o[18];
  legacy_read(c, true, o, 18);
  double* d = (double*)s;
  for (int i = 0; i < 18; i++)
    if (!(i == 8 || i == 11 || i == 14 || i == 17)) d[i] = o[i];
}
void Display(Cassie2d* c, bool display) { if (c) c->display = display; }
void Render(Cassie2d*) {}

}  // extern "C"
