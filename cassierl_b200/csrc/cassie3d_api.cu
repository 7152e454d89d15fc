// C-ABI of the batched 3-D step path (include/cassie3d.h).  Host glue only: the arithmetic is tree_engine.cuh.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <dlfcn.h>

#include "../../include/cassie3d.h"
#include "mjcf_flatten.h"
#include "tree_kernels.cuh"

using namespace cassie;
using namespace cassie::tree;

static thread_local std::string g3_last_error;
static int fail3(const std::string& msg) {
  g3_last_error = msg;
  return -1;
}
#define CU3_OK(expr)                                                                             \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail3(std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

static std::string default_xml3() {
  if (const char* e = getenv("CASSIE3D_XML")) return e;
  Dl_info info;
  if (dladdr((void*)&default_xml3, &info) && info.dli_fname) {
    std::string p = info.dli_fname;  // <pkg>/lib/libcassie2d.so -> <pkg>/model/cassie3d_stiff.xml
    size_t s = p.rfind('/');
    if (s != std::string::npos) {
      p = p.substr(0, s);
      size_t s2 = p.rfind('/');
      if (s2 != std::string::npos) return p.substr(0, s2) + "/model/cassie3d_stiff.xml";
    }
  }
  return "cassie3d_stiff.xml";
}

struct Cassie3dBatch {
  int n = 0, device = 0, precision = 32, lanes = 32;
  TreeModel<double> model;
  TreeModel<float> m32;
  TreeModel<double> m64;
  TreeBatchView<float> v32;
  TreeBatchView<double> v64;
  std::vector<void*> allocs;
  std::vector<double> reset_q, reset_qd;
  void* d_reset_q = nullptr;
  void* d_reset_qd = nullptr;
  void* d_action = nullptr;   // host-variant staging
  uint8_t* d_done = nullptr;
  cudaStream_t own_stream = nullptr;
  size_t rs() const { return precision == 64 ? 8 : 4; }
};

namespace {
struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

int dmalloc(Cassie3dBatch* h, void** p, size_t bytes) {
  CU3_OK(cudaMalloc(p, bytes));
  h->allocs.push_back(*p);
  CU3_OK(cudaMemset(*p, 0, bytes));
  return 0;
}

template <typename T>
int setup(Cassie3dBatch* h, TreeBatchView<T>& v) {
  const size_t n = (size_t)h->n;
  const TreeModel<double>& m = h->model;
  cast_tree_model(&h->m32, h->model);
  cast_tree_model(&h->m64, h->model);
  v.n = h->n;
  if (dmalloc(h, (void**)&v.qpos, sizeof(T) * n * m.nq) || dmalloc(h, (void**)&v.qvel, sizeof(T) * n * m.nv) ||
      dmalloc(h, (void**)&v.warm, sizeof(T) * n * m.nv) || dmalloc(h, (void**)&v.stats, sizeof(int32_t) * 4 * n) ||
      dmalloc(h, (void**)&v.resets, sizeof(int32_t) * n) ||
      dmalloc(h, (void**)&v.order, sizeof(int32_t) * n) || dmalloc(h, (void**)&v.bins, sizeof(int32_t) * 2 * kTreeBins) ||
      dmalloc(h, (void**)&v.resume, sizeof(int32_t) * n) || dmalloc(h, &h->d_reset_q, sizeof(T) * m.nq) ||
      dmalloc(h, &h->d_reset_qd, sizeof(T) * m.nv) || dmalloc(h, &h->d_action, sizeof(T) * n * m.nu) ||
      dmalloc(h, (void**)&h->d_done, n))
    return -1;
  return 0;
}

template <typename T>
int upload_reset(Cassie3dBatch* h) {
  std::vector<T> q(h->reset_q.begin(), h->reset_q.end()), v(h->reset_qd.begin(), h->reset_qd.end());
  CU3_OK(cudaMemcpy(h->d_reset_q, q.data(), sizeof(T) * q.size(), cudaMemcpyHostToDevice));
  CU3_OK(cudaMemcpy(h->d_reset_qd, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return 0;
}

template <typename T> const TreeModel<T>& model_of(Cassie3dBatch* h);
template <> const TreeModel<float>& model_of<float>(Cassie3dBatch* h) { return h->m32; }
template <> const TreeModel<double>& model_of<double>(Cassie3dBatch* h) { return h->m64; }
template <typename T> TreeBatchView<T>& view(Cassie3dBatch* h);
template <> TreeBatchView<float>& view<float>(Cassie3dBatch* h) { return h->v32; }
template <> TreeBatchView<double>& view<double>(Cassie3dBatch* h) { return h->v64; }

template <typename T>
int step(Cassie3dBatch* h, const void* action, int n_sub, double z_done, int auto_reset, uint8_t* done, cudaStream_t s) {
  TreeStepArgs a;
  a.action = action; a.n_sub = n_sub; a.z_done = z_done; a.auto_reset = auto_reset; a.done = done;
  a.reset_q = h->d_reset_q; a.reset_qd = h->d_reset_qd;
  CU3_OK(TreeLaunch<T>::step(model_of<T>(h), view<T>(h), a, h->lanes, s));
  return 0;
}
}  // namespace

#define DISPATCH3(h, expr32, expr64) ((h)->precision == 64 ? (expr64) : (expr32))

extern "C" {

const char* Cassie3dGetLastError(void) { return g3_last_error.c_str(); }

Cassie3dBatch* Cassie3dBatchCreate(const char* xml_path, int n_envs, int device, int precision) {
  g3_last_error.clear();
  if (n_envs <= 0) { fail3("n_envs must be positive"); return nullptr; }
  if (precision != 32 && precision != 64) { fail3("precision must be 32 or 64"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { fail3("no CUDA device: libcassie2d has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= ndev) { fail3("bad device index"); return nullptr; }
  Cassie3dBatch* h = new Cassie3dBatch();
  h->n = n_envs; h->device = device; h->precision = precision;
  if (const char* e = getenv("CASSIE3D_LANES")) {
    const int l = atoi(e);
    if (l == 8 || l == 16 || l == 32) h->lanes = l;
  }
  std::string err;
  if (!flatten_tree_file(xml_path ? xml_path : default_xml3(), &h->model, &err)) { fail3("model: " + err); delete h; return nullptr; }
  DeviceGuard g(device);
  const int rc = precision == 64 ? setup<double>(h, h->v64) : setup<float>(h, h->v32);
  if (rc) { Cassie3dBatchDestroy(h); return nullptr; }
  // default reset state: the planar constructor pose (Cassie2d.cpp:56-58) on the 3-D joints, abduction = yaw = 0;
  // joints are matched by their MuJoCo order per leg: abduction, yaw, hip, knee, ankle, toe, achilles rod
  const TreeModel<double>& m = h->model;
  h->reset_q.assign(m.nq, 0.0);
  h->reset_qd.assign(m.nv, 0.0);
  for (int i = 0; i < 7; i++) h->reset_q[i] = m.qpos0[i];
  for (int d = 6; d < m.nv; d++) h->reset_q[m.user_dof[d] + 1] = m.qpos0[d + 1];
  if (m.nq == 21) {
    const double leg[7] = {0.0, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407};
    h->reset_q[2] = 0.939;
    for (int i = 0; i < 7; i++) { h->reset_q[7 + i] = leg[i]; h->reset_q[14 + i] = leg[i]; }
  }
  if (DISPATCH3(h, upload_reset<float>(h), upload_reset<double>(h))) { Cassie3dBatchDestroy(h); return nullptr; }
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { fail3("cudaStreamCreate"); Cassie3dBatchDestroy(h); return nullptr; }
  if (Cassie3dBatchResetAll(h, nullptr, nullptr)) { Cassie3dBatchDestroy(h); return nullptr; }
  cudaDeviceSynchronize();
  return h;
}

void Cassie3dBatchDestroy(Cassie3dBatch* h) {
  if (!h) return;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int Cassie3dBatchSizes(Cassie3dBatch* h, int32_t* out) {
  if (!h || !out) return fail3("null argument");
  out[0] = h->model.nq; out[1] = h->model.nv; out[2] = h->model.nu; out[3] = tree::kMaxRows; out[4] = tree::kMaxCon;
  out[5] = (int32_t)(h->precision == 64 ? sizeof(Scratch<double, tree::kFastRows, tree::kFastCon>) : sizeof(Scratch<float, tree::kFastRows, tree::kFastCon>));
  return 0;
}

int Cassie3dBatchSetLanes(Cassie3dBatch* h, int lanes) {
  if (!h) return fail3("null handle");
  if (lanes != 8 && lanes != 16 && lanes != 32) return fail3("lanes must be 8, 16 or 32");
  h->lanes = lanes;
  return 0;
}

int Cassie3dBatchSetResetState(Cassie3dBatch* h, const double* qpos, const double* qvel) {
  if (!h || !qpos || !qvel) return fail3("null argument");
  DeviceGuard g(h->device);
  h->reset_q.assign(qpos, qpos + h->model.nq);
  h->reset_qd.assign(qvel, qvel + h->model.nv);
  cudaDeviceSynchronize();
  return DISPATCH3(h, upload_reset<float>(h), upload_reset<double>(h));
}

int Cassie3dBatchGetResetState(Cassie3dBatch* h, double* qpos, double* qvel) {
  if (!h || !qpos || !qvel) return fail3("null argument");
  memcpy(qpos, h->reset_q.data(), sizeof(double) * h->reset_q.size());
  memcpy(qvel, h->reset_qd.data(), sizeof(double) * h->reset_qd.size());
  return 0;
}

int Cassie3dBatchResetAll(Cassie3dBatch* h, const uint8_t* mask, void* stream) {
  if (!h) return fail3("null handle");
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->precision == 64)
    CU3_OK(TreeLaunch<double>::set_all(h->v64, h->model.nq, h->model.nv, (const double*)h->d_reset_q, (const double*)h->d_reset_qd, mask, s));
  else
    CU3_OK(TreeLaunch<float>::set_all(h->v32, h->model.nq, h->model.nv, (const float*)h->d_reset_q, (const float*)h->d_reset_qd, mask, s));
  return 0;
}

static void* qpos_of(Cassie3dBatch* h) { return h->precision == 64 ? (void*)h->v64.qpos : (void*)h->v32.qpos; }
static void* qvel_of(Cassie3dBatch* h) { return h->precision == 64 ? (void*)h->v64.qvel : (void*)h->v32.qvel; }
static void* warm_of(Cassie3dBatch* h) { return h->precision == 64 ? (void*)h->v64.warm : (void*)h->v32.warm; }
static int32_t* stats_of(Cassie3dBatch* h) { return h->precision == 64 ? h->v64.stats : h->v32.stats; }
static int32_t* resets_of(Cassie3dBatch* h) { return h->precision == 64 ? h->v64.resets : h->v32.resets; }

int Cassie3dBatchSetState(Cassie3dBatch* h, const void* qpos, const void* qvel, void* stream) {
  if (!h || !qpos || !qvel) return fail3("null argument");
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->n;
  CU3_OK(cudaMemcpyAsync(qpos_of(h), qpos, h->rs() * n * h->model.nq, cudaMemcpyDeviceToDevice, s));
  CU3_OK(cudaMemcpyAsync(qvel_of(h), qvel, h->rs() * n * h->model.nv, cudaMemcpyDeviceToDevice, s));
  CU3_OK(cudaMemsetAsync(warm_of(h), 0, h->rs() * n * h->model.nv, s));
  return 0;
}

int Cassie3dBatchGetState(Cassie3dBatch* h, void* qpos, void* qvel, void* stream) {
  if (!h || !qpos || !qvel) return fail3("null argument");
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->n;
  CU3_OK(cudaMemcpyAsync(qpos, qpos_of(h), h->rs() * n * h->model.nq, cudaMemcpyDeviceToDevice, s));
  CU3_OK(cudaMemcpyAsync(qvel, qvel_of(h), h->rs() * n * h->model.nv, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int Cassie3dBatchGetWarmStart(Cassie3dBatch* h, void* w, void* stream) {
  if (!h || !w) return fail3("null argument");
  DeviceGuard g(h->device);
  CU3_OK(cudaMemcpyAsync(w, warm_of(h), h->rs() * (size_t)h->n * h->model.nv, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int Cassie3dBatchSetWarmStart(Cassie3dBatch* h, const void* w, void* stream) {
  if (!h || !w) return fail3("null argument");
  DeviceGuard g(h->device);
  CU3_OK(cudaMemcpyAsync(warm_of(h), w, h->rs() * (size_t)h->n * h->model.nv, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int Cassie3dBatchStep(Cassie3dBatch* h, const void* action, int n_substeps, double z_done, int auto_reset, uint8_t* done,
                      void* stream) {
  if (!h) return fail3("null handle");
  if (n_substeps < 0) return fail3("n_substeps must not be negative");
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  return DISPATCH3(h, step<float>(h, action, n_substeps, z_done, auto_reset, done, s),
                   step<double>(h, action, n_substeps, z_done, auto_reset, done, s));
}

int Cassie3dBatchStepHost(Cassie3dBatch* h, const void* action, int n_substeps, double z_done, int auto_reset,
                          void* qpos_out, void* qvel_out, uint8_t* done_out) {
  if (!h) return fail3("null handle");
  DeviceGuard g(h->device);
  cudaStream_t s = h->own_stream;
  const size_t n = (size_t)h->n;
  if (action) CU3_OK(cudaMemcpyAsync(h->d_action, action, h->rs() * n * h->model.nu, cudaMemcpyHostToDevice, s));
  if (DISPATCH3(h, step<float>(h, action ? h->d_action : nullptr, n_substeps, z_done, auto_reset, h->d_done, s),
                step<double>(h, action ? h->d_action : nullptr, n_substeps, z_done, auto_reset, h->d_done, s)))
    return -1;
  if (qpos_out) CU3_OK(cudaMemcpyAsync(qpos_out, qpos_of(h), h->rs() * n * h->model.nq, cudaMemcpyDeviceToHost, s));
  if (qvel_out) CU3_OK(cudaMemcpyAsync(qvel_out, qvel_of(h), h->rs() * n * h->model.nv, cudaMemcpyDeviceToHost, s));
  if (done_out) CU3_OK(cudaMemcpyAsync(done_out, h->d_done, n, cudaMemcpyDeviceToHost, s));
  CU3_OK(cudaStreamSynchronize(s));
  return 0;
}

int Cassie3dBatchGetStats(Cassie3dBatch* h, int32_t* stats, void* stream) {
  if (!h || !stats) return fail3("null argument");
  DeviceGuard g(h->device);
  CU3_OK(cudaMemcpyAsync(stats, stats_of(h), sizeof(int32_t) * 4 * (size_t)h->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int Cassie3dBatchGetResets(Cassie3dBatch* h, int32_t* resets, void* stream) {
  if (!h || !resets) return fail3("null argument");
  DeviceGuard g(h->device);
  CU3_OK(cudaMemcpyAsync(resets, resets_of(h), sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
