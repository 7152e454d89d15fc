// Lane runtime of the quad engine: FOUR LANES PER ENV (one "quad"), eight envs per warp.
//
// BASELINE.json north_star: "each env maps to one warp or sub-warp using shuffle reductions".  Lane ql = 2 L + h of a
// quad works on leg L (0 left, 1 right) and on half h of whatever the stage splits two ways (right-hand sides of
// the multi-RHS solves, constraint rows, tasks).  Cross-lane traffic is warp shuffles of width 4 and a per-warp
// shared-memory block laid out [field][8 envs] (a field read by all lanes of a warp touches 8 consecutive words:
// conflict free, and broadcast within a quad).
//
// On the host (tests/host_harness/quad_harness.cpp) the four lanes of ONE env run as four threads that meet at a
// spin barrier for every collective, so the very same code is unit-tested against the oracle without a GPU.
#pragma once
#include "planar_engine.cuh"

#if !defined(__CUDACC__)
#include <atomic>
#include <cstring>
#include <thread>
#endif

namespace cassie {
namespace quad {

#if defined(__CUDACC__)
constexpr int kES = 8;   // env stride of shared arrays: element i of env e lives at [i * 8 + e]
#define QUAD_FN __device__ __forceinline__
#define QUAD_NOINLINE __device__ __noinline__

struct Lane {
  int ql, L, h;   // lane in quad, leg, half
};
QUAD_FN Lane lane_id() {
  Lane l;
  l.ql = (int)(threadIdx.x & 3u);
  l.L = l.ql >> 1;
  l.h = l.ql & 1;
  return l;
}
template <typename V> QUAD_FN V shfl(V v, int src) { return __shfl_sync(0xffffffffu, v, src, 4); }
template <typename V> QUAD_FN V shx(V v, int m) { return __shfl_xor_sync(0xffffffffu, v, m, 4); }
QUAD_FN void wsync() { __syncwarp(); }
QUAD_FN bool wany(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// any over the whole CTA (a barrier: every thread of the CTA must call it).  Used for decisions that select between large
// unrolled code regions: the warps of a lock-step CTA share instruction-cache fills only while they run the SAME region.
#ifndef CASSIE_QUAD_CTA_TIER
#define CASSIE_QUAD_CTA_TIER 1
#endif
QUAD_FN bool cta_any(bool p) {
  if (CASSIE_QUAD_CTA_TIER && blockDim.x > 32) return __syncthreads_or(p) != 0;
  return wany(p);
}
// any / all over the four lanes of the quad only
QUAD_FN bool qany(bool p) {
  const unsigned b = __ballot_sync(0xffffffffu, p);
  return ((b >> ((threadIdx.x & 31u) & ~3u)) & 0xfu) != 0u;
}
// 4-bit mask: bit c is set when lane position c of ANY quad of the warp has p (warp uniform)
QUAD_FN unsigned wlanes(bool p) {
  unsigned b = __ballot_sync(0xffffffffu, p);
  b |= b >> 16; b |= b >> 8; b |= b >> 4;
  return b & 0xfu;
}

#else  // ------------------------------------------------------------------ host emulation (one quad)
constexpr int kES = 1;
#define QUAD_FN inline
#define QUAD_NOINLINE inline

struct Lane {
  int ql, L, h;
};
struct Emu {
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  alignas(8) unsigned char slot[4][8];
  bool pred[4];
  void barrier() {
    const int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == 4) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_acq_rel);
    } else {
      int spins = 0;
      while (gen.load(std::memory_order_acquire) == g) {
        if (++spins > 200) { std::this_thread::yield(); spins = 0; }
      }
    }
  }
};
inline Emu& emu() { static Emu e; return e; }
inline int& emu_lane() { static thread_local int l = 0; return l; }
QUAD_FN Lane lane_id() {
  Lane l;
  l.ql = emu_lane();
  l.L = l.ql >> 1;
  l.h = l.ql & 1;
  return l;
}
template <typename V> QUAD_FN V shfl(V v, int src) {
  static_assert(sizeof(V) <= 8, "shuffle payload");
  Emu& e = emu();
  std::memcpy(e.slot[emu_lane()], &v, sizeof(V));
  e.barrier();
  V r;
  std::memcpy(&r, e.slot[src & 3], sizeof(V));
  e.barrier();
  return r;
}
template <typename V> QUAD_FN V shx(V v, int m) { return shfl(v, emu_lane() ^ m); }
QUAD_FN void wsync() { emu().barrier(); }
QUAD_FN bool wany(bool p) {
  Emu& e = emu();
  e.pred[emu_lane()] = p;
  e.barrier();
  const bool r = e.pred[0] || e.pred[1] || e.pred[2] || e.pred[3];
  e.barrier();
  return r;
}
QUAD_FN bool qany(bool p) { return wany(p); }
QUAD_FN bool cta_any(bool p) { return wany(p); }
QUAD_FN unsigned wlanes(bool p) {
  Emu& e = emu();
  e.pred[emu_lane()] = p;
  e.barrier();
  const unsigned r = (e.pred[0] ? 1u : 0u) | (e.pred[1] ? 2u : 0u) | (e.pred[2] ? 4u : 0u) | (e.pred[3] ? 8u : 0u);
  e.barrier();
  return r;
}
// runs fn(lane) on four threads
template <typename F> inline void run_quad(F fn) {
  std::thread th[4];
  for (int l = 0; l < 4; l++) th[l] = std::thread([l, &fn]() { emu_lane() = l; fn(l); });
  for (int l = 0; l < 4; l++) th[l].join();
}
#endif

// Phase barrier of a multi-warp CTA (quad_kernels.cuh): keeps the warps of a CTA in the same code region so that they
// share instruction-cache fills.  LEVEL = how fine a grain the call site is; CASSIE_QUAD_PHASE_SYNC selects up to which
// level barriers are compiled in (0: only the per-step barrier of the kernels).
#ifndef CASSIE_QUAD_PHASE_SYNC
#define CASSIE_QUAD_PHASE_SYNC 0
#endif
template <int LEVEL> QUAD_FN void phase_sync() {
#if defined(__CUDACC__)
  if (LEVEL <= CASSIE_QUAD_PHASE_SYNC && blockDim.x > 32) __syncthreads();
#endif
}

// view of one env's column of an [n][kES] shared array
template <typename V>
struct SV {
  V* p;
  QUAD_FN V& operator[](int i) const { return p[i * kES]; }
  QUAD_FN SV<V> at(int i) const { return SV<V>{p + i * kES}; }
};

// sum over the two legs (lanes ql and ql ^ 2), same value on both afterwards
template <typename V> QUAD_FN V sum_legs(V v) { return v + shx(v, 2); }

}  // namespace quad
}  // namespace cassie
