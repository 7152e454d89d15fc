// Dependent-chain latencies on the GPU at hand: FFMA, DFMA, SHFL (32- and 64-bit), LDS, MUFU.RCP, rsqrt(double), DP div.
// One warp, one block; clock64() around an unrolled dependent chain.  Also DFMA throughput with 8 independent chains
// and with 1..8 warps per SM sub-partition.  usage: ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int MODE> __global__ void k(double* out, long long* cyc, int warps_report) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = threadIdx.x * 1e-3;
  __syncthreads();
  double d = out[0], e = out[0] * 1e-9 + 1.0;
  float f = (float)d, g = (float)e;
  int idx = threadIdx.x & 31;
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) : "d"(d), "f"(f) : "memory");
  if (MODE == 0) { _Pragma("unroll") for (int i = 0; i < N; i++) f = fmaf(f, g, 0.5f); }
  if (MODE == 1) { _Pragma("unroll") for (int i = 0; i < N; i++) d = fma(d, e, 0.5); }
  if (MODE == 2) { _Pragma("unroll") for (int i = 0; i < N; i++) f = __shfl_sync(0xffffffffu, f, (idx + 1) & 3, 4) + 1.0f; }
  if (MODE == 3) { _Pragma("unroll") for (int i = 0; i < N; i++) d = __shfl_sync(0xffffffffu, d, (idx + 1) & 3, 4); }
  if (MODE == 4) { _Pragma("unroll") for (int i = 0; i < N; i++) { idx = (int)sm[idx & 63] & 63; } d = idx; }
  if (MODE == 5) { _Pragma("unroll") for (int i = 0; i < N; i++) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f)); } }
  if (MODE == 6) { _Pragma("unroll") for (int i = 0; i < N / 8; i++) d = rsqrt(d) + 1.0; }
  if (MODE == 7) { _Pragma("unroll") for (int i = 0; i < N / 8; i++) d = 1.0 / d + 1.0; }
  if (MODE == 8) {   // 8 independent DFMA chains
    double a0 = d, a1 = d + 1, a2 = d + 2, a3 = d + 3, a4 = d + 4, a5 = d + 5, a6 = d + 6, a7 = d + 7;
    _Pragma("unroll") for (int i = 0; i < N / 8; i++) {
      a0 = fma(a0, e, 0.5); a1 = fma(a1, e, 0.5); a2 = fma(a2, e, 0.5); a3 = fma(a3, e, 0.5);
      a4 = fma(a4, e, 0.5); a5 = fma(a5, e, 0.5); a6 = fma(a6, e, 0.5); a7 = fma(a7, e, 0.5);
    }
    d = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  }
  if (MODE == 9) { _Pragma("unroll") for (int i = 0; i < N; i++) f = f > 0.5f ? g : f * g; }   // FSEL/FMUL chain
  if (MODE == 10) { _Pragma("unroll") for (int i = 0; i < N; i++) d = d + e; }   // DADD
  if (MODE == 11) { _Pragma("unroll") for (int i = 0; i < N; i++) d = (double)(float)d + 1.0; }   // F2F round trip + DADD
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "d"(d), "f"(f), "r"(idx) : "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x + blockIdx.x * blockDim.x + 1] = d + f + idx;
}
template <int MODE> void run(const char* name, int per, int threads = 32, int blocks = 1) {
  double* out; long long* cyc; cudaMalloc(&out, 8 * (threads * blocks + 2)); cudaMalloc(&cyc, 8);
  cudaMemset(out, 0, 8 * (threads * blocks + 2));
  double one = 1.5; cudaMemcpy(out, &one, 8, cudaMemcpyHostToDevice);
  k<MODE><<<blocks, threads>>>(out, cyc, 0); k<MODE><<<blocks, threads>>>(out, cyc, 0);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %8.1f cycles per op  (threads/block %d)\n", name, (double)c / per, threads);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA dependent", N); run<1>("DFMA dependent", N); run<10>("DADD dependent", N);
  run<2>("SHFL.32 + FADD dependent", N); run<3>("SHFL 64-bit dependent", N); run<4>("LDS dependent (incl. cvt)", N);
  run<5>("MUFU.RCP dependent", N); run<6>("rsqrt(double) + DADD", N / 8); run<7>("1.0/double + DADD", N / 8);
  run<9>("FSETP+FSEL+FMUL chain", N); run<11>("F2F.f32<->f64 round trip + DADD", N);
  run<8>("DFMA 8 independent chains, per DFMA", N, 32); run<8>("  same, 4 warps (1 per SMSP)", N, 128);
  run<8>("  same, 8 warps (2 per SMSP)", N, 256); run<8>("  same, 16 warps (4 per SMSP)", N, 512);
  run<1>("DFMA dependent, 16 warps (4 per SMSP)", N, 512);
  run<0>("FFMA dependent, 16 warps (4 per SMSP)", N, 512);
  return 0;
}
