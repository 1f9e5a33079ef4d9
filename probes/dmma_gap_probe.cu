// Probe: do other warps fill the DMMA pipe while a warp is busy with non-MMA work?  Every warp loops
// { NM m16n8k16.f64 MMAs ; a dependent chain of GAP integer multiply-adds (~4 cycles each, no FP64, no memory) } with
// a per-warp phase offset.  W warps on one SM (W/4 per sub-partition).  Reported: DMMA pipe utilisation
// (128 cycles per MMA per sub-partition = 100 %) against the no-contention expectation min(1, (W/4) * t_mma / (t_mma + t_gap)).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_gap_probe dmma_gap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int NM>
__global__ void gap(double* out, long long* cyc, int iters, int gapn, double seed, int useBar) {
  double a[8], b[4], d[4][4];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
  for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) d[j][i] = 0;
  const int warp = threadIdx.x >> 5;
  unsigned x = threadIdx.x * 2654435761u + 12345u;
  // phase offset: warp w starts with a partial gap
  for (int k = 0; k < (gapn * (warp >> 2)) / 4; k++) x = x * 1664525u + 1013904223u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NM; j++) mma16816(d[j & 3], a, b);
    if (useBar) __syncthreads();
#pragma unroll 1
    for (int k = 0; k < gapn; k++) x = x * 1664525u + 1013904223u;       // dependent IMAD chain
    a[0] += (x == 7u) ? 1.0 : 0.0;                                         // keep the chain alive
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) s += d[j][i];
  out[threadIdx.x] = s + x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NM>
void run(int warps, int gapn, int useBar) {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 400;
  gap<NM><<<1, warps * 32>>>(out, cyc, 4, gapn, 1.0, useBar);
  gap<NM><<<1, warps * 32>>>(out, cyc, iters, gapn, 1.0, useBar);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_iter = (double)h / iters;
  const double util = (warps / 4.0) * NM * 128.0 / per_iter;
  printf("warps=%2d MMAs/iter=%d gap=%4d imads bar=%d : %.0f cycles per iteration, DMMA pipe utilisation %.1f %%\n", warps, NM, gapn, useBar,
         per_iter, 100.0 * (util > 1 ? 1 : util));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int bar : {0, 1})
    for (int w : {4, 8, 16})
      for (int g : {0, 100, 250, 500}) run<8>(w, g, bar);
  return 0;
}
