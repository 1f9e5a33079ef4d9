// Latency probe: dependent-chain latencies of FP64 ops, LDS, SHFL, BAR on sm_100a (clock64 based).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double seed) {
  __shared__ double sm[256];
  sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  double x = seed, y = seed * 0.5, z = 1.0000001;
  long long t0, t1;
  const int N = 256;
  // DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = fma(x, z, y);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / N;
  // DADD chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = x + y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / N;
  // DMUL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = x * z;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / N;
  // 8 independent DFMA chains (per-instr issue cost)
  double c[8]; for (int k = 0; k < 8; k++) c[k] = seed + k;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = fma(c[k], z, y);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0) / N;   // per 8 instr
  for (int k = 0; k < 8; k++) x += c[k];
  // SHFL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) / N;
  // LDS dependent chain (pointer chase)
  int idx = threadIdx.x & 31;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) { double v = sm[idx]; idx = ((int)v + i) & 31; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0) / N;   // includes F2I
  x += idx;
  // __syncthreads
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0) / N;
  // rsqrt.approx.f64 + accuracy
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(seed * 3.0));
  double rc; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(seed * 3.0));
  if (threadIdx.x == 0) { out[1] = r; out[2] = rc; }
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(y));
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0) / N;
  // exact sqrt / div chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 64; i++) y = sqrt(y + 1.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0) / 64;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 64; i++) y = 1.0 / (y + 1.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = (t1 - t0) / 64;
  out[0] = x + y;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 128);
  for (int threads : {32, 160}) {
    lat<<<1, threads>>>(out, cyc, 1.37);
    cudaDeviceSynchronize();
    long long h[10]; double ho[3];
    cudaMemcpy(h, cyc, 80, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 24, cudaMemcpyDeviceToHost);
    printf("threads=%d  DFMA %lld  DADD %lld  DMUL %lld  8xDFMA %lld  SHFL %lld  LDSchase %lld  BAR %lld  RSQ64 %lld  sqrt %lld  div %lld cycles\n",
           threads, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
    double ex = 1.0 / sqrt(1.37 * 3.0);
    printf("rsqrt.approx rel err %.3e   rcp.approx rel err %.3e\n", (ho[1] - ex) / ex, (ho[2] - 1.0 / (1.37 * 3.0)) * (1.37 * 3.0));
  }
  return 0;
}
