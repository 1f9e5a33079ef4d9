#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double seed) {
  __shared__ __align__(16) double sm[256];
  sm[threadIdx.x & 255] = seed + threadIdx.x;
  __syncthreads();
  double x = seed, y = seed * 0.5, z = 1.0000001;
  long long t0, t1;
  const int N = 256;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(z), "d"(y));
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(y));
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(z));
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  double c0 = seed, c1 = seed + 1, c2 = seed + 2, c3 = seed + 3, c4 = seed + 4, c5 = seed + 5, c6 = seed + 6, c7 = seed + 7;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N / 8; i++) {
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c0) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c1) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c2) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c3) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c4) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c5) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c6) : "d"(z), "d"(y));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c7) : "d"(z), "d"(y));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  x += c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 64; i++) {
    sm[threadIdx.x & 31] = x;
    __syncwarp();
    double2 v = *reinterpret_cast<const double2*>(&sm[(i * 2) & 30]);
    x = v.x + v.y;
    __syncwarp();
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) * 4;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, i & 31);
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) { __syncthreads(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 128);
  for (int threads : {32, 160}) {
    lat<<<1, threads>>>(out, cyc, 1.37);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("threads=%d per-op cycles: DFMAdep %.1f DADDdep %.1f DMULdep %.1f DFMA8chains %.1f STS-LDS128-roundtrip(+DADD) %.1f SHFL64dep %.1f BAR %.1f\n",
           threads, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 256.0, h[6] / 256.0);
  }
  return 0;
}
