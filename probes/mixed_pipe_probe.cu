// Probe: do FP64 DMMA (mma.sync m16n8k16.f64) and DFMA issue to the same execution units on sm_100a, or can the
// two pipes run side by side?  Warps 0-3 of every CTA run DMMA chains, warps 4-7 DFMA chains; each half is timed
// alone and then together.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mixed_pipe_probe mixed_pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
// which: bit 0 = DMMA warps work, bit 1 = DFMA warps work
__global__ void __launch_bounds__(256) mixed(double* out, int it_mma, int it_fma, int which, double seed) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp < 4) {
    if (which & 1) {
      double a[8], b[4], d[8][4];
      for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
      for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
      for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) d[j][i] = 0;
      for (int it = 0; it < it_mma; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) mma16816(d[j], a, b);
      }
      for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) s += d[j][i];
    }
  } else if (which & 2) {
    double a[4], b[4], d[8][4];
    for (int i = 0; i < 4; i++) { a[i] = seed + threadIdx.x * 1e-9 + i; b[i] = seed * 0.5 + i; }
    for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) d[j][i] = 0;
    for (int it = 0; it < it_fma; it++) {
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) d[j][i] = fma(a[i], b[i], d[j][i]);
    }
    for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) s += d[j][i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, (size_t)nsm * 4 * 256 * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int it_mma = 4000, it_fma = 4000 * 64;    // 8 MMAs x 4096 flop = 8 x 4 x 64 x (64 DFMA iterations) per loop trip
  for (int cps : {1, 2}) {
    float t[4] = {0, 0, 0, 0};
    for (int which = 1; which <= 3; which++) {
      mixed<<<nsm * cps, 256>>>(out, 100, 100, which, 1.0);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        mixed<<<nsm * cps, 256>>>(out, it_mma, it_fma, which, 1.0);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      t[which] = best;
    }
    const double fl_mma = (double)nsm * cps * 4 * it_mma * 8 * 4096.0, fl_fma = (double)nsm * cps * 4 * it_fma * 8 * 4 * 64.0;
    printf("MIXED ctas/sm=%d  dmma-only %.3f ms (%.2f TF)  dfma-only %.3f ms (%.2f TF)  both %.3f ms (%.2f TF)  sum-of-parts %.3f ms\n",
           cps, t[1], fl_mma / t[1] * 1e-9, t[2], fl_fma / t[2] * 1e-9, t[3], (fl_mma + fl_fma) / t[3] * 1e-9, t[1] + t[2]);
  }
  return 0;
}
