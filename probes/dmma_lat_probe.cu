// Probe: latency / issue behaviour of mma.sync m16n8k16.f64 on sm_100a.  One CTA on one SM; W warps, each with
// CH independent accumulator chains; cycles per MMA per warp and per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lat_probe dmma_lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int CH>
__global__ void lat(double* out, long long* cyc, int iters, double seed) {
  double a[8], b[4], d[CH][4];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
  for (int j = 0; j < CH; j++) for (int i = 0; i < 4; i++) d[j][i] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < CH; j++) mma16816(d[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < CH; j++) for (int i = 0; i < 4; i++) s += d[j][i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH>
void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  lat<CH><<<1, warps * 32>>>(out, cyc, 10, 1.0);
  lat<CH><<<1, warps * 32>>>(out, cyc, iters, 1.0);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_warp_mma = (double)h / (iters * CH);
  const double warps_per_smsp = warps / 4.0;
  printf("warps=%2d chains=%d : %.1f cycles per MMA per warp, %.1f cycles per MMA per sub-partition\n", warps, CH, per_warp_mma,
         per_warp_mma / (warps_per_smsp < 1 ? 1 : warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
