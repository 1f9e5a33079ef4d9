// Cost of the per-step communication skeleton of the panel kernel: STS -> BAR -> 5 LDS -> SHFL chain.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(160) skel(double* out, long long* cyc, int iters) {
  __shared__ double red[2][5][32];
  __shared__ __align__(16) double xs[2][5][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = lane + warp * 0.001, acc = 0.0;
  long long t0 = clock64();
  for (int j = 0; j < iters; j++) {
    const int buf = j & 1;
    if (MODE >= 1) {   // x broadcast + short dot
      double2 x0 = *reinterpret_cast<const double2*>(&xs[buf][warp][(lane >> 3) * 8]);
      double2 x1 = *reinterpret_cast<const double2*>(&xs[buf][warp][(lane >> 3) * 8 + 2]);
      v = fma(x0.x, v, x0.y) + fma(x1.x, v, x1.y);
    }
    if (MODE >= 2) {   // 3-shuffle fold
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
    }
    red[buf][warp][lane] = v;
    __syncthreads();
    double tot = (red[buf][0][lane] + red[buf][1][lane]) + (red[buf][2][lane] + red[buf][3][lane]) + red[buf][4][lane];
    double s = __shfl_sync(0xffffffffu, tot, j & 31);
    v = tot * 1e-3 + s * 1e-4;
    if (MODE == 3 || MODE == 4) {   // sw broadcast shuffles
      double a0 = __shfl_sync(0xffffffffu, v, lane & 7), a1 = __shfl_sync(0xffffffffu, v, (lane & 7) + 8);
      double a2 = __shfl_sync(0xffffffffu, v, (lane & 7) + 16), a3 = __shfl_sync(0xffffffffu, v, (lane & 7) + 24);
      v = (a0 + a1) + (a2 + a3);
    }
    if (MODE == 3 || MODE == 5) {   // publish
      if ((lane & 7) == ((j + 1) & 7)) {
        *reinterpret_cast<double2*>(&xs[buf ^ 1][warp][(lane >> 3) * 8]) = make_double2(v, v);
        *reinterpret_cast<double2*>(&xs[buf ^ 1][warp][(lane >> 3) * 8 + 2]) = make_double2(v, v);
      }
    }
    __syncwarp();
    acc += v;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 160 + threadIdx.x] = acc;
}
template <int MODE> void run(int ctas_per_sm) {
  int nsm = 148, grid = nsm * ctas_per_sm, iters = 4096;
  double* out; long long* cyc;
  cudaMalloc(&out, grid * 160 * 8); cudaMalloc(&cyc, grid * 8);
  skel<MODE><<<grid, 160>>>(out, cyc, iters); cudaDeviceSynchronize();
  skel<MODE><<<grid, 160>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  printf("mode %d ctas/sm %d : %.1f cycles per step per CTA\n", MODE, ctas_per_sm, (double)h[0] / iters);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int c : {1, 3}) { run<0>(c); run<2>(c); run<3>(c); run<4>(c); run<5>(c); }
  return 0;
}
