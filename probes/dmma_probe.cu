// Probe: FP64 mma.sync fragment layouts + throughput of DMMA shapes vs DFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/dmma_probe probes/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void mma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// Layout check: A is 16x16 row-major, B is 16x8 (k x n) row-major, D 16x8.
// Hypothesis H: a_i: row = g + 8*(i&1), col = t + 4*(i>>1); b_i: k = t + 4*i, n = g; d: row g (+8 for i>=2), col 2t+(i&1)
__global__ void layout_k16(const double* A, const double* B, double* D) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[8], b[4], d[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; i++) a[i] = A[(g + 8 * (i & 1)) * 16 + t + 4 * (i >> 1)];
  for (int i = 0; i < 4; i++) b[i] = B[(t + 4 * i) * 8 + g];
  mma16816(d, a, b);
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
  D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}
__global__ void layout_k8(const double* A, const double* B, double* D) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[4], b[2], d[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) a[i] = A[(g + 8 * (i & 1)) * 16 + t + 4 * (i >> 1)];
  for (int i = 0; i < 2; i++) b[i] = B[(t + 4 * i) * 8 + g];
  mma1688(d, a, b);
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
  D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}
__global__ void layout_k4(const double* A, const double* B, double* D) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[2], d[4] = {0, 0, 0, 0};
  for (int i = 0; i < 2; i++) a[i] = A[(g + 8 * (i & 1)) * 16 + t];
  double b = B[t * 8 + g];
  mma1684(d, a, b);
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
  D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}
__global__ void layout_884(const double* A, const double* B, double* D) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double d[2] = {0, 0};
  mma884(d, A[g * 16 + t], B[t * 8 + g]);
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
}

template <int MODE>
__global__ void __launch_bounds__(256) tput(double* out, int iters, double seed) {
  // MODE 0: DFMA, 1: m8n8k4, 2: m16n8k4, 3: m16n8k8, 4: m16n8k16. 8 independent accumulator chains.
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
  constexpr int NACC = 8;
  double d[NACC][4];
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) d[j][i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) d[j][i] = fma(a[i], b[i], d[j][i]);
      } else if (MODE == 1) {
        double dd[2] = {d[j][0], d[j][1]}; mma884(dd, a[0], b[0]); d[j][0] = dd[0]; d[j][1] = dd[1];
      } else if (MODE == 2) {
        double aa[2] = {a[0], a[1]}; mma1684(d[j], aa, b[0]);
      } else if (MODE == 3) {
        double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; mma1688(d[j], aa, bb);
      } else {
        mma16816(d[j], a, b);
      }
    }
  }
  double s = 0;
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += d[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run_tput(const char* name, double flops_per_warp_instr, int ctas_per_sm) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int grid = nsm * ctas_per_sm, iters = 20000;
  double* out; CK(cudaMalloc(&out, (size_t)grid * 256 * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  tput<MODE><<<grid, 256>>>(out, 1000, 1.0);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    tput<MODE><<<grid, 256>>>(out, iters, 1.0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double warps = (double)grid * 8;
  double fl = warps * iters * 8 * flops_per_warp_instr;
  printf("TPUT %-10s ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", name, ctas_per_sm, best, fl / best * 1e-9);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d clock=%d kHz smem/block optin=%zu L2=%d\n", p.name, p.major, p.minor,
         p.multiProcessorCount, p.clockRate, p.sharedMemPerBlockOptin, p.l2CacheSize);
  double hA[256], hB[128], hD[128], ref[128];
  srand(1);
  for (int i = 0; i < 256; i++) hA[i] = (rand() % 17) - 8;
  for (int i = 0; i < 128; i++) hB[i] = (rand() % 13) - 6;
  double *dA, *dB, *dD;
  CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, sizeof(hD)));
  CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  for (int K : {16, 8, 4, -4}) {
    int M = (K == -4) ? 8 : 16, KK = (K == -4) ? 4 : K;
    for (int i = 0; i < M; i++) for (int j = 0; j < 8; j++) {
      double s = 0; for (int k = 0; k < KK; k++) s += hA[i * 16 + k] * hB[k * 8 + j]; ref[i * 8 + j] = s; }
    CK(cudaMemset(dD, 0, sizeof(hD)));
    if (K == 16) layout_k16<<<1, 32>>>(dA, dB, dD);
    else if (K == 8) layout_k8<<<1, 32>>>(dA, dB, dD);
    else if (K == 4) layout_k4<<<1, 32>>>(dA, dB, dD);
    else layout_884<<<1, 32>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
    int bad = 0; for (int i = 0; i < M * 8; i++) if (hD[i] != ref[i]) bad++;
    printf("LAYOUT m%dn8k%d hypothesis H: %s (%d mismatches)\n", M, KK, bad ? "WRONG" : "OK", bad);
  }
  for (int c : {1, 2, 4}) {
    run_tput<0>("dfma", 4 * 32 * 2.0, c);
    run_tput<1>("m8n8k4", 8 * 8 * 4 * 2.0, c);
    run_tput<2>("m16n8k4", 16 * 8 * 4 * 2.0, c);
    run_tput<3>("m16n8k8", 16 * 8 * 8 * 2.0, c);
    run_tput<4>("m16n8k16", 16 * 8 * 16 * 2.0, c);
  }
  return 0;
}
