// Tall-times-small FP64 GEMM on the tensor cores (DMMA, mma.sync m16n8k16.f64):
//     C[m x n] = A[m x k] * B[k x n],   m huge, k,n <= ~1024.
// Replaces `dmatmul` -> cblas_dgemm (pyLOM/vmmath/src/vector_matrix.c:206-221,234-242) for the two
// calls per tsqr_svd (svd.c:673,708) and the one per POD.reconstruct (POD/wrapper.pyx:324-351).
//
// CTA tile 128 x 64, BK = 16, 8 warps as 4(M) x 2(N) with 32 x 32 warp tiles (2 m16 x 4 n8 MMAs per
// k16 step), 3-stage cp.async pipeline (128 x 32 / 128 x 16 tiles for narrow outputs).  B is tiny and is first packed (zero padded to kp % 16 == 0,
// np % 64 == 0) so that every B access is aligned and in range; A is read in place with row / k
// predication (16-byte cp.async when lda is even and the base is aligned, 8-byte loads otherwise).
// The CTAs that share a row tile (blockIdx.x = column tile, fastest) re-read A from L2, so DRAM
// sees A once.
#include "pl_common.cuh"
#include "caqr.h"

namespace pl {

constexpr int BM = 128, BK = 16, STAGES = 3;
constexpr int AS = BK + 4;    // 20: A-fragment reads (row g, col t) conflict free

// Column-tile width BN = 64 (8 warps as 4 x 2, 32 x 32 warp tiles) for the n x n back-multiply; BN = 32 / 16
// (8 warps as 8 x 1, 16 x BN warp tiles) for narrow outputs -- the sketch products A*Omega, A*Q2 of randomized_qr
// (n = r <= 32) are HBM bound and a 64-wide tile would spend the tensor pipe on zero padding.
template <int BN>
struct GemmSmem {
  double A[STAGES][BM][AS];
  double B[STAGES][BK][BN + 4];   // BN + 4 = 4 mod 16: B-fragment reads (row t, col g) conflict free
};

template <bool ALIGNED, int BN>
__device__ __forceinline__ void gemm_load_stage(GemmSmem<BN>& S, int stage, const double* __restrict__ A, int64_t lda,
                                                const double* __restrict__ Bp, int64_t ldb, int64_t row_base,
                                                int64_t m, int64_t k, int k0, int col_base, int tid) {
  // A tile: 128 rows x 16 cols
  if (ALIGNED) {
    for (int e = tid; e < BM * (BK / 2); e += 256) {
      const int r = e >> 3, c2 = (e & 7) * 2;
      const int64_t gr = row_base + r;
      const bool ok = (gr < m) && (k0 + c2 < k);      // k even in the aligned path => pair is all-in or all-out
      cp_async16(&S.A[stage][r][c2], A + (ok ? gr : 0) * lda + (ok ? k0 + c2 : 0), ok);
    }
  } else {
    for (int e = tid; e < BM * BK; e += 256) {
      const int r = e >> 4, c = e & 15;
      const int64_t gr = row_base + r;
      double v = 0.0;
      if (gr < m && k0 + c < k) v = A[gr * lda + k0 + c];
      S.A[stage][r][c] = v;
    }
  }
  // B tile: 16 rows x BN cols (always aligned / in range thanks to the packing)
  for (int e = tid; e < BK * (BN / 2); e += 256) {
    const int r = e / (BN / 2), c2 = (e % (BN / 2)) * 2;
    cp_async16(&S.B[stage][r][c2], Bp + (int64_t)(k0 + r) * ldb + col_base + c2, true);
  }
}

template <bool ALIGNED, int BN>
__global__ void __launch_bounds__(256, 2)
gemm_tall_kernel(double* __restrict__ C, int64_t ldc, const double* __restrict__ A, int64_t lda,
                 const double* __restrict__ Bp, int64_t ldb, int64_t m, int n, int k, int kp) {
  constexpr int WN = (BN == 64) ? 2 : 1, WM = 8 / WN;          // warp grid
  constexpr int M2 = BM / (16 * WM), N2 = BN / (8 * WN);       // m16 / n8 blocks per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmSmem<BN>& S = *reinterpret_cast<GemmSmem<BN>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int col_base = blockIdx.x * BN;
  const int64_t row_base = (int64_t)blockIdx.y * BM;
  const int nk = kp / BK;

  double acc[M2][N2][4];
#pragma unroll
  for (int a = 0; a < M2; a++)
#pragma unroll
    for (int b = 0; b < N2; b++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[a][b][c] = 0.0;

  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) gemm_load_stage<ALIGNED, BN>(S, s, A, lda, Bp, ldb, row_base, m, k, s * BK, col_base, tid);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = kt + STAGES - 1;
      if (nx < nk) gemm_load_stage<ALIGNED, BN>(S, nx % STAGES, A, lda, Bp, ldb, row_base, m, k, nx * BK, col_base, tid);
      cp_async_commit();
    }
    const int st = kt % STAGES;
    double fa[M2][8], fb[N2][4];
#pragma unroll
    for (int m2 = 0; m2 < M2; m2++)
#pragma unroll
      for (int x = 0; x < 8; x++) fa[m2][x] = S.A[st][(16 * M2) * wm + 16 * m2 + g + 8 * (x & 1)][t4 + 4 * (x >> 1)];
#pragma unroll
    for (int n2 = 0; n2 < N2; n2++)
#pragma unroll
      for (int x = 0; x < 4; x++) fb[n2][x] = S.B[st][t4 + 4 * x][(8 * N2) * wn + 8 * n2 + g];
#pragma unroll
    for (int m2 = 0; m2 < M2; m2++)
#pragma unroll
      for (int n2 = 0; n2 < N2; n2++) mma16816(acc[m2][n2], fa[m2], fb[n2]);
  }
  cp_async_wait<0>();
  // ---- epilogue: predicated stores (16-byte when the destination pair is aligned)
  const bool vec = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
  for (int m2 = 0; m2 < M2; m2++)
#pragma unroll
    for (int n2 = 0; n2 < N2; n2++)
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const int64_t gr = row_base + (16 * M2) * wm + 16 * m2 + g + 8 * hh;
        const int gc = col_base + (8 * N2) * wn + 8 * n2 + 2 * t4;
        if (gr < m) {
          double* dst = C + gr * ldc + gc;
          if (vec && gc + 1 < n) {
            *reinterpret_cast<double2*>(dst) = make_double2(acc[m2][n2][2 * hh], acc[m2][n2][2 * hh + 1]);
          } else {
            if (gc < n) dst[0] = acc[m2][n2][2 * hh];
            if (gc + 1 < n) dst[1] = acc[m2][n2][2 * hh + 1];
          }
        }
      }
}

template <int BN>
static int gemm_tall_launch(double* C, int64_t ldc, const double* A, int64_t lda, const double* Bp, int64_t ldb, int64_t m,
                            int64_t n, int64_t k, cudaStream_t st) {
  const int64_t kp = round_up(k, BK), np = round_up(n, BN);
  static DevOnce attr;
  if (first_on_device(attr)) {
    PL_CUDA(cudaFuncSetAttribute(gemm_tall_kernel<true, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem<BN>)));
    PL_CUDA(cudaFuncSetAttribute(gemm_tall_kernel<false, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem<BN>)));
  }
  const bool aligned = ((lda & 1) == 0) && ((k & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  // rows on grid.y (<= 65535 tiles per launch)
  int64_t tiles = ceil_div(m, BM), done = 0;
  while (done < tiles) {
    int64_t ny = tiles - done; if (ny > 65535) ny = 65535;
    dim3 grid((unsigned)(np / BN), (unsigned)ny);
    const int64_t r0 = done * BM;
    if (aligned)
      gemm_tall_kernel<true, BN><<<grid, 256, sizeof(GemmSmem<BN>), st>>>(C + r0 * ldc, ldc, A + r0 * lda, lda, Bp, ldb, m - r0, (int)n, (int)k, (int)kp);
    else
      gemm_tall_kernel<false, BN><<<grid, 256, sizeof(GemmSmem<BN>), st>>>(C + r0 * ldc, ldc, A + r0 * lda, lda, Bp, ldb, m - r0, (int)n, (int)k, (int)kp);
    PL_LAUNCH_CHECK();
    done += ny;
  }
  return 0;
}

// Bp: packed B, at least round_up(k, 16) rows and ldb >= round_up(n, 64) zero-padded columns
int gemm_tall(double* C, int64_t ldc, const double* A, int64_t lda, const double* Bp, int64_t ldb, int64_t m,
              int64_t n, int64_t k, cudaStream_t st) {
  if (m <= 0 || n <= 0) return 0;
  if (ldb < round_up(n, 64)) { set_error("gemm_tall: packed B too narrow"); return -6; }
  if (n <= 16) return gemm_tall_launch<16>(C, ldc, A, lda, Bp, ldb, m, n, k, st);
  if (n <= 32) return gemm_tall_launch<32>(C, ldc, A, lda, Bp, ldb, m, n, k, st);
  return gemm_tall_launch<64>(C, ldc, A, lda, Bp, ldb, m, n, k, st);
}

// dst (rows_p x cols_p, zero padded) <- diag(rowscale) * src (rows x cols, lds)
__global__ void pad_small_kernel(double* dst, int64_t rows_p, int64_t cols_p, const double* src, int64_t lds, int64_t rows,
                                 int64_t cols, const double* rowscale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_p * cols_p) return;
  int64_t r = idx / cols_p, c = idx - r * cols_p;
  double v = 0.0;
  if (r < rows && c < cols) { v = src[r * lds + c]; if (rowscale) v *= rowscale[r]; }
  dst[idx] = v;
}
int pad_small(double* dst, int64_t rows_p, int64_t cols_p, const double* src, int64_t lds, int64_t rows, int64_t cols,
              const double* rowscale, cudaStream_t st) {
  int64_t tot = rows_p * cols_p;
  if (tot == 0) return 0;
  pad_small_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(dst, rows_p, cols_p, src, lds, rows, cols, rowscale);
  PL_LAUNCH_CHECK();
  return 0;
}

}  // namespace pl
