// Transposed-tall FP64 GEMM on the tensor cores (DMMA, mma.sync m16n8k16.f64):
//     C[a x b] = X^T Y,   X (m x a), Y (m x b) row-major, m huge (the row-distributed dimension), a, b <= ~1024.
// This is the rank-local part of the reference's `matmulp` (pyLOM/vmmath/maths.py:93-110; dmatmulp,
// pyLOM/vmmath/src/vector_matrix.c:344-356: local cblas_dgemm + MPI_Allreduce) for the only shapes pyLOM uses it
// with -- `matmulp(Ai.T, Qi)` and `matmulp(Qi.T, Ai)` in randomized_qr (vmmath/svd.py:139,143) and
// `matmulp(U.T, Y)` in DMD -- i.e. a reduction over the rows of two tall operands.  The caller all-reduces C.
//
// The rows are the contraction index, so both operands are read exactly as they lie in memory: a stage holds KS = 16
// consecutive rows of a TA-column slice of X and a TB-column slice of Y; the A fragment of the MMA is the transposed
// read X[k][i] (conflict free with a row stride = 4 mod 16 doubles), the B fragment is Y[k][j].  Split-K over the
// rows: grid = (output tiles, row splits); every CTA writes its TA x TB partial sum to a scratch slab and a second
// kernel adds the slabs in a fixed order (deterministic, unlike atomics).  With a small (the randomized sketch, a = r)
// the kernel is HBM bound -- Y is streamed once, 8 bytes per 2a flop -- so the tile shapes for a <= 16 / <= 32 keep
// the padding waste out of the tensor pipe: (TA, TB) = (16, 128), (32, 128), (64, 64).
#include "pl_common.cuh"
#include "caqr.h"

namespace pl {

constexpr int TN_KS = 16, TN_ST = 4, TN_TARGET_CTAS = 592;

template <int TA, int TB>
struct TnSmem {
  double X[TN_ST][TN_KS][TA + 4];
  double Y[TN_ST][TN_KS][TB + 4];
};

// One operand slice: KS rows x W columns starting at (row k0, column c0); rows >= k_end and columns >= ncols read as 0.
template <int W, bool ALIGNED>
__device__ __forceinline__ void tn_load_slice(double (*dst)[W + 4], const double* __restrict__ src, int64_t ld, int64_t k0,
                                              int64_t k_end, int c0, int ncols, int tid) {
  if (ALIGNED) {
#pragma unroll
    for (int e = tid; e < TN_KS * (W / 2); e += 256) {
      const int r = e / (W / 2), c2 = (e % (W / 2)) * 2;
      const int64_t gr = k0 + r;
      const int gc = c0 + c2;
      int bytes = 0;
      if (gr < k_end && gc < ncols) bytes = (gc + 1 < ncols) ? 16 : 8;
      const double* p = src + (bytes ? gr * ld + gc : 0);
      unsigned s = (unsigned)__cvta_generic_to_shared(&dst[r][c2]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(p), "r"(bytes));
    }
  } else {
    for (int e = tid; e < TN_KS * W; e += 256) {
      const int r = e / W, c = e % W;
      const int64_t gr = k0 + r;
      const int gc = c0 + c;
      const bool ok = gr < k_end && gc < ncols;
      unsigned s = (unsigned)__cvta_generic_to_shared(&dst[r][c]);
      const double* p = src + (ok ? gr * ld + gc : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(p), "r"(ok ? 8 : 0));
    }
  }
}

template <int TA, int TB, int WA, bool ALIGNED>
__global__ void __launch_bounds__(256, 2)
gemm_tn_kernel(double* __restrict__ part, int a_pad, int b_pad, const double* __restrict__ X, int64_t ldx, int a,
               const double* __restrict__ Y, int64_t ldy, int b, int64_t m, int64_t rows_per_split, int tiles_b) {
  constexpr int WB = 8 / WA, MA = TA / (16 * WA), NN = TB / (8 * WB);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TnSmem<TA, TB>& S = *reinterpret_cast<TnSmem<TA, TB>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int wa = warp / WB, wb = warp % WB;
  const int a0 = (blockIdx.x / tiles_b) * TA, b0 = (blockIdx.x % tiles_b) * TB;
  const int64_t k_beg = (int64_t)blockIdx.y * rows_per_split;
  int64_t k_end = k_beg + rows_per_split; if (k_end > m) k_end = m;
  const int nk = k_end > k_beg ? (int)((k_end - k_beg + TN_KS - 1) / TN_KS) : 0;

  double acc[MA][NN][4];
#pragma unroll
  for (int i = 0; i < MA; i++)
#pragma unroll
    for (int j = 0; j < NN; j++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[i][j][c] = 0.0;

  for (int s = 0; s < TN_ST - 1; s++) {
    if (s < nk) {
      tn_load_slice<TA, ALIGNED>(S.X[s], X, ldx, k_beg + (int64_t)s * TN_KS, k_end, a0, a, tid);
      tn_load_slice<TB, ALIGNED>(S.Y[s], Y, ldy, k_beg + (int64_t)s * TN_KS, k_end, b0, b, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<TN_ST - 2>();
    __syncthreads();
    {
      const int nx = kt + TN_ST - 1;
      if (nx < nk) {
        tn_load_slice<TA, ALIGNED>(S.X[nx % TN_ST], X, ldx, k_beg + (int64_t)nx * TN_KS, k_end, a0, a, tid);
        tn_load_slice<TB, ALIGNED>(S.Y[nx % TN_ST], Y, ldy, k_beg + (int64_t)nx * TN_KS, k_end, b0, b, tid);
      }
      cp_async_commit();
    }
    const int st = kt % TN_ST;
    double fa[MA][8], fb[NN][4];
#pragma unroll
    for (int i = 0; i < MA; i++)
#pragma unroll
      for (int x = 0; x < 8; x++) fa[i][x] = S.X[st][t4 + 4 * (x >> 1)][wa * (16 * MA) + 16 * i + g + 8 * (x & 1)];
#pragma unroll
    for (int j = 0; j < NN; j++)
#pragma unroll
      for (int x = 0; x < 4; x++) fb[j][x] = S.Y[st][t4 + 4 * x][wb * (8 * NN) + 8 * j + g];
#pragma unroll
    for (int i = 0; i < MA; i++)
#pragma unroll
      for (int j = 0; j < NN; j++) mma16816(acc[i][j], fa[i], fb[j]);
  }
  cp_async_wait<0>();
  double* slab = part + (int64_t)blockIdx.y * a_pad * b_pad;
#pragma unroll
  for (int i = 0; i < MA; i++)
#pragma unroll
    for (int j = 0; j < NN; j++)
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const int r = a0 + wa * (16 * MA) + 16 * i + g + 8 * hh;
        const int c = b0 + wb * (8 * NN) + 8 * j + 2 * t4;
        *reinterpret_cast<double2*>(slab + (int64_t)r * b_pad + c) = make_double2(acc[i][j][2 * hh], acc[i][j][2 * hh + 1]);
      }
}

// C[i][j] (or C[j][i]) = sum over the split slabs, in slab order
__global__ void gemm_tn_reduce_kernel(double* __restrict__ C, int64_t ldc, int transpose_out, const double* __restrict__ part,
                                      int nsplit, int a_pad, int b_pad, int a, int b) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)a * b) return;
  const int i = (int)(idx / b), j = (int)(idx % b);
  const double* p = part + (int64_t)i * b_pad + j;
  const int64_t slab = (int64_t)a_pad * b_pad;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int s = 0;
  for (; s + 3 < nsplit; s += 4) { s0 += p[s * slab]; s1 += p[(s + 1) * slab]; s2 += p[(s + 2) * slab]; s3 += p[(s + 3) * slab]; }
  for (; s < nsplit; s++) s0 += p[s * slab];
  const double v = (s0 + s1) + (s2 + s3);
  if (transpose_out) C[(int64_t)j * ldc + i] = v; else C[(int64_t)i * ldc + j] = v;
}

struct TnCfg { int ta, tb, tiles_a, tiles_b, a_pad, b_pad, nsplit; int64_t rows_per_split; };
static TnCfg tn_config(int64_t a, int64_t b, int64_t m) {
  TnCfg c;
  if (a <= 16) { c.ta = 16; c.tb = 128; } else if (a <= 32) { c.ta = 32; c.tb = 128; } else { c.ta = 64; c.tb = 64; }
  c.tiles_a = (int)ceil_div(a, c.ta); c.tiles_b = (int)ceil_div(b, c.tb);
  c.a_pad = c.tiles_a * c.ta; c.b_pad = c.tiles_b * c.tb;
  const int64_t tiles = (int64_t)c.tiles_a * c.tiles_b;
  int64_t ns = ceil_div(TN_TARGET_CTAS, tiles);
  const int64_t max_by_rows = ceil_div(m, 1024);
  if (ns > max_by_rows) ns = max_by_rows;
  if (ns < 1) ns = 1;
  c.rows_per_split = round_up(ceil_div(m, ns), TN_KS);
  c.nsplit = (int)ceil_div(m, c.rows_per_split);
  if (c.nsplit < 1) c.nsplit = 1;
  return c;
}
size_t gemm_tn_workspace_bytes(int64_t a, int64_t b) {
  if (a > b) { int64_t t = a; a = b; b = t; }
  TnCfg c = tn_config(a, b, (int64_t)1 << 40);
  return (size_t)c.nsplit * c.a_pad * c.b_pad * 8 + 256;
}

template <int TA, int TB, int WA>
static int tn_launch(const TnCfg& c, double* part, const double* X, int64_t ldx, int a, const double* Y, int64_t ldy, int b,
                     int64_t m, cudaStream_t st) {
  const bool aligned = ((ldx & 1) == 0) && ((ldy & 1) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  const size_t smem = sizeof(TnSmem<TA, TB>);
  static DevOnce attr;
  if (first_on_device(attr)) {
    PL_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<TA, TB, WA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PL_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<TA, TB, WA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((unsigned)(c.tiles_a * c.tiles_b), (unsigned)c.nsplit);
  if (aligned)
    gemm_tn_kernel<TA, TB, WA, true><<<grid, 256, smem, st>>>(part, c.a_pad, c.b_pad, X, ldx, a, Y, ldy, b, m, c.rows_per_split, c.tiles_b);
  else
    gemm_tn_kernel<TA, TB, WA, false><<<grid, 256, smem, st>>>(part, c.a_pad, c.b_pad, X, ldx, a, Y, ldy, b, m, c.rows_per_split, c.tiles_b);
  PL_LAUNCH_CHECK();
  return 0;
}

// C (a x b, ldc) = X^T Y; `ws` >= gemm_tn_workspace_bytes(a, b)
int gemm_tn(double* C, int64_t ldc, const double* X, int64_t ldx, int64_t a, const double* Y, int64_t ldy, int64_t b, int64_t m,
            double* ws, cudaStream_t st) {
  if (a <= 0 || b <= 0) return 0;
  if (m <= 0) {   // empty contraction: C = 0
    for (int64_t i = 0; i < a; i++) PL_CUDA(cudaMemsetAsync(C + i * ldc, 0, (size_t)b * 8, st));
    return 0;
  }
  int transpose_out = 0;
  if (a > b) {   // the narrow operand goes on the MMA's m side (tile shapes are chosen on it): C^T = Y^T X
    const double* tp = X; X = Y; Y = tp;
    int64_t t = ldx; ldx = ldy; ldy = t;
    t = a; a = b; b = t;
    transpose_out = 1;
  }
  if (a > 65535 || b > 65535) { set_error("gemm_tn: output too large"); return -5; }
  const TnCfg c = tn_config(a, b, m);
  int rc;
  ProfScope ps(PROF_GEMM, st);
  if (c.ta == 16) rc = tn_launch<16, 128, 1>(c, ws, X, ldx, (int)a, Y, ldy, (int)b, m, st);
  else if (c.ta == 32) rc = tn_launch<32, 128, 1>(c, ws, X, ldx, (int)a, Y, ldy, (int)b, m, st);
  else rc = tn_launch<64, 64, 2>(c, ws, X, ldx, (int)a, Y, ldy, (int)b, m, st);
  if (rc) return rc;
  gemm_tn_reduce_kernel<<<(unsigned)ceil_div(a * b, 256), 256, 0, st>>>(C, ldc, transpose_out, ws, c.nsplit, c.a_pad, c.b_pad, (int)a, (int)b);
  PL_LAUNCH_CHECK();
  return 0;
}

}  // namespace pl
