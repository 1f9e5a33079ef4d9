// Streaming row kernels: temporal_mean / subtract_mean / fused centering (+ optional padding into
// the factorisation buffer), vecmat, RMSE partial sums.  All HBM-bound: one read (+ one write) of
// the snapshot matrix, 8/16-byte coalesced accesses, a group of lanes per row, shuffle reductions.
//
// Reference semantics: pyLOM/vmmath/src/averaging.c:29-46 (dtemporal_mean: row mean over the n
// snapshots), :109-124 (dsubtract_mean), pyLOM/vmmath/src/vector_matrix.c:401-414 (dvecmat),
// pyLOM/vmmath/src/stats.c:44-72 (dRMSE_relative sums).
#include "pl_common.cuh"

namespace pl {

enum RowMode { ROW_MEAN = 0, ROW_SUB = 1, ROW_CENTER = 2, ROW_COPY = 3, ROW_VAR = 4, ROW_NORMVAR = 5, ROW_CENTERVAR = 6 };

// One group of GS lanes per row (GS = 32 for n >= 32, smaller powers of two for short rows).
// The row is read twice in ROW_CENTER; the second read hits L1/L2 (a row is <= 8 KiB).
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) row_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src,
                                                  int64_t lds, double* __restrict__ mean_out,
                                                  const double* __restrict__ mean_in, double* __restrict__ var_out,
                                                  const double* __restrict__ var_in, int64_t m, int n, int gs,
                                                  int pad_to) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % gs;                 // lane inside the row group
  const int grp = lane / gs;                 // row group inside the warp
  const int gpw = 32 / gs;                   // row groups per warp
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const double inv_n = 1.0 / (double)n;
  for (int64_t row0 = warp * gpw; row0 < m; row0 += nwarps * gpw) {
    const int64_t row = row0 + grp;
    const bool live = row < m;
    const double* x = src + (live ? row : 0) * lds;
    double mu = 0.0;
    double scl = 1.0;   // divisor applied after centering (the row variance in the variance-normalising modes)
    if (MODE == ROW_MEAN || MODE == ROW_CENTER || MODE == ROW_CENTERVAR) {
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      if (live) {
        if (VEC) {
          const double2* xv = reinterpret_cast<const double2*>(x);
          const int nv = n >> 1;
          int j = sub;
          for (; j + 3 * gs < nv; j += 4 * gs) {
            double2 a = xv[j], b = xv[j + gs], c = xv[j + 2 * gs], d = xv[j + 3 * gs];
            s0 += a.x + a.y; s1 += b.x + b.y; s2 += c.x + c.y; s3 += d.x + d.y;
          }
          for (; j < nv; j += gs) { double2 a = xv[j]; s0 += a.x + a.y; }
        } else {
          int j = sub;
          for (; j + 3 * gs < n; j += 4 * gs) {
            s0 += x[j]; s1 += x[j + gs]; s2 += x[j + 2 * gs]; s3 += x[j + 3 * gs];
          }
          for (; j < n; j += gs) s0 += x[j];
        }
      }
      double s = (s0 + s1) + (s2 + s3);
      for (int o = gs >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      mu = s * inv_n;
      if (live && sub == 0 && mean_out) mean_out[row] = mu;
    } else if (MODE == ROW_SUB || MODE == ROW_VAR || MODE == ROW_NORMVAR) {
      if (live) mu = mean_in[row];
    }
    if (MODE == ROW_VAR || MODE == ROW_CENTERVAR) {   // population variance (1/n), src/averaging.c:70-90
      double v0 = 0, v1 = 0;
      if (live) {
        int j = sub;
        for (; j + gs < n; j += 2 * gs) { double a = x[j] - mu, b = x[j + gs] - mu; v0 += a * a; v1 += b * b; }
        for (; j < n; j += gs) { double a = x[j] - mu; v0 += a * a; }
      }
      double v = v0 + v1;
      for (int o = gs >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      v *= inv_n;
      if (live && sub == 0 && var_out) var_out[row] = v;
      scl = v;
    } else if (MODE == ROW_NORMVAR) {
      if (live) scl = var_in[row];
    }
    if (MODE != ROW_MEAN && MODE != ROW_VAR && live) {
      double* y = dst + row * ldd;
      if (VEC) {
        const double2* xv = reinterpret_cast<const double2*>(x);
        double2* yv = reinterpret_cast<double2*>(y);
        const int nv = n >> 1;
        for (int j = sub; j < nv; j += gs) {
          double2 a = xv[j];
          a.x -= mu; a.y -= mu;
          if (MODE == ROW_NORMVAR || MODE == ROW_CENTERVAR) { a.x /= scl; a.y /= scl; }
          yv[j] = a;
        }
        for (int j = nv + sub; j < (pad_to >> 1); j += gs) yv[j] = make_double2(0.0, 0.0);
      } else {
        for (int j = sub; j < n; j += gs) y[j] = (MODE == ROW_NORMVAR || MODE == ROW_CENTERVAR) ? (x[j] - mu) / scl : (x[j] - mu);
        for (int j = n + sub; j < pad_to; j += gs) y[j] = 0.0;
      }
    }
  }
}

static int pick_gs(int64_t n, bool vec) {
  int64_t e = vec ? (n >> 1) : n;
  int gs = 1;
  while (gs < 32 && gs < e) gs <<= 1;
  return gs;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE>
static int launch_row(double* dst, int64_t ldd, const double* src, int64_t lds, double* mean_out,
                      const double* mean_in, int64_t m, int64_t n, int64_t pad_to, cudaStream_t st,
                      double* var_out = nullptr, const double* var_in = nullptr) {
  if (m <= 0 || n <= 0) return 0;
  bool vec = (n % 2 == 0) && (lds % 2 == 0) && aligned16(src) &&
             (MODE == ROW_MEAN || MODE == ROW_VAR || ((ldd % 2 == 0) && aligned16(dst) && (pad_to % 2 == 0)));
  int gs = pick_gs(n, vec);
  int64_t rows_per_block = 8 * (32 / gs);
  int64_t blocks = ceil_div(m, rows_per_block);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (vec)
    row_kernel<MODE, true><<<(unsigned)blocks, 256, 0, st>>>(dst, ldd, src, lds, mean_out, mean_in, var_out, var_in, m, (int)n, gs, (int)pad_to);
  else
    row_kernel<MODE, false><<<(unsigned)blocks, 256, 0, st>>>(dst, ldd, src, lds, mean_out, mean_in, var_out, var_in, m, (int)n, gs, (int)pad_to);
  PL_LAUNCH_CHECK();
  return 0;
}

int temporal_mean(double* out, const double* X, int64_t m, int64_t n, cudaStream_t st) {
  return launch_row<ROW_MEAN>(nullptr, 0, X, n, out, nullptr, m, n, n, st);
}
int subtract_mean(double* out, int64_t ldo, const double* X, const double* mean, int64_t m, int64_t n, int64_t pad_to,
                  cudaStream_t st) {
  return launch_row<ROW_SUB>(out, ldo, X, n, nullptr, mean, m, n, pad_to, st);
}
int center_rows(double* Y, int64_t ldy, double* mean, const double* X, int64_t m, int64_t n, int64_t pad_to,
                cudaStream_t st) {
  return launch_row<ROW_CENTER>(Y, ldy, X, n, mean, nullptr, m, n, pad_to, st);
}
int temporal_variance(double* out, const double* X, const double* mean, int64_t m, int64_t n, cudaStream_t st) {
  return launch_row<ROW_VAR>(nullptr, 0, X, n, nullptr, mean, m, n, n, st, out, nullptr);
}
int norm_variance(double* out, int64_t ldo, const double* X, const double* mean, const double* var, int64_t m, int64_t n,
                  int64_t pad_to, cudaStream_t st) {
  return launch_row<ROW_NORMVAR>(out, ldo, X, n, nullptr, mean, m, n, pad_to, st, nullptr, var);
}
int center_var_rows(double* Y, int64_t ldy, double* mean, double* var, const double* X, int64_t m, int64_t n, int64_t pad_to,
                    cudaStream_t st) {
  return launch_row<ROW_CENTERVAR>(Y, ldy, X, n, mean, nullptr, m, n, pad_to, st, var, nullptr);
}
int copy_pad(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t m, int64_t n, int64_t pad_to,
             cudaStream_t st) {
  return launch_row<ROW_COPY>(dst, ldd, src, lds, nullptr, nullptr, m, n, pad_to, st);
}

// ---- fp32 <-> fp64 streaming conversion (the fp32 entry points of the Python layer compute in fp64) ----------------
__global__ void __launch_bounds__(256) widen_kernel(double* __restrict__ dst, const float* __restrict__ src, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t n4 = count >> 2;
    for (int64_t k = i; k < n4; k += stride) {
      const float4 v = reinterpret_cast<const float4*>(src)[k];
      reinterpret_cast<double2*>(dst)[2 * k] = make_double2((double)v.x, (double)v.y);
      reinterpret_cast<double2*>(dst)[2 * k + 1] = make_double2((double)v.z, (double)v.w);
    }
    for (int64_t k = 4 * n4 + i; k < count; k += stride) dst[k] = (double)src[k];
  } else {
    for (int64_t k = i; k < count; k += stride) dst[k] = (double)src[k];
  }
}
__global__ void __launch_bounds__(256) narrow_kernel(float* __restrict__ dst, const double* __restrict__ src, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t n4 = count >> 2;
    for (int64_t k = i; k < n4; k += stride) {
      const double2 a = reinterpret_cast<const double2*>(src)[2 * k], b = reinterpret_cast<const double2*>(src)[2 * k + 1];
      reinterpret_cast<float4*>(dst)[k] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
    }
    for (int64_t k = 4 * n4 + i; k < count; k += stride) dst[k] = (float)src[k];
  } else {
    for (int64_t k = i; k < count; k += stride) dst[k] = (float)src[k];
  }
}
int widen_f32(double* dst, const float* src, int64_t count, cudaStream_t st) {
  if (count <= 0) return 0;
  int64_t blocks = ceil_div(count, 256 * 8); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  widen_kernel<<<(unsigned)blocks, 256, 0, st>>>(dst, src, count);
  PL_LAUNCH_CHECK();
  return 0;
}
int narrow_f64(float* dst, const double* src, int64_t count, cudaStream_t st) {
  if (count <= 0) return 0;
  int64_t blocks = ceil_div(count, 256 * 8); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  narrow_kernel<<<(unsigned)blocks, 256, 0, st>>>(dst, src, count);
  PL_LAUNCH_CHECK();
  return 0;
}

// ---- complex <-> real embedding (complex128 tsqr_svd runs as a real SVD of [[Ar, -Ai], [Ai, Ar]]) -------------------
// Ahat (2m x 2n, row major) from A (m x n complex, interleaved re/im)
__global__ void __launch_bounds__(256) complex_embed_kernel(double* __restrict__ Ah, const double2* __restrict__ A, int64_t m, int64_t n) {
  const int64_t tot = m * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / n, j = idx - i * n;
    const double2 a = A[idx];
    Ah[i * 2 * n + j] = a.x;           Ah[i * 2 * n + n + j] = -a.y;
    Ah[(m + i) * 2 * n + j] = a.y;     Ah[(m + i) * 2 * n + n + j] = a.x;
  }
}
// Uc (m x n complex) = (P_top - Q_bot) + i (Q_top + P_bot); P, Q are 2m x n real, Q may be null (zero)
__global__ void __launch_bounds__(256) complex_pack_kernel(double2* __restrict__ Uc, const double* __restrict__ P, const double* __restrict__ Q,
                                                           int64_t m, int64_t n) {
  const int64_t tot = m * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    double re = P[idx], im = P[tot + idx];
    if (Q) { re -= Q[tot + idx]; im += Q[idx]; }
    Uc[idx] = make_double2(re, im);
  }
}
int complex_embed(double* Ah, const double* A, int64_t m, int64_t n, cudaStream_t st) {
  if (m * n <= 0) return 0;
  int64_t blocks = ceil_div(m * n, 256 * 4); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  complex_embed_kernel<<<(unsigned)blocks, 256, 0, st>>>(Ah, reinterpret_cast<const double2*>(A), m, n);
  PL_LAUNCH_CHECK();
  return 0;
}
int complex_pack(double* Uc, const double* P, const double* Q, int64_t m, int64_t n, cudaStream_t st) {
  if (m * n <= 0) return 0;
  int64_t blocks = ceil_div(m * n, 256 * 4); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
  complex_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<double2*>(Uc), P, Q, m, n);
  PL_LAUNCH_CHECK();
  return 0;
}

// ---- vecmat: C[i,:] = v[i] * A[i,:] --------------------------------------------------------
__global__ void vecmat_kernel(double* C, int64_t ldc, const double* v, const double* A, int64_t lda, int64_t m, int64_t n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = m * n;
  for (; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = idx / n, j = idx - i * n;
    C[i * ldc + j] = v[i] * A[i * lda + j];
  }
}
int vecmat(double* C, int64_t ldc, const double* v, const double* A, int64_t lda, int64_t m, int64_t n, cudaStream_t st) {
  if (m * n == 0) return 0;
  int64_t blocks = ceil_div(m * n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  vecmat_kernel<<<(unsigned)blocks, 256, 0, st>>>(C, ldc, v, A, lda, m, n);
  PL_LAUNCH_CHECK();
  return 0;
}

// ---- RMSE sums: out[0] = sum (A-B)^2, out[1] = sum A^2 (deterministic two-stage reduction) ----
constexpr int RM_BLOCKS = 148 * 4;
__global__ void __launch_bounds__(256) sumsq_partial(double* part, const double* A, const double* B, int64_t cnt) {
  double s1 = 0, s2 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
    double a = A[i], d = a - B[i];
    s1 += d * d; s2 += a * a;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  __shared__ double sh[2][8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; i++) { a += sh[0][i]; b += sh[1][i]; }
    part[blockIdx.x] = a; part[RM_BLOCKS + blockIdx.x] = b;
  }
}
__global__ void sumsq_final(double* out, const double* part) {
  double s1 = 0, s2 = 0;
  for (int i = threadIdx.x; i < RM_BLOCKS; i += 32) { s1 += part[i]; s2 += part[RM_BLOCKS + i]; }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (threadIdx.x == 0) { out[0] = s1; out[1] = s2; }
}
int sumsq_diff(double* out2, double* scratch, const double* A, const double* B, int64_t cnt, cudaStream_t st) {
  sumsq_partial<<<RM_BLOCKS, 256, 0, st>>>(scratch, A, B, cnt);
  PL_LAUNCH_CHECK();
  sumsq_final<<<1, 32, 0, st>>>(out2, scratch);
  PL_LAUNCH_CHECK();
  return 0;
}

}  // namespace pl
