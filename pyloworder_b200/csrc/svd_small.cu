// SVD of the small n x n triangular factor on one GPU: one-sided Jacobi on the ROWS of R.
//
// Replaces `dsvd` = LAPACKE_dgesvd / dgesdd on R (pyLOM/vmmath/src/svd.c:83-139, called at
// svd.c:706).  Left rotations G <- Rot * G make the rows of G = J R mutually orthogonal, so
// G = diag(S) VT and R = J^T diag(S) VT, i.e. Ur = J^T.  Rows are contiguous in memory, the
// rotations of one round-robin round are independent (n/2 CTAs), and Jacobi delivers singular
// values to high *relative* accuracy, which the 1e-10 parity target on small sigma needs.
// Orthogonalising the rows of R is the Drmac-Veselic "apply Jacobi to R^T" preconditioned variant.
// Output ordering: S descending, VT rows = right singular vectors (the `V` every reference
// routine returns, POD/wrapper.py:80).
#include "pl_common.cuh"
#include "caqr.h"
#include <cmath>
#include <cstdlib>

namespace pl {

// G = P R, J = P with P the permutation that sorts the rows of R by decreasing norm (rank[r] = new position
// of row r): one-sided Jacobi converges in fewer sweeps on a norm-ordered (graded) matrix (de Rijk).
__global__ void jacobi_init_kernel(double* Gm, double* J, const double* R, int64_t ldr, int n, const int* rank) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  int r = (int)(idx / n), c = (int)(idx % n);
  const int64_t dst = (int64_t)rank[r] * n + c;
  Gm[dst] = R[(int64_t)r * ldr + c];
  J[dst] = (r == c) ? 1.0 : 0.0;
}

__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double (*sh)[4]) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = a; sh[1][w] = b; sh[2][w] = c; }
  __syncthreads();
  a = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
  b = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
  c = sh[2][0] + sh[2][1] + sh[2][2] + sh[2][3];
}

// One round of the round-robin ordering: CTA i rotates row pair (p, q).
__global__ void __launch_bounds__(128) jacobi_round_kernel(double* __restrict__ Gm, double* __restrict__ J, int n, int ne,
                                                           int round, double tol, int* __restrict__ rotations, const double* __restrict__ fl) {
  __shared__ double sh[3][4];
  const int i = blockIdx.x;
  int p, q;
  if (i == 0) { p = ne - 1; q = round; }
  else { p = (round + i) % (ne - 1); q = (round - i + (ne - 1)) % (ne - 1); }
  if (p >= n || q >= n) return;
  if (p > q) { int tmp = p; p = q; q = tmp; }
  double* gp = Gm + (int64_t)p * n; double* gq = Gm + (int64_t)q * n;
  double a = 0, b = 0, c = 0;
  for (int j = threadIdx.x; j < n; j += 128) { double x = gp[j], y = gq[j]; a += x * x; b += y * y; c += x * y; }
  block_sum3(a, b, c, sh);
  if (c == 0.0 || fabs(c) <= tol * sqrt(a) * sqrt(b) || fmin(a, b) <= fl[0]) return;
  if (threadIdx.x == 0) atomicAdd(rotations, 1);
  const double zeta = (b - a) / (2.0 * c);
  const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  double* jp = J + (int64_t)p * n; double* jq = J + (int64_t)q * n;
  for (int j = threadIdx.x; j < n; j += 128) {
    double x = gp[j], y = gq[j];
    gp[j] = cs * x - sn * y; gq[j] = sn * x + cs * y;
    double u = jp[j], v = jq[j];
    jp[j] = cs * u - sn * v; jq[j] = sn * u + cs * v;
  }
}

// One whole sweep (ne-1 rounds) in a single cooperative launch.  ONE WARP owns one row pair of every
// round: the two rows of G (and then of J) are held in registers (EPL elements per lane per row), the
// three inner products are shuffle reductions, so a rotation reads and writes each element once and
// needs no block barrier.  Rounds are separated by a grid-wide barrier (monotone counter in global
// memory); rows move between SMs from round to round, so G and J are read with ld.global.cg (L2) and
// never through a possibly stale L1 line.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
  }
  __syncthreads();
}

// Convergence state in device memory, so that the host never waits for a sweep (the call stays stream-ordered):
//   st[1] grid-barrier counter, st[2] converged flag, st[3] sweeps executed, st[4..6] rotation counters of three
//   consecutive sweeps (rotating).
// ALL sweeps run inside ONE cooperative launch (a launch per sweep had to win its SM slots back from the concurrently
// running Q formation up to 40 times, most of them for an empty "already converged" launch).  Sweep s counts its
// rotations in st[4 + s % 3]; every rotation is counted before the sweep's last grid barrier, so after it every CTA
// reads the same final count and all of them leave the loop together.  Block 0 clears the counter of sweep s + 1 at
// the start of sweep s: its last readers (end of sweep s - 2) have since arrived at a barrier of sweep s - 1, and its
// next writers first pass a barrier of sweep s, which block 0 reaches after the store (fence inside grid_barrier).
__device__ __forceinline__ bool sweep_done(const int* st) { return *reinterpret_cast<const volatile int*>(st + 2) != 0; }
__device__ __forceinline__ int* sweep_begin(int* st, int sweep) {
  if (blockIdx.x == 0 && threadIdx.x == 0) st[4 + (sweep + 1) % 3] = 0;
  return st + 4 + sweep % 3;
}
__device__ __forceinline__ bool sweep_end(int* st, const int* cnt) {     // call after the last grid barrier of the sweep
  const int nrot = *reinterpret_cast<const volatile int*>(cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st[3] += 1;
    if (nrot == 0) st[2] = 1;
  }
  return nrot == 0;
}
// Not converged within the budget: poison S so that the failure is loud wherever the result is read.
__global__ void jacobi_check_kernel(const int* st, double* S, int n) {
  if (st[2]) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) S[i] = __longlong_as_double(0x7ff8000000000000LL);
}

constexpr int JW = 8;   // warps (row pairs) per CTA

template <int EPL>
__global__ void __launch_bounds__(JW * 32) jacobi_sweep_kernel(double* __restrict__ Gm, double* __restrict__ J, int n, int ne,
                                                               double tol, int* __restrict__ state, unsigned* bar, const double* __restrict__ fl,
                                                               int max_sweeps) {
  if (sweep_done(state)) return;
  const double floor2 = fl[0];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * JW + (threadIdx.x >> 5);       // pair index inside a round
  const bool have = i < ne / 2;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
  int* rotations = sweep_begin(state, sweep);
  const unsigned bar_base = (unsigned)sweep * (unsigned)(ne - 1) * gridDim.x;
  for (int round = 0; round < ne - 1; round++) {
    int p = 0, q = 0;
    if (have) {
      if (i == 0) { p = ne - 1; q = round; }
      else { p = (round + i) % (ne - 1); q = (round - i + (ne - 1)) % (ne - 1); }
    }
    if (have && p < n && q < n) {
      if (p > q) { int tmp = p; p = q; q = tmp; }
      double* gp = Gm + (int64_t)p * n; double* gq = Gm + (int64_t)q * n;
      double* jp = J + (int64_t)p * n; double* jq = J + (int64_t)q * n;
      double x[EPL], y[EPL], u[EPL], v[EPL];
      double a = 0, b = 0, c = 0;
#pragma unroll
      for (int e = 0; e < EPL; e++) {       // all four rows are requested up front (one L2 round trip)
        const int j = lane + 32 * e;
        x[e] = (j < n) ? __ldcg(gp + j) : 0.0;
        y[e] = (j < n) ? __ldcg(gq + j) : 0.0;
        u[e] = (j < n) ? __ldcg(jp + j) : 0.0;
        v[e] = (j < n) ? __ldcg(jq + j) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < EPL; e++) { a = fma(x[e], x[e], a); b = fma(y[e], y[e], b); c = fma(x[e], y[e], c); }
      a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
      if (!(c == 0.0 || fabs(c) <= tol * sqrt(a) * sqrt(b) || fmin(a, b) <= floor2)) {
        if (lane == 0) atomicAdd(rotations, 1);
        const double zeta = (b - a) / (2.0 * c);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int j = lane + 32 * e;
          if (j < n) {
            gp[j] = cs * x[e] - sn * y[e]; gq[j] = sn * x[e] + cs * y[e];
            jp[j] = cs * u[e] - sn * v[e]; jq[j] = sn * u[e] + cs * v[e];
          }
        }
      }
    }
    grid_barrier(bar, bar_base + (unsigned)(round + 1) * gridDim.x);
  }
  if (sweep_end(state, rotations)) break;
  }
}

// Block one-sided Jacobi: a CTA owns a PAIR OF ROW BLOCKS (2*BR rows of G and of J, staged in shared
// memory) and performs all rotations between them with only block-level barriers; the grid-wide barrier
// (the expensive part, ~5 us) is needed once per block round, i.e. BR times less often than in the
// row-pair kernel above.  Round 0 of a sweep rotates every pair of the 2*BR local rows (this covers the
// pairs inside a block exactly once per sweep), later rounds rotate only the BR*BR cross pairs.
template <int BR, int EPL>
__global__ void __launch_bounds__(256) jacobi_block_sweep_kernel(double* __restrict__ Gm, double* __restrict__ J, int n, int nbe,
                                                                 double tol, int* __restrict__ state, unsigned* bar, const double* __restrict__ fl,
                                                                 int max_sweeps) {
  extern __shared__ __align__(16) double jsm[];
  if (sweep_done(state)) return;
  const double floor2 = fl[0];
  double* Gs = jsm;                       // [2*BR][n]
  double* Js = jsm + (size_t)2 * BR * n;  // [2*BR][n]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i = blockIdx.x;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
  int* rotations = sweep_begin(state, sweep);
  const unsigned bar_base = (unsigned)sweep * (unsigned)(nbe - 1) * gridDim.x;
  for (int round = 0; round < nbe - 1; round++) {
    int bp, bq;
    if (i == 0) { bp = nbe - 1; bq = round; }
    else { bp = (round + i) % (nbe - 1); bq = (round - i + (nbe - 1)) % (nbe - 1); }
    if (bp > bq) { int tmp = bp; bp = bq; bq = tmp; }
    // ---- stage the 2*BR rows (rows >= n are phantom: zeros, never stored)
    for (int lr = warp; lr < 2 * BR; lr += 8) {
      const int gr = (lr < BR ? bp * BR + lr : bq * BR + (lr - BR));
      const bool ok = gr < n;
      double gv[EPL], jv[EPL];
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const int j = lane + 32 * e;
        gv[e] = (ok && j < n) ? __ldcg(Gm + (int64_t)gr * n + j) : 0.0;
        jv[e] = (ok && j < n) ? __ldcg(J + (int64_t)gr * n + j) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const int j = lane + 32 * e;
        if (j < n) { Gs[(size_t)lr * n + j] = gv[e]; Js[(size_t)lr * n + j] = jv[e]; }
      }
    }
    __syncthreads();
    const int nmini = (round == 0) ? (2 * BR - 1) : BR;
    for (int k = 0; k < nmini; k++) {
      if (warp < BR) {
        int a, b;
        if (round == 0) {      // round-robin over the 2*BR local rows
          const int ne2 = 2 * BR;
          if (warp == 0) { a = ne2 - 1; b = k; }
          else { a = (k + warp) % (ne2 - 1); b = (k - warp + (ne2 - 1)) % (ne2 - 1); }
        } else { a = warp; b = BR + (warp + k) % BR; }
        double* x = Gs + (size_t)a * n; double* y = Gs + (size_t)b * n;
        double xv[EPL], yv[EPL];
        double sa = 0, sb = 0, sc = 0;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int j = lane + 32 * e;
          xv[e] = (j < n) ? x[j] : 0.0; yv[e] = (j < n) ? y[j] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < EPL; e++) { sa = fma(xv[e], xv[e], sa); sb = fma(yv[e], yv[e], sb); sc = fma(xv[e], yv[e], sc); }
        sa = warp_sum(sa); sb = warp_sum(sb); sc = warp_sum(sc);
        if (!(sc == 0.0 || fabs(sc) <= tol * sqrt(sa) * sqrt(sb) || fmin(sa, sb) <= floor2)) {
          if (lane == 0) atomicAdd(rotations, 1);
          const double zeta = (sb - sa) / (2.0 * sc);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          double* u = Js + (size_t)a * n; double* v = Js + (size_t)b * n;
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const int j = lane + 32 * e;
            if (j < n) { x[j] = cs * xv[e] - sn * yv[e]; y[j] = sn * xv[e] + cs * yv[e]; }
          }
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const int j = lane + 32 * e;
            xv[e] = (j < n) ? u[j] : 0.0; yv[e] = (j < n) ? v[j] : 0.0;
          }
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const int j = lane + 32 * e;
            if (j < n) { u[j] = cs * xv[e] - sn * yv[e]; v[j] = sn * xv[e] + cs * yv[e]; }
          }
        }
      }
      __syncthreads();
    }
    // ---- write the rows back
    for (int lr = warp; lr < 2 * BR; lr += 8) {
      const int gr = (lr < BR ? bp * BR + lr : bq * BR + (lr - BR));
      if (gr < n) {
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int j = lane + 32 * e;
          if (j < n) { Gm[(int64_t)gr * n + j] = Gs[(size_t)lr * n + j]; J[(int64_t)gr * n + j] = Js[(size_t)lr * n + j]; }
        }
      }
    }
    grid_barrier(bar, bar_base + (unsigned)(round + 1) * gridDim.x);
  }
  if (sweep_end(state, rotations)) break;
  }
}

// Block one-sided Jacobi, second generation: only G is staged in shared memory.  The left rotations of a block round are
// ACCUMULATED in a small orthogonal matrix Om (2 BR x 2 BR, shared memory; a Givens rotation costs 2 x 2 BR entries
// instead of 2 n) and applied to the 2 BR rows of J once per round as a dense product J_rows <- Om J_rows read from /
// written to L2.  Half the shared memory per row means twice the rows per block at a given n (BR = 16 up to n = 512,
// BR = 8 up to n = 1024): half the block rounds and grid barriers per sweep, half the CTAs on the machine (the sweep
// launches share the GPU with the Q formation), and the mini-steps touch G only.
// Rotation that orthogonalises two rows with squared norms a, b and inner product c != 0 (the smaller angle):
// t = sgn(d) h / (|d| + sqrt(d^2 + h^2)), d = b - a, h = 2 c;  cs = 1 / sqrt(1 + t^2), sn = cs t  -- the same numbers as
// zeta = d / h, t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)), without the two divisions and two square roots of the
// textbook form (each a ~150-cycle dependent step of every pair): rsqrt / rcp seeds + two Newton steps, like the
// Householder scalars of the panel kernels.
__device__ __forceinline__ void jacobi_rotation(double a, double b, double c, double& cs, double& sn) {
  const double d = b - a, h = 2.0 * c;
  const double q = fma(d, d, h * h);
  double y, rc;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
  const double hq = 0.5 * q;
  y = y * fma(-hq * y, y, 1.5);
  y = y * fma(-hq * y, y, 1.5);
  const double den = fabs(d) + q * y;                 // |d| + sqrt(d^2 + h^2) > 0
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));
  rc = rc * fma(-den, rc, 2.0);
  rc = rc * fma(-den, rc, 2.0);
  const double t = (signbit(d) ? -h : h) * rc;        // sgn(d) h / den   (the sign bit of +-0 counts, like copysign(1, zeta))
  const double w = fma(t, t, 1.0);
  double z;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(w));
  const double hw = 0.5 * w;
  z = z * fma(-hw * z, z, 1.5);
  z = z * fma(-hw * z, z, 1.5);
  cs = z; sn = z * t;
}

template <int BR, int EPL>
__global__ void __launch_bounds__(256) jacobi_block_sweep2_kernel(double* __restrict__ Gm, double* __restrict__ J, int n, int nbe,
                                                                  double tol, int* __restrict__ state, unsigned* bar, const double* __restrict__ fl,
                                                                  int max_sweeps) {
  extern __shared__ __align__(16) double jsm[];
  if (sweep_done(state)) return;
  constexpr int R2 = 2 * BR, LDO = R2 + 1;
  const double floor2 = fl[0], tol2 = tol * tol;
  double* Gs = jsm;                       // [R2][n]
  double* Om = jsm + (size_t)R2 * n;      // [R2][LDO]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i = blockIdx.x;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
  int* rotations = sweep_begin(state, sweep);
  const unsigned bar_base = (unsigned)sweep * (unsigned)(nbe - 1) * gridDim.x;
  for (int round = 0; round < nbe - 1; round++) {
    int bp, bq;
    if (i == 0) { bp = nbe - 1; bq = round; }
    else { bp = (round + i) % (nbe - 1); bq = (round - i + (nbe - 1)) % (nbe - 1); }
    if (bp > bq) { int tmp = bp; bp = bq; bq = tmp; }
    auto grow = [&](int lr) { return lr < BR ? bp * BR + lr : bq * BR + (lr - BR); };
    // ---- stage the 2*BR rows of G (rows >= n are phantom: zeros, never stored); Om <- I
    for (int lr = warp; lr < R2; lr += 8) {
      const int gr = grow(lr);
      const bool ok = gr < n;
      double gv[EPL];
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const int j = lane + 32 * e;
        gv[e] = (ok && j < n) ? __ldcg(Gm + (int64_t)gr * n + j) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const int j = lane + 32 * e;
        if (j < n) Gs[(size_t)lr * n + j] = gv[e];
      }
    }
    for (int e = tid; e < R2 * R2; e += 256) Om[(e / R2) * LDO + (e % R2)] = (e / R2 == e % R2) ? 1.0 : 0.0;
    __syncthreads();
    const int nmini = (round == 0) ? (R2 - 1) : BR;
    for (int k = 0; k < nmini; k++) {
      for (int pw = warp; pw < BR; pw += 8) {
        int a, b;
        if (round == 0) {      // round-robin over the 2*BR local rows
          if (pw == 0) { a = R2 - 1; b = k; }
          else { a = (k + pw) % (R2 - 1); b = (k - pw + (R2 - 1)) % (R2 - 1); }
        } else { a = pw; b = BR + (pw + k) % BR; }
        double* x = Gs + (size_t)a * n; double* y = Gs + (size_t)b * n;
        double xv[EPL], yv[EPL];
        double sa = 0, sb = 0, sc = 0;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int j = lane + 32 * e;
          xv[e] = (j < n) ? x[j] : 0.0; yv[e] = (j < n) ? y[j] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < EPL; e++) { sa = fma(xv[e], xv[e], sa); sb = fma(yv[e], yv[e], sb); sc = fma(xv[e], yv[e], sc); }
        sa = warp_sum(sa); sb = warp_sum(sb); sc = warp_sum(sc);
        if (!(sc == 0.0 || sc * sc <= tol2 * (sa * sb) || fmin(sa, sb) <= floor2)) {
          if (lane == 0) atomicAdd(rotations, 1);
          double cs, sn;
          jacobi_rotation(sa, sb, sc, cs, sn);
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const int j = lane + 32 * e;
            if (j < n) { x[j] = cs * xv[e] - sn * yv[e]; y[j] = sn * xv[e] + cs * yv[e]; }
          }
          if (lane < R2) {       // the same rotation on rows a, b of Om
            const double u = Om[a * LDO + lane], v = Om[b * LDO + lane];
            Om[a * LDO + lane] = cs * u - sn * v; Om[b * LDO + lane] = sn * u + cs * v;
          }
        }
      }
      __syncthreads();
    }
    // ---- write the rows of G back; J_rows <- Om J_rows (one column per thread and pass, rows of J from / to L2)
    for (int lr = warp; lr < R2; lr += 8) {
      const int gr = grow(lr);
      if (gr < n) {
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int j = lane + 32 * e;
          if (j < n) Gm[(int64_t)gr * n + j] = Gs[(size_t)lr * n + j];
        }
      }
    }
    for (int j = tid; j < n; j += 256) {
      double jo[R2];
#pragma unroll
      for (int sidx = 0; sidx < R2; sidx++) { const int gr = grow(sidx); jo[sidx] = gr < n ? __ldcg(J + (int64_t)gr * n + j) : 0.0; }
#pragma unroll 4
      for (int r = 0; r < R2; r++) {
        const int gr = grow(r);
        if (gr >= n) continue;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int sidx = 0; sidx < R2; sidx += 2) { acc0 = fma(Om[r * LDO + sidx], jo[sidx], acc0); acc1 = fma(Om[r * LDO + sidx + 1], jo[sidx + 1], acc1); }
        J[(int64_t)gr * n + j] = acc0 + acc1;
      }
    }
    grid_barrier(bar, bar_base + (unsigned)(round + 1) * gridDim.x);
  }
  if (sweep_end(state, rotations)) break;
  }
}

__global__ void __launch_bounds__(128) row_norm_kernel(double* s, const double* Gm, int n, int64_t ld) {
  __shared__ double sh[3][4];
  const double* g = Gm + (int64_t)blockIdx.x * ld;
  // scaled two-pass norm is unnecessary here: rows are O(sigma), far from over/underflow for POD data
  double a = 0, b = 0, c = 0;
  for (int j = threadIdx.x; j < n; j += 128) { double x = g[j]; a += x * x; }
  block_sum3(a, b, c, sh);
  if (threadIdx.x == 0) s[blockIdx.x] = sqrt(a);
}

// Noise floor of the singular values: rows of G whose norm falls below n eps max_k |row_k| are numerically zero (two such
// rows are never orthogonal RELATIVE to their own norms, so the relative criterion alone would rotate them forever --
// e.g. an all-zero snapshot column plus a duplicated one).  fl[0] = floor^2 (pairs with a row below it count as converged),
// fl[1] = floor (rows of V^T below it get an orthonormal completion, like exactly zero ones).
__global__ void __launch_bounds__(256) jacobi_floor_kernel(double* fl, const double* s, int n) {
  __shared__ double sh[256];
  double mx = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) mx = fmax(mx, s[k]);
  sh[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
  if (threadIdx.x == 0) { const double f = (double)n * 2.220446049250313e-16 * sh[0]; fl[1] = f; fl[0] = f * f; }
}

__global__ void rank_kernel(int* rank, const double* s, int n) {
  for (int k = threadIdx.x + blockIdx.x * blockDim.x; k < n; k += blockDim.x * gridDim.x) {
    const double sk = s[k];
    int r = 0;
    for (int j = 0; j < n; j++) { double sj = s[j]; r += (sj > sk) || (sj == sk && j < k); }
    rank[k] = r;
  }
}

// Row k of G is sigma_k * (right singular vector), row k of J is the left singular vector.  Ten million
// slightly non-orthogonal rotations (n ~ 1000, 20 sweeps) let the norms of the rows of J random-walk by
// ~1e-12, mostly on the modes with tiny sigma; the rows stay mutually orthogonal to ~1e-14, so they are
// renormalised here (the classical Jacobi-SVD clean-up, cf. the normalisation of the columns of U in xGESVJ).
__global__ void __launch_bounds__(128) svd_scatter_kernel(double* Ur, int64_t ldu, double* S, double* VT, int64_t ldvt,
                                                          const double* Gm, const double* J, const double* s,
                                                          const int* rank, int n) {
  __shared__ double sh[3][4];
  const int k = blockIdx.x, r = rank[k];
  const double sk = s[k], inv = sk > 0.0 ? 1.0 / sk : 0.0;
  double a = 0, b = 0, c = 0;
  for (int j = threadIdx.x; j < n; j += 128) { const double x = J[(int64_t)k * n + j]; a += x * x; }
  block_sum3(a, b, c, sh);
  const double jinv = a > 0.0 ? 1.0 / sqrt(a) : 0.0;
  if (threadIdx.x == 0) S[r] = sk;
  for (int j = threadIdx.x; j < n; j += 128) {
    VT[(int64_t)r * ldvt + j] = Gm[(int64_t)k * n + j] * inv;
    Ur[(int64_t)j * ldu + r] = J[(int64_t)k * n + j] * jinv;
  }
}

// Zero (or numerically zero: below the noise floor) singular values, e.g. an all-zero snapshot column or the null
// direction of centred data: LAPACK returns an orthonormal completion of V^T; do the same.  One CTA; for every such row
// try the unit vectors e_0, e_1, ..., orthogonalise against all rows already in place (classical Gram-Schmidt, all inner
// products of a pass computed at once, one warp per row) and keep the first candidate whose remainder is long enough.
// A remaining null space of dimension >= 1 holds some e_j with squared projection >= 1/n (the squared projections of
// e_0..e_{n-1} sum to the dimension), so the threshold 1/(2n) always finds one; three passes + a renormalisation in
// between bring the row to working-precision orthogonality whatever the cancellation of the first pass was.
__global__ void __launch_bounds__(1024) svd_complete_kernel(double* VT, int64_t ldvt, const double* S, int n, const double* __restrict__ fl) {
  extern __shared__ double dots[];      // n inner products
  __shared__ double sh[32];
  __shared__ double s_nn;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  auto norm2 = [&](const double* v) {
    double a = 0.0;
    for (int j = tid; j < n; j += 1024) a = fma(v[j], v[j], a);
    a = warp_sum(a);
    if (l == 0) sh[w] = a;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int i = 0; i < 32; i++) t += sh[i]; s_nn = t; }
    __syncthreads();
    return s_nn;
  };
  auto cgs_pass = [&](double* vz, int z) {
    for (int r = w; r < z; r += 32) {
      const double* vr = VT + (int64_t)r * ldvt;
      double d = 0.0;
      for (int j = l; j < n; j += 32) d = fma(vz[j], vr[j], d);
      d = warp_sum(d);
      if (l == 0) dots[r] = d;
    }
    __syncthreads();
    for (int j = tid; j < n; j += 1024) {
      double a0 = 0.0, a1 = 0.0;
      int r = 0;
      for (; r + 1 < z; r += 2) { a0 = fma(dots[r], VT[(int64_t)r * ldvt + j], a0); a1 = fma(dots[r + 1], VT[(int64_t)(r + 1) * ldvt + j], a1); }
      if (r < z) a0 = fma(dots[r], VT[(int64_t)r * ldvt + j], a0);
      vz[j] -= a0 + a1;
    }
    __syncthreads();
  };
  int cand = 0;
  const double thr = 0.5 / (double)n;
  for (int z = 0; z < n; z++) {
    if (S[z] > fl[1]) continue;          // rows are sorted: the (numerically) zero ones are at the end, earlier rows are complete
    double* vz = VT + (int64_t)z * ldvt;
    for (; cand < n; cand++) {
      for (int j = tid; j < n; j += 1024) vz[j] = (j == cand) ? 1.0 : 0.0;
      __syncthreads();
      cgs_pass(vz, z);
      double nn = norm2(vz);
      if (!(nn > thr)) continue;
      for (int pass = 0; pass < 2; pass++) {
        const double inv = 1.0 / sqrt(nn);
        for (int j = tid; j < n; j += 1024) vz[j] *= inv;
        __syncthreads();
        cgs_pass(vz, z);
        nn = norm2(vz);
      }
      const double inv = 1.0 / sqrt(nn);
      for (int j = tid; j < n; j += 1024) vz[j] *= inv;
      __syncthreads();
      cand++;
      break;
    }
  }
}

int64_t svd_small_scratch_doubles(int64_t n) { return 2 * n * n + 2 * n + 64; }

int svd_small(double* Ur, int64_t ldu, double* S, double* VT, int64_t ldvt, const double* R, int64_t ldr, int64_t n,
              double* scratch, int* sweeps_out, cudaStream_t st) {
  if (n <= 0) return 0;
  double* Gm = scratch;
  double* J = Gm + n * n;
  double* s = J + n * n;
  int* rank = reinterpret_cast<int*>(s + n);
  int* rot = rank + n;     // inside the 64-double tail
  const int ni = (int)n, ne = ni + (ni & 1);
  double* fl = scratch + 2 * n * n + 2 * n + 56;   // {floor^2, floor}: inside the 64-double tail, behind rank / state
  row_norm_kernel<<<ni, 128, 0, st>>>(s, R, ni, ldr);
  jacobi_floor_kernel<<<1, 256, 0, st>>>(fl, s, ni);
  if (getenv("PL_JACOBI_NOSORT")) PL_CUDA(cudaMemsetAsync(s, 0, (size_t)n * 8, st));     // equal keys -> rank = identity
  rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(rank, s, ni);
  jacobi_init_kernel<<<(unsigned)ceil_div(n * n, 256), 256, 0, st>>>(Gm, J, R, ldr, ni, rank);
  PL_LAUNCH_CHECK();
  count_launches(2);
  const double tol = 2.0 * std::sqrt((double)n) * 2.220446049250313e-16;
  int sweeps = 0;
  int* state = rot;                           // {-, barrier counter, converged, sweeps executed, 3 rotation counters, -}
  unsigned* bar = reinterpret_cast<unsigned*>(rot + 1);
  bool async_path = false;
  if (ni > 1) {
    // cooperative single-launch sweeps when all CTAs can be co-resident
    int dev = 0, coop = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int sweep_blocks = (ne / 2 + JW - 1) / JW;
    void* sweep_fn = ni <= 128 ? (void*)jacobi_sweep_kernel<4> : ni <= 256 ? (void*)jacobi_sweep_kernel<8>
                   : ni <= 512 ? (void*)jacobi_sweep_kernel<16> : (void*)jacobi_sweep_kernel<32>;
    const bool use_coop = coop && ni <= 1024 && sweep_blocks <= nsm;
    // block variant: BR rows per block, 2*BR*n*16 bytes of shared memory per CTA
    int br = 0;
    static const bool jac_old = getenv("PL_JACOBI_OLD") != nullptr;          // first-generation block kernel (A/B timing only)
    static const int br_force = getenv("PL_JACOBI_BR") ? atoi(getenv("PL_JACOBI_BR")) : 0;
    int nbe = 0; size_t bsm = 0; void* bfn = nullptr;
    if (coop && ni >= 64 && !getenv("PL_JACOBI_ROWPAIR") && !jac_old && ni <= 1024) {
      // second generation: G only in shared memory (2 BR n doubles + Om): BR = 16 up to n = 512, BR = 8 up to n = 1024
      br = 8;
      if (br_force == 8 || br_force == 16) br = br_force;
      if ((size_t)2 * br * ni * 8 > 200 * 1024) br = 8;
      const int nblk = (ni + br - 1) / br;
      nbe = nblk + (nblk & 1);
      bsm = (size_t)2 * br * ni * 8 + (size_t)2 * br * (2 * br + 1) * 8;
      if (br == 16) bfn = ni <= 128 ? (void*)jacobi_block_sweep2_kernel<16, 4> : ni <= 256 ? (void*)jacobi_block_sweep2_kernel<16, 8>
                        : (void*)jacobi_block_sweep2_kernel<16, 16>;
      else bfn = ni <= 128 ? (void*)jacobi_block_sweep2_kernel<8, 4> : ni <= 256 ? (void*)jacobi_block_sweep2_kernel<8, 8>
               : ni <= 512 ? (void*)jacobi_block_sweep2_kernel<8, 16> : (void*)jacobi_block_sweep2_kernel<8, 32>;
      if (nbe / 2 > nsm || nbe < 2) br = 0;
      else PL_CUDA(cudaFuncSetAttribute(bfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    } else if (coop && ni >= 64 && !getenv("PL_JACOBI_ROWPAIR")) {
      if ((size_t)2 * 8 * ni * 16 <= 200 * 1024) br = 8; else if (ni <= 1024 && (size_t)2 * 4 * ni * 16 <= 200 * 1024) br = 4;
      if (br) {
        const int nblk = (ni + br - 1) / br;
        nbe = nblk + (nblk & 1);
        bsm = (size_t)2 * br * ni * 16;
        if (br == 8) bfn = ni <= 128 ? (void*)jacobi_block_sweep_kernel<8, 4> : ni <= 256 ? (void*)jacobi_block_sweep_kernel<8, 8>
                         : ni <= 512 ? (void*)jacobi_block_sweep_kernel<8, 16> : (void*)jacobi_block_sweep_kernel<8, 25>;
        else bfn = (void*)jacobi_block_sweep_kernel<4, 32>;
        if (nbe / 2 > nsm) br = 0;
        else PL_CUDA(cudaFuncSetAttribute(bfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
      }
    }
    PL_CUDA(cudaMemsetAsync(state, 0, 8 * sizeof(int), st));
    if (br || use_coop) {
      // Stream-ordered: ONE cooperative launch runs sweeps until one ends without a rotation (at most `budget`);
      // convergence is tracked on the device (no host sync).  PL_JACOBI_MULTILAUNCH=1: one launch per sweep (A/B timing).
      async_path = true;
      static const int budget = getenv("PL_JACOBI_SWEEPS") ? atoi(getenv("PL_JACOBI_SWEEPS")) : 40;
      static const bool multi = getenv("PL_JACOBI_MULTILAUNCH") != nullptr;
      const int launches = multi ? budget : 1;
      int per_launch = multi ? 1 : budget;
      for (int k = 0; k < launches; k++) {
        if (k) { PL_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), st)); PL_CUDA(cudaMemsetAsync(state + 4, 0, 3 * sizeof(int), st)); }
        double tol_ = tol; int ni_ = ni, nbe_ = nbe, ne_ = ne;
        if (br) {
          void* args[] = {&Gm, &J, &ni_, &nbe_, &tol_, &state, &bar, &fl, &per_launch};
          PL_CUDA(cudaLaunchCooperativeKernel(bfn, dim3(nbe / 2), dim3(256), args, bsm, st));
        } else {
          void* args[] = {&Gm, &J, &ni_, &ne_, &tol_, &state, &bar, &fl, &per_launch};
          PL_CUDA(cudaLaunchCooperativeKernel(sweep_fn, dim3(sweep_blocks), dim3(JW * 32), args, 0, st));
        }
      }
      count_launches(launches);
    } else {
      // n > 1024 (or no cooperative launch): one launch per round and a host check per sweep.  This rare path
      // synchronises the stream.
      const int max_sweeps = 60;
      for (; sweeps < max_sweeps;) {
        PL_CUDA(cudaMemsetAsync(state, 0, sizeof(int), st));
        for (int r = 0; r < ne - 1; r++) {
          jacobi_round_kernel<<<ne / 2, 128, 0, st>>>(Gm, J, ni, ne, r, tol, rot, fl);
        }
        PL_LAUNCH_CHECK();
        count_launches(ne - 2);
        int h = 0;
        PL_CUDA(cudaMemcpyAsync(&h, rot, sizeof(int), cudaMemcpyDeviceToHost, st));
        PL_CUDA(cudaStreamSynchronize(st));
        sweeps++;
        if (h == 0) break;
      }
      if (sweeps >= max_sweeps) { set_error("svd_small: Jacobi did not converge in %d sweeps", max_sweeps); return 2; }
    }
  }
  row_norm_kernel<<<ni, 128, 0, st>>>(s, Gm, ni, n);
  rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(rank, s, ni);
  svd_scatter_kernel<<<ni, 128, 0, st>>>(Ur, ldu, S, VT, ldvt, Gm, J, s, rank, ni);
  PL_LAUNCH_CHECK();
  count_launches(2);
  {   // fill the rows of V^T that belong to exactly zero singular values
    svd_complete_kernel<<<1, 1024, (size_t)ni * sizeof(double), st>>>(VT, ldvt, S, ni, fl);
    PL_LAUNCH_CHECK();
  }
  if (async_path) {
    jacobi_check_kernel<<<1, 128, 0, st>>>(state, S, ni);
    PL_LAUNCH_CHECK();
  }
  if (sweeps_out || getenv("PL_DEBUG")) {     // debugging only: this synchronises
    if (async_path) {
      int h[4] = {0, 0, 0, 0};
      PL_CUDA(cudaMemcpyAsync(h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
      PL_CUDA(cudaStreamSynchronize(st));
      sweeps = h[3];
      if (!h[2]) { set_error("svd_small: Jacobi did not converge in %d sweeps", sweeps); return 2; }
    }
    if (sweeps_out) *sweeps_out = sweeps;
    if (getenv("PL_DEBUG")) fprintf(stderr, "[pl] svd_small n=%d sweeps=%d\n", ni, sweeps);
  }
  return 0;
}

}  // namespace pl
