// Communication-avoiding Householder QR of a tall-skinny fp64 matrix on one GPU (sm_100a).
//
// Replaces the reference's per-rank `dqr` = LAPACKE_dgeqrf + LAPACKE_dorgqr
// (pyLOM/vmmath/src/svd.c:280-321, called from dtsqr at svd.c:594).  Same mathematics
// (Householder reflectors, backward stable, explicit thin Q), different organisation:
//
//   * columns are processed in panels of NB = 32;
//   * rows are cut into tiles of TB = 128 rows (G = 4 blocks of NB rows); a CTA owns a STRIP of
//     up to SMAX consecutive tiles and reduces them with a flat tree (the NB x NB triangle R and
//     the NB carried rows Z stay in shared memory while the strip streams through);
//   * the strips' NB-row tops are reduced by the same two kernels applied recursively
//     (levels), so the whole factorisation is 2 kernel launches per (panel, level);
//   * panel kernel  (caqr_panel_kernel):  Householder on [R; tile] with one row per thread,
//     one batched 32-value warp transpose-reduction per column (norm, trailing products and the
//     T-factor products come out of the same reduction), compact-WY T built on the fly;
//   * update kernel (caqr_update_kernel): W = V^T C, W' = op(T) W, C -= V W' as FP64 DMMA
//     (mma.sync m16n8k16.f64) GEMMs out of shared memory, tile staged with cp.async.
//
// The executable specification of exactly this decomposition is tests/model_caqr.py.
#include "pl_common.cuh"
#include "caqr.h"
#include <vector>

namespace pl {

// =============================================================================================
// planner
// =============================================================================================
Plan make_plan(int64_t m, int64_t n) {
  Plan P;
  P.m = m; P.n = n;
  P.npad = round_up(n, NB);
  P.K = (int)(P.npad / NB);
  P.mrows = m + P.npad + NB;
  P.t_tiles = 0; P.vup_tiles = 0;
  P.panels.resize(P.K);
  for (int p = 0; p < P.K; p++) {
    int64_t m_act = m - (int64_t)p * NB;
    int64_t nblk = ceil_div(m_act, NB), bs = NB;
    int li = 0;
    while (true) {
      Level L;
      L.nblk = nblk; L.bs = bs;
      L.ntiles = ceil_div(nblk, G);
      if (L.ntiles > 148) {
        int64_t s = L.ntiles / 592;
        L.s = (int)(s < 1 ? 1 : (s > SMAX ? SMAX : s));
      } else {
        L.s = (int)(L.ntiles < 4 ? L.ntiles : 4);
      }
      L.nstrips = ceil_div(L.ntiles, L.s);
      L.t_off = P.t_tiles; P.t_tiles += L.ntiles;
      if (li > 0) { L.v_off = P.vup_tiles; P.vup_tiles += L.ntiles; } else L.v_off = -1;
      P.panels[p].push_back(L);
      if (L.nstrips == 1) break;
      nblk = L.nstrips; bs = bs * G * L.s; li++;
    }
  }
  return P;
}

// =============================================================================================
// panel kernel
// =============================================================================================
// 160 threads: warp 0 holds the NB x NB pivot block Rp (lane = row), warps 1..4 hold the tile
// body (one row per thread, the 32 panel entries of the row live in registers).
__global__ void __launch_bounds__(160, 3)
caqr_panel_kernel(double* __restrict__ Vb, int64_t ld, int64_t row0, int col0, int64_t nblk, int64_t bs,
                  int64_t ntiles, int s, int upper, double* __restrict__ Tl, double* __restrict__ Vupl) {
  __shared__ double red[2][5][32];
  __shared__ __align__(16) double prow[2][32];
  __shared__ __align__(16) double wbuf[5][32];
  __shared__ double Ts[32][33];
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t0 = (int64_t)blockIdx.x * s;
  const int64_t pivblk = t0 * G;
  double a[32];

  // pivot block -> warp 0
  if (warp == 0) {
    const double2* src = reinterpret_cast<const double2*>(Vb + (row0 + pivblk * bs + lane) * ld + col0);
#pragma unroll
    for (int k = 0; k < 16; k++) { double2 v = src[k]; a[2 * k] = v.x; a[2 * k + 1] = v.y; }
    if (upper) {
#pragma unroll
      for (int k = 0; k < 32; k++) if (k < lane) a[k] = 0.0;
    }
  }

  for (int i = 0; i < s; i++) {
    const int64_t t = t0 + i;
    if (t >= ntiles) break;
    // ---- load the body rows of this tile
    int q = -1;
    if (warp >= 1) q = (i == 0) ? (warp <= 3 ? warp : -1) : (warp - 1);
    const int64_t kblk = t * G + q;
    const bool valid = (q >= 0) && (kblk < nblk);
    double* rowp = Vb + (row0 + (valid ? kblk : 0) * bs + lane) * ld + col0;
    if (warp >= 1) {
      if (valid) {
        const double2* src = reinterpret_cast<const double2*>(rowp);
#pragma unroll
        for (int k = 0; k < 16; k++) { double2 v = src[k]; a[2 * k] = v.x; a[2 * k + 1] = v.y; }
        if (upper) {
#pragma unroll
          for (int k = 0; k < 32; k++) if (k < lane) a[k] = 0.0;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; k++) a[k] = 0.0;
      }
    }
    for (int e = threadIdx.x; e < 32 * 33; e += 160) (&Ts[0][0])[e] = 0.0;
    __syncthreads();

    // ---- 32 Householder steps
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int buf = j & 1;
      double x;
      if (warp == 0) x = (lane > j) ? a[j] : 0.0; else x = a[j];
      // batched reduction of x * a[k], k = 0..31 (lane k ends up with the warp total of index k)
      double v[16];
      {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          double lo = x * a[k], hi = x * a[k + 16];
          double send = up ? lo : hi, keep = up ? hi : lo;
          v[k] = keep + __shfl_xor_sync(FULL, send, 16);
        }
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int k = 0; k < off; k++) {
          double send = up ? v[k] : v[k + off], keep = up ? v[k + off] : v[k];
          v[k] = keep + __shfl_xor_sync(FULL, send, off);
        }
      }
      red[buf][warp][lane] = v[0];
      if (warp == 0 && lane == j) {
#pragma unroll
        for (int k = 0; k < 16; k++) reinterpret_cast<double2*>(prow[buf])[k] = make_double2(a[2 * k], a[2 * k + 1]);
      }
      __syncthreads();
      const double tot = red[buf][0][lane] + red[buf][1][lane] + red[buf][2][lane] + red[buf][3][lane] + red[buf][4][lane];
      const double alpha = prow[buf][j];
      const double sigma2 = __shfl_sync(FULL, tot, j);
      double beta = alpha, tau = 0.0, scale = 0.0;
      if (sigma2 != 0.0) {
        beta = -copysign(sqrt(alpha * alpha + sigma2), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      const double zz = prow[buf][lane] + scale * tot;   // lane>j: pre-w ; lane<j: T-factor product
      wbuf[warp][lane] = (lane > j) ? tau * zz : 0.0;
      __syncwarp();
      double vr;
      if (warp == 0) vr = (lane > j) ? x * scale : ((lane == j) ? 1.0 : 0.0); else vr = x * scale;
      if (j < 31) {
#pragma unroll
        for (int k = (j + 2) & ~1; k < 32; k += 2) {
          double2 w2 = *reinterpret_cast<const double2*>(&wbuf[warp][k]);
          a[k] -= vr * w2.x; a[k + 1] -= vr * w2.y;
        }
        if (((j + 1) & 1)) a[j + 1] -= vr * wbuf[warp][j + 1];
      }
      if (warp == 0) { if (lane > j) a[j] = vr; else if (lane == j) a[j] = beta; } else a[j] = vr;
      // compact-WY T, column j (warp 4): T[0:j,j] = -tau * T[0:j,0:j] * z,  T[j][j] = tau
      if (warp == 4) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < j; l++) {
          double zl = __shfl_sync(FULL, zz, l);
          acc += Ts[lane][l] * zl;
        }
        if (lane < j) Ts[lane][j] = -tau * acc;
        if (lane == j) Ts[lane][j] = tau;
        __syncwarp();
      }
      __syncwarp();
    }
    __syncthreads();

    // ---- write reflectors, T
    double* Tt = Tl + t * (NB * NB);
    for (int e = threadIdx.x; e < NB * NB; e += 160) Tt[e] = Ts[e >> 5][e & 31];
    if (!upper) {
      if (warp >= 1 && valid) {
        double2* dst = reinterpret_cast<double2*>(rowp);
#pragma unroll
        for (int k = 0; k < 16; k++) dst[k] = make_double2(a[2 * k], a[2 * k + 1]);
      }
      if (warp == 0 && i == 0) {   // strictly-lower part of the pivot block = reflector entries
        double* dst = Vb + (row0 + pivblk * bs + lane) * ld + col0;
#pragma unroll
        for (int k = 0; k < 32; k++) if (k < lane) dst[k] = a[k];
      }
    } else {
      double* Vt = Vupl + t * (TB * NB);
      if (warp >= 1 && q >= 0) {   // body block q of the tile (zeros when the block is missing)
        double2* dst = reinterpret_cast<double2*>(Vt + (q * NB + lane) * NB);
#pragma unroll
        for (int k = 0; k < 16; k++) dst[k] = make_double2(a[2 * k], a[2 * k + 1]);
      }
      if (i == 0) {
        if (warp == 0) {           // explicit unit-lower pivot block
          double* dst = Vt + lane * NB;
#pragma unroll
          for (int k = 0; k < 32; k++) dst[k] = (k < lane) ? a[k] : ((k == lane) ? 1.0 : 0.0);
        }
        if (warp == 4) {           // tile 0 has only 3 body blocks; block slot 0 is the pivot
          // nothing: slots 1..3 written by warps 1..3
        }
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 32; k++) if (k < lane) a[k] = 0.0;   // carry only the triangle
    }
    __syncthreads();
  }
  // ---- R of the strip -> upper triangle of its pivot block
  if (warp == 0) {
    double* dst = Vb + (row0 + pivblk * bs + lane) * ld + col0;
#pragma unroll
    for (int k = 0; k < 32; k++) if (k >= lane) dst[k] = a[k];
  }
}

// =============================================================================================
// update kernel (DMMA)
// =============================================================================================
constexpr int SP = 36;   // padded row stride (doubles) of every 32-column shared tile: 36 = 4 mod 16
struct UpdSmem {
  double Vs[G][NB][SP];
  double Cs[G][NB][SP];
  double Zs[NB][SP];
  double Ws[NB][SP];
  double Wp[NB][SP];
  double Ts[NB][SP];
};

struct UpdArgs {
  const double* Vb; int64_t ld; int64_t row0; int col0;
  int64_t nblk, bs, ntiles; int s; int upper; int forward;
  const double* Tl; const double* Vupl;
  double* C0; int64_t ldc0; int coff0; int nchunk0;
  double* C1; int64_t ldc1; int coff1;
};

__global__ void __launch_bounds__(256, 2) caqr_update_kernel(UpdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UpdSmem& S = *reinterpret_cast<UpdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int chunk = blockIdx.x;
  double* Cb; int64_t ldc; int coff;
  if (chunk < A.nchunk0) { Cb = A.C0; ldc = A.ldc0; coff = A.coff0 + NB * chunk; }
  else { Cb = A.C1; ldc = A.ldc1; coff = A.coff1 + NB * (chunk - A.nchunk0); }
  const int64_t t0 = (int64_t)blockIdx.y * A.s;
  int cnt = A.s;
  if (t0 + cnt > A.ntiles) cnt = (int)(A.ntiles - t0);
  const int64_t pivrow = A.row0 + t0 * G * A.bs;

  if (!A.forward) {   // backward: carried rows come from memory
    for (int e = tid; e < NB * 16; e += 256) {
      int r = e >> 4, c2 = (e & 15) * 2;
      cp_async16(&S.Zs[r][c2], Cb + (pivrow + r) * ldc + coff + c2, true);
    }
  }

  for (int it = 0; it < cnt; it++) {
    const int i = A.forward ? it : (cnt - 1 - it);
    const int64_t t = t0 + i;
    const bool first = (i == 0);
    // ---- stage V, C, T of this tile
    for (int e = tid; e < G * NB * 16; e += 256) {
      const int q = e >> 9, r = (e >> 4) & 31, c2 = (e & 15) * 2;
      const int64_t kblk = t * G + q;
      const bool valid = kblk < A.nblk;
      const int64_t grow = A.row0 + (valid ? kblk : 0) * A.bs + r;
      const double* vsrc = A.upper ? (A.Vupl + (t * TB + q * NB + r) * NB + c2)
                                   : (A.Vb + grow * A.ld + A.col0 + c2);
      cp_async16(&S.Vs[q][r][c2], vsrc, valid || A.upper);
      if (first && q == 0) {
        if (A.forward) cp_async16(&S.Zs[r][c2], Cb + grow * ldc + coff + c2, true);
      } else {
        cp_async16(&S.Cs[q][r][c2], Cb + grow * ldc + coff + c2, valid);
      }
    }
    {
      const double* Tt = A.Tl + t * (NB * NB);
      for (int e = tid; e < NB * 16; e += 256) {
        int r = e >> 4, c2 = (e & 15) * 2;
        cp_async16(&S.Ts[r][c2], Tt + r * NB + c2, true);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (first && !A.upper) {   // explicit unit-lower pivot block from the in-place storage
      for (int e = tid; e < NB * NB; e += 256) {
        int r = e >> 5, c = e & 31;
        if (c > r) S.Vs[0][r][c] = 0.0; else if (c == r) S.Vs[0][r][c] = 1.0;
      }
      __syncthreads();
    }

    // ---- GEMM1: W = V^T C (+ Z)      W is NB x NB: warp -> (m16 block mb, n8 group ng)
    const int mb = warp >> 2, ng = warp & 3;
    {
      double acc[4] = {0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < G; q++) {
        const double (*Cq)[SP] = (first && q == 0) ? S.Zs : S.Cs[q];
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          double fa[8], fb[4];
#pragma unroll
          for (int x = 0; x < 8; x++) fa[x] = S.Vs[q][t4 + 4 * (x >> 1) + 16 * kk][16 * mb + g + 8 * (x & 1)];
#pragma unroll
          for (int x = 0; x < 4; x++) fb[x] = Cq[t4 + 4 * x + 16 * kk][8 * ng + g];
          mma16816(acc, fa, fb);
        }
      }
      const int r = 16 * mb + g, c = 8 * ng + 2 * t4;
      if (!first) {
        acc[0] += S.Zs[r][c]; acc[1] += S.Zs[r][c + 1];
        acc[2] += S.Zs[r + 8][c]; acc[3] += S.Zs[r + 8][c + 1];
      }
      *reinterpret_cast<double2*>(&S.Ws[r][c]) = make_double2(acc[0], acc[1]);
      *reinterpret_cast<double2*>(&S.Ws[r + 8][c]) = make_double2(acc[2], acc[3]);
    }
    __syncthreads();
    // ---- W' = op(T) W     forward: T^T, backward: T
    {
      double acc[4] = {0, 0, 0, 0};
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
        double fa[8], fb[4];
#pragma unroll
        for (int x = 0; x < 8; x++) {
          const int mm = 16 * mb + g + 8 * (x & 1), kq = t4 + 4 * (x >> 1) + 16 * kk;
          fa[x] = A.forward ? S.Ts[kq][mm] : S.Ts[mm][kq];
        }
#pragma unroll
        for (int x = 0; x < 4; x++) fb[x] = S.Ws[t4 + 4 * x + 16 * kk][8 * ng + g];
        mma16816(acc, fa, fb);
      }
      const int r = 16 * mb + g, c = 8 * ng + 2 * t4;
      *reinterpret_cast<double2*>(&S.Wp[r][c]) = make_double2(acc[0], acc[1]);
      *reinterpret_cast<double2*>(&S.Wp[r + 8][c]) = make_double2(acc[2], acc[3]);
    }
    __syncthreads();
    // ---- GEMM2: C -= V W'     warp -> (slab q, column half h)
    {
      const int q = warp >> 1, h = warp & 1;
      const bool piv = first && q == 0;
      double (*Cq)[SP] = piv ? S.Zs : S.Cs[q];
      const int64_t kblk = t * G + q;
      const bool valid = kblk < A.nblk;
      double fa[2][2][8];
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++)
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
#pragma unroll
          for (int x = 0; x < 8; x++) fa[m2][kk][x] = S.Vs[q][16 * m2 + g + 8 * (x & 1)][t4 + 4 * (x >> 1) + 16 * kk];
#pragma unroll
      for (int nn = 0; nn < 2; nn++) {
        double fb[2][4];
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
#pragma unroll
          for (int x = 0; x < 4; x++) fb[kk][x] = -S.Wp[t4 + 4 * x + 16 * kk][16 * h + 8 * nn + g];
#pragma unroll
        for (int m2 = 0; m2 < 2; m2++) {
          const int r = 16 * m2 + g, c = 16 * h + 8 * nn + 2 * t4;
          double2 c01 = *reinterpret_cast<const double2*>(&Cq[r][c]);
          double2 c23 = *reinterpret_cast<const double2*>(&Cq[r + 8][c]);
          double acc[4] = {c01.x, c01.y, c23.x, c23.y};
          mma16816(acc, fa[m2][0], fb[0]);
          mma16816(acc, fa[m2][1], fb[1]);
          if (piv) {
            *reinterpret_cast<double2*>(&S.Zs[r][c]) = make_double2(acc[0], acc[1]);
            *reinterpret_cast<double2*>(&S.Zs[r + 8][c]) = make_double2(acc[2], acc[3]);
          } else if (valid) {
            double* dst = Cb + (A.row0 + kblk * A.bs + r) * ldc + coff + c;
            *reinterpret_cast<double2*>(dst) = make_double2(acc[0], acc[1]);
            *reinterpret_cast<double2*>(dst + 8 * ldc) = make_double2(acc[2], acc[3]);
          }
        }
      }
    }
    if (!first) {
      for (int e = tid; e < NB * NB; e += 256) { int r = e >> 5, c = e & 31; S.Zs[r][c] -= S.Wp[r][c]; }
    }
    __syncthreads();
  }
  // ---- carried rows back to the pivot block rows
  for (int e = tid; e < NB * 16; e += 256) {
    int r = e >> 4, c2 = (e & 15) * 2;
    *reinterpret_cast<double2*>(Cb + (pivrow + r) * ldc + coff + c2) = *reinterpret_cast<const double2*>(&S.Zs[r][c2]);
  }
}

// =============================================================================================
// small helper kernels
// =============================================================================================
// R (n x n, ld n) <- upper triangle of Vb[0:n, 0:n]
__global__ void extract_r_kernel(double* R, int64_t ldr, const double* Vb, int64_t ld, int n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  int r = (int)(idx / n), c = (int)(idx % n);
  R[(int64_t)r * ldr + c] = (c >= r) ? Vb[(int64_t)r * ld + c] : 0.0;
}
// zero the R entries to the right of each diagonal block (rows < npad)
__global__ void zero_r_right_kernel(double* Vb, int64_t ld, int npad) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)npad * npad) return;
  int r = (int)(idx / npad), c = (int)(idx % npad);
  if (c >= (r / NB + 1) * NB) Vb[(int64_t)r * ld + c] = 0.0;
}
// Ptmp rows [row0, mrows): zero, identity block at row0
__global__ void ptmp_init_kernel(double* Ptmp, int64_t row0, int64_t mrows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // double2 index
  int64_t tot = (mrows - row0) * (NB / 2);
  for (; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx / (NB / 2); int c = (int)(idx % (NB / 2)) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (r < NB) { if (c == r) v.x = 1.0; if (c + 1 == r) v.y = 1.0; }
    reinterpret_cast<double2*>(Ptmp + (row0 + r) * NB)[c >> 1] = v;
  }
}
// Vb[:, col0:col0+NB] <- Ptmp (rows >= row0), 0 (rows < row0)
__global__ void ptmp_copyback_kernel(double* Vb, int64_t ld, int col0, const double* Ptmp, int64_t row0, int64_t mrows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = mrows * (NB / 2);
  for (; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx / (NB / 2); int c = (int)(idx % (NB / 2)) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (r >= row0) v = reinterpret_cast<const double2*>(Ptmp + r * NB)[c >> 1];
    *reinterpret_cast<double2*>(Vb + r * ld + col0 + c) = v;
  }
}

// =============================================================================================
// drivers
// =============================================================================================
static int launch_update(const Plan& P, int p, const Level& L, int li, const double* Vb, const double* Tws,
                         const double* Vup, double* C0, int64_t ldc0, int coff0, int nchunk0, double* C1, int64_t ldc1,
                         int coff1, int nchunk1, int forward, cudaStream_t st) {
  if (nchunk0 + nchunk1 <= 0) return 0;
  UpdArgs A;
  A.Vb = Vb; A.ld = P.npad; A.row0 = (int64_t)p * NB; A.col0 = p * NB;
  A.nblk = L.nblk; A.bs = L.bs; A.ntiles = L.ntiles; A.s = L.s; A.upper = li > 0; A.forward = forward;
  A.Tl = Tws + L.t_off * (NB * NB);
  A.Vupl = (li > 0) ? (Vup + L.v_off * (TB * NB)) : nullptr;
  A.C0 = C0; A.ldc0 = ldc0; A.coff0 = coff0; A.nchunk0 = nchunk0;
  A.C1 = C1; A.ldc1 = ldc1; A.coff1 = coff1;
  static bool attr_set = false;
  if (!attr_set) {
    PL_CUDA(cudaFuncSetAttribute(caqr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UpdSmem)));
    attr_set = true;
  }
  // strips on grid.y (<= 65535): split very long strip lists over several launches
  int64_t done = 0;
  while (done < L.nstrips) {
    int64_t ny = L.nstrips - done; if (ny > 65535) ny = 65535;
    UpdArgs B = A;
    // shift the strip origin by `done` strips: expressed through row0/tile offsets
    B.row0 = A.row0 + done * L.s * G * L.bs;
    B.nblk = L.nblk - done * L.s * G;
    B.ntiles = L.ntiles - done * L.s;
    B.Tl = A.Tl + done * L.s * (NB * NB);
    if (B.Vupl) B.Vupl = A.Vupl + done * L.s * (TB * NB);
    dim3 grid((unsigned)(nchunk0 + nchunk1), (unsigned)ny);
    {
      ProfScope ps(forward ? PROF_UPDATE_F : PROF_UPDATE_Q, st);
      caqr_update_kernel<<<grid, 256, sizeof(UpdSmem), st>>>(B);
    }
    PL_LAUNCH_CHECK();
    done += ny;
  }
  return 0;
}

int caqr_factor(const Plan& P, double* Vb, double* Tws, double* Vup, cudaStream_t st) {
  for (int p = 0; p < P.K; p++) {
    const int64_t row0 = (int64_t)p * NB; const int col0 = p * NB;
    const int ntrail = (int)((P.npad - col0 - NB) / NB);
    for (size_t li = 0; li < P.panels[p].size(); li++) {
      const Level& L = P.panels[p][li];
      {
        ProfScope ps(PROF_PANEL, st);
        caqr_panel_kernel<<<(unsigned)L.nstrips, 160, 0, st>>>(Vb, P.npad, row0, col0, L.nblk, L.bs, L.ntiles, L.s,
                                                               li > 0, Tws + L.t_off * (NB * NB),
                                                               li > 0 ? Vup + L.v_off * (TB * NB) : nullptr);
      }
      PL_LAUNCH_CHECK();
      int rc = launch_update(P, p, L, (int)li, Vb, Tws, Vup, nullptr, 0, 0, 0, Vb, P.npad, col0 + NB, ntrail, 1, st);
      if (rc) return rc;
    }
  }
  return 0;
}

int caqr_extract_r(const Plan& P, const double* Vb, double* R, int64_t ldr, cudaStream_t st) {
  int64_t tot = P.n * P.n;
  extract_r_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(R, ldr, Vb, P.npad, (int)P.n);
  PL_LAUNCH_CHECK();
  return 0;
}

// Overwrite the reflectors in Vb with the explicit thin Q (rows < m, columns < n are meaningful).
int caqr_form_q(const Plan& P, double* Vb, const double* Tws, const double* Vup, double* Ptmp, cudaStream_t st) {
  {
    int64_t tot = P.npad * P.npad;
    zero_r_right_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(Vb, P.npad, (int)P.npad);
    PL_LAUNCH_CHECK();
  }
  for (int p = P.K - 1; p >= 0; p--) {
    const int64_t row0 = (int64_t)p * NB; const int col0 = p * NB;
    const int ntrail = (int)((P.npad - col0 - NB) / NB);
    {
      ProfScope ps(PROF_MISC, st);
      ptmp_init_kernel<<<148 * 8, 256, 0, st>>>(Ptmp, row0, P.mrows);
    }
    PL_LAUNCH_CHECK();
    for (int li = (int)P.panels[p].size() - 1; li >= 0; li--) {
      const Level& L = P.panels[p][li];
      int rc = launch_update(P, p, L, li, Vb, Tws, Vup, Ptmp, NB, 0, 1, Vb, P.npad, col0 + NB, ntrail, 0, st);
      if (rc) return rc;
    }
    {
      ProfScope ps(PROF_MISC, st);
      ptmp_copyback_kernel<<<148 * 8, 256, 0, st>>>(Vb, P.npad, col0, Ptmp, row0, P.mrows);
    }
    PL_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace pl
