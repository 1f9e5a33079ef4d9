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
//   * panel kernel  (caqr_panel_kernel):  Householder on [R; tile], one column per lane,
//     lane-local column inner products (norm, trailing products and the T-factor products
//     come out of the same 32 FMA chains), compact-WY T built on the fly;
//   * update kernel (caqr_update_kernel): W = V^T C, W' = op(T) W, C -= V W' as FP64 DMMA
//     (mma.sync m16n8k16.f64) GEMMs out of shared memory, tile staged with cp.async.
//
// The executable specification of exactly this decomposition is tests/model_caqr.py.
#include "pl_common.cuh"
#include "caqr.h"
#include <vector>
#include <cstdlib>
#include <type_traits>

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
  P.t_tiles = 0; P.vup_tiles = 0; P.vpiv_strips = 0;
  P.panels.resize(P.K);
  for (int p = 0; p < P.K; p++) {
    int64_t m_act = m - (int64_t)p * NB;
    int64_t nblk = ceil_div(m_act, NB), bs = NB;
    int li = 0;
    while (true) {
      Level L;
      L.nblk = nblk; L.bs = bs;
      L.ntiles = ceil_div(nblk, G);
      // tiles per strip: long strips (fewer, cheaper upper levels) while >= ~4 strips per SM remain
      static const int smax = getenv("PL_SMAX") ? atoi(getenv("PL_SMAX")) : SMAX;
      static const int starget = getenv("PL_STRIPS") ? atoi(getenv("PL_STRIPS")) : 592;
      static const int ssmall = getenv("PL_SSMALL") ? atoi(getenv("PL_SSMALL")) : 2;
      if (L.ntiles > 148) {
        int64_t s = L.ntiles / starget;
        L.s = (int)(s < 1 ? 1 : (s > smax ? smax : s));
      } else {
        L.s = (int)(L.ntiles < ssmall ? L.ntiles : ssmall);
      }
      L.nstrips = ceil_div(L.ntiles, L.s);
      L.t_off = P.t_tiles; P.t_tiles += L.ntiles;
      if (li > 0) { L.v_off = P.vup_tiles; P.vup_tiles += L.ntiles; L.p_off = -1; }
      else { L.v_off = -1; L.p_off = P.vpiv_strips; P.vpiv_strips += L.nstrips; }
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
// 160 threads = 5 warps; each warp owns one NB x NB block of the stacked [Rp; tile] matrix
// (warp 0: the pivot block Rp, warps 1..4: the four body blocks of the tile).  In the body warps lane
// (cg, rg) HOLDS 4 COLUMNS x 8 ROWS of the block in registers (columns cg+8i, rows 8rg..8rg+7); the
// pivot block lives in shared memory (Rs) because it needs row access (the pivot row) as well.  With
// this layout
//   * a lane needs only the 8 entries of the pivot column x that belong to its rows (4 quarter-warp
//     broadcast loads per step instead of 32 -- the first column-per-lane version was bound by the
//     shared-memory broadcast of x) and keeps them in registers for the rank-1 update;
//   * the partial inner products x^T P_k are lane-local FMA chains; 3 shuffles fold the 4 row groups
//     so that lane k ends up with column k (norm for lane j, trailing products for lanes > j,
//     compact-WY T-factor products for lanes < j);
//   * the column loop is a RUNTIME loop (no register array is indexed by j), so the kernel is a few
//     KB of code instead of 250 KB -- the fully unrolled first version was instruction-fetch bound.
// Reflector columns are kept UNSCALED (u = x, pivot entry implied) during the 32 steps and scaled by
// their 1/(alpha - beta) once at the end.
__device__ __forceinline__ double fast_rsqrt(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double h = 0.5 * s;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = y * fma(-d, y, 2.0);
  y = y * fma(-d, y, 2.0);
  return y;
}

// One Householder column step for the body warps, slot JI (= j >> 3) static so that the register array
// is never indexed dynamically.  Lane (cg, rg) holds rows 8rg..8rg+7 of columns cg, cg+8, cg+16, cg+24.
#define PL_SLOT_SWITCH(ji, STMT)                                                              \
  switch (ji) { case 0: { constexpr int JI = 0; STMT } break; case 1: { constexpr int JI = 1; STMT } break; \
                case 2: { constexpr int JI = 2; STMT } break; default: { constexpr int JI = 3; STMT } break; }

#ifdef PL_PANEL_TIMING
__device__ unsigned long long g_panel_dbg[8];
extern "C" int pl_debug_panel_read(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_panel_dbg, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_panel_dbg, z, sizeof(z));
  return 0;
}
#define PT_DECL long long tacc0 = 0, tacc1 = 0, tacc2 = 0, tacc3 = 0, tacc4 = 0, tacc5 = 0, tprev = clock64();
#define PT_MARK(k) do { long long _t = clock64(); tacc##k += _t - tprev; tprev = _t; } while (0)
#define PT_FLUSH do { if (lane == 0 && warp == PT_WARP) { atomicAdd(&g_panel_dbg[0], (unsigned long long)tacc0); atomicAdd(&g_panel_dbg[1], (unsigned long long)tacc1); \
  atomicAdd(&g_panel_dbg[2], (unsigned long long)tacc2); atomicAdd(&g_panel_dbg[3], (unsigned long long)tacc3); atomicAdd(&g_panel_dbg[4], (unsigned long long)tacc4); \
  atomicAdd(&g_panel_dbg[5], (unsigned long long)tacc5); } } while (0)
#else
#define PT_DECL
#define PT_MARK(k)
#define PT_FLUSH
#endif
template <int MINB>
__global__ void __launch_bounds__(160, MINB)
caqr_panel_kernel(double* __restrict__ Vb, int64_t ld, int64_t row0, int col0, int64_t nblk, int64_t bs,
                  int64_t ntiles, int s, int upper, double* __restrict__ Tl, double* __restrict__ Vupl,
                  double* __restrict__ Vpivl, const double* __restrict__ Asrc) {
  __shared__ __align__(16) double xs[2][5][32];   // pivot column of step j (buffer j&1), pre-published during step j-1
  __shared__ double red[2][5][32];
  __shared__ double prow[2][32];
  __shared__ double Rs[32][33];      // pivot block (needs row AND column access -> shared memory)
  __shared__ double Ts[32][33];
  __shared__ __align__(16) double zsm[32];
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = lane & 7, rg = lane >> 3;     // body warps: column group / row group of this lane
  const int64_t t0 = (int64_t)blockIdx.x * s;
  const int64_t pivblk = t0 * G;
  double a[4][8];        // body warps: a[i][r] = P[8*rg + r][cg + 8*i] of this warp's NB x NB block
  double mysc = 0.0;     // 1/(alpha-beta) of column `lane` (every lane owns the bookkeeping of one column)
  PT_DECL

  // Asrc != nullptr (first panel of a factorisation that was not preceded by a copy): the input is read from the
  // caller's matrix (same leading dimension), everything is written to Vb.
  const double* Lb = Asrc ? Asrc : Vb;
  if (warp == 0) {   // pivot block -> shared memory (row-wise, coalesced)
    const double* src = Lb + (row0 + pivblk * bs) * ld + col0 + lane;
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const double v = src[(int64_t)r * ld];
      Rs[r][lane] = (upper && r > lane) ? 0.0 : v;
    }
  }

  for (int i = 0; i < s; i++) {
    const int64_t t = t0 + i;
    if (t >= ntiles) break;
    const bool dense_piv = (i == 0);          // later tiles see the carried triangle (x = 0 in the pivot block)
    const int twarp = dense_piv ? 4 : 0;      // the warp with spare time builds T
    int q = -1;
    if (warp >= 1) q = (i == 0) ? (warp <= 3 ? warp : -1) : (warp - 1);
    const int64_t kblk = t * G + q;
    const bool valid = (q >= 0) && (kblk < nblk);
    const int64_t blko = (row0 + (valid ? kblk : 0) * bs + 8 * rg) * ld + col0 + cg;
    double* blkp = Vb + blko;
    if (warp >= 1) {
      if (valid) {
        const double* lp = Lb + blko;
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int ii = 0; ii < 4; ii++) a[ii][r] = lp[(int64_t)r * ld + 8 * ii];
        if (upper) {
#pragma unroll
          for (int r = 0; r < 8; r++)
#pragma unroll
            for (int ii = 0; ii < 4; ii++) if (8 * rg + r > cg + 8 * ii) a[ii][r] = 0.0;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int ii = 0; ii < 4; ii++) a[ii][r] = 0.0;
      }
    }
    for (int e = threadIdx.x; e < 32 * 33; e += 160) (&Ts[0][0])[e] = 0.0;
    mysc = 0.0;
    if (warp >= 1 && cg == 0) {   // column 0 of the body blocks for step 0 (later columns are pre-published)
#pragma unroll
      for (int r = 0; r < 8; r += 2) *reinterpret_cast<double2*>(&xs[0][warp][8 * rg + r]) = make_double2(a[0][r], a[0][r + 1]);
    }
    __syncthreads();

    PT_MARK(0);
#pragma unroll 1
    for (int j = 0; j < 32; j++) {
      const int buf = j & 1;
      double x[8];       // body warps: the pivot column restricted to this lane's 8 rows
      double dsum = 0.0;
      // 1. lane-local partial inner products x^T P_k and their reduction over the 4 row groups
      if (warp == 0) {
        prow[buf][lane] = Rs[j][lane];
        if (dense_piv) {
          xs[buf][0][lane] = (lane > j) ? Rs[lane][j] : 0.0;     // only rows below the pivot are active
          __syncwarp();
          double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
          for (int r = 0; r < 32; r += 4) {
            const double2 xa = *reinterpret_cast<const double2*>(&xs[buf][0][r]);
            const double2 xb = *reinterpret_cast<const double2*>(&xs[buf][0][r + 2]);
            d0 = fma(xa.x, Rs[r][lane], d0); d1 = fma(xa.y, Rs[r + 1][lane], d1);
            d2 = fma(xb.x, Rs[r + 2][lane], d2); d3 = fma(xb.y, Rs[r + 3][lane], d3);
          }
          dsum = (d0 + d1) + (d2 + d3);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; r += 2) {
          const double2 xv = *reinterpret_cast<const double2*>(&xs[buf][warp][8 * rg + r]);
          x[r] = xv.x; x[r + 1] = xv.y;
        }
        double p[4], pb[4];
#pragma unroll
        for (int ii = 0; ii < 4; ii++) { p[ii] = 0.0; pb[ii] = 0.0; }
#pragma unroll
        for (int r = 0; r < 8; r += 2)
#pragma unroll
          for (int ii = 0; ii < 4; ii++) { p[ii] = fma(x[r], a[ii][r], p[ii]); pb[ii] = fma(x[r + 1], a[ii][r + 1], pb[ii]); }
#pragma unroll
        for (int ii = 0; ii < 4; ii++) p[ii] += pb[ii];
        // transposed reduction over the row groups: lane (cg, rg) ends with the total of column cg + 8*rg = lane
        const bool hi2 = (rg & 2) != 0, hi1 = (rg & 1) != 0;
        const double s0 = hi2 ? p[0] : p[2], s1 = hi2 ? p[1] : p[3];
        const double k0 = hi2 ? p[2] : p[0], k1 = hi2 ? p[3] : p[1];
        const double q0 = k0 + __shfl_xor_sync(FULL, s0, 16), q1 = k1 + __shfl_xor_sync(FULL, s1, 16);
        const double s2 = hi1 ? q0 : q1, k2 = hi1 ? q1 : q0;
        dsum = k2 + __shfl_xor_sync(FULL, s2, 8);
      }
      PT_MARK(1);
      red[buf][warp][lane] = dsum;
#ifdef PL_SKIP_BAR
      __syncwarp();
#else
      __syncthreads();
#endif
      PT_MARK(2);
      const double tot = (red[buf][0][lane] + red[buf][1][lane]) + (red[buf][2][lane] + red[buf][3][lane]) + red[buf][4][lane];
      const double alpha = prow[buf][j];
      const double sigma2 = __shfl_sync(FULL, tot, j);
      double beta = alpha, tau = 0.0, scale = 0.0;
#ifdef PL_SKIP_SCALARS
      if (false) {
#else
      if (sigma2 != 0.0) {
#endif
        const double s2 = fma(alpha, alpha, sigma2);
        if (s2 > 1e-280 && s2 < 1e280) {
          // |beta| = sqrt(s2), scale = 1/(alpha - beta) = sgn/(|alpha| + |beta|), tau = (|alpha| + |beta|)/|beta|.
          // The reciprocal is seeded from the UNREFINED rsqrt so that its MUFU overlaps the rsqrt Newton steps.
          double y, rc;
          asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s2));
          const double aa = fabs(alpha);
          const double dd0 = fma(s2, y, aa);
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dd0));
          const double h = 0.5 * s2;
          y = y * fma(-h * y, y, 1.5);
          y = y * fma(-h * y, y, 1.5);                   // 1 / |beta|   (2^-22 -> 2^-43 -> full)
          const double nrm = s2 * y;
          const double dd = aa + nrm;
          rc = rc * fma(-dd, rc, 2.0);
          rc = rc * fma(-dd, rc, 2.0);                   // 1 / (|alpha| + |beta|)
          beta = -copysign(nrm, alpha);
          scale = copysign(rc, alpha);
          tau = dd * y;
        } else {
          beta = -copysign(sqrt(s2), alpha);
          tau = (beta - alpha) / beta;
          scale = 1.0 / (alpha - beta);
        }
      }
      if (lane == j) mysc = scale;
      const double zz = fma(scale, tot, prow[buf][lane]);
      const double w = (lane > j) ? tau * zz : 0.0;
      // 2. trailing update P[r][k] -= (x_r * scale) * w_k, and pre-publication of the next pivot column
      const double sw = scale * w;
      PT_MARK(3);
      if (warp == 0) {
        if (dense_piv) {
#pragma unroll
          for (int r = 0; r < 32; r += 2) {
            const double2 xa = *reinterpret_cast<const double2*>(&xs[buf][0][r]);
            Rs[r][lane] = fma(-xa.x, sw, Rs[r][lane]);
            Rs[r + 1][lane] = fma(-xa.y, sw, Rs[r + 1][lane]);
          }
        }
        Rs[j][lane] = (lane == j) ? beta : (Rs[j][lane] - w);    // pivot row (v = 1), new diagonal
      } else {
        double swi[4];
#pragma unroll
        for (int ii = 0; ii < 4; ii++) swi[ii] = __shfl_sync(FULL, sw, cg + 8 * ii);
#ifndef PL_SKIP_UPDATE
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int ii = 0; ii < 4; ii++) a[ii][r] = fma(-x[r], swi[ii], a[ii][r]);
#endif
        const int jn = j + 1;
        if (jn < 32 && cg == (jn & 7)) {   // the 4 lanes (one per row group) that hold the next pivot column
          PL_SLOT_SWITCH(jn >> 3,
            _Pragma("unroll") for (int r = 0; r < 8; r += 2)
              *reinterpret_cast<double2*>(&xs[buf ^ 1][warp][8 * rg + r]) = make_double2(a[JI][r], a[JI][r + 1]);
          )
        }
      }
      PT_MARK(4);
      // 3. compact-WY T, column j:  T[0:j,j] = -tau * T[0:j,0:j] * z ,  T[j][j] = tau.
      //    Columns >= j of Ts are still zero, so the sum runs over all 32 entries (static unroll).
#ifndef PL_SKIP_T
      if (warp == twarp) {
        zsm[lane] = (lane < j) ? mysc * zz : 0.0;
        __syncwarp();
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
        for (int l = 0; l < 32; l += 4) {
          const double2 za = *reinterpret_cast<const double2*>(&zsm[l]);
          const double2 zb = *reinterpret_cast<const double2*>(&zsm[l + 2]);
          c0 = fma(Ts[lane][l], za.x, c0); c1 = fma(Ts[lane][l + 1], za.y, c1);
          c2 = fma(Ts[lane][l + 2], zb.x, c2); c3 = fma(Ts[lane][l + 3], zb.y, c3);
        }
        const double acc = (c0 + c1) + (c2 + c3);
        if (lane < j) Ts[lane][j] = -tau * acc;
        if (lane == j) Ts[lane][j] = tau;
      }
#endif
      __syncwarp();
      PT_MARK(5);
    }
    __syncthreads();

    // ---- scale the reflector columns, write reflectors and T
    if (warp == 0) {
      if (dense_piv) {
#pragma unroll 8
        for (int r = 1; r < 32; r++) if (r > lane) Rs[r][lane] *= mysc;
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int ii = 0; ii < 4; ii++) {
        const double sc = __shfl_sync(FULL, mysc, cg + 8 * ii);
#pragma unroll
        for (int r = 0; r < 8; r++) a[ii][r] *= sc;
      }
    }
    double* Tt = Tl + t * (NB * NB);
    for (int e = threadIdx.x; e < NB * NB; e += 160) Tt[e] = Ts[e >> 5][e & 31];
    if (!upper) {
      if (warp >= 1 && valid) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int ii = 0; ii < 4; ii++) blkp[(int64_t)r * ld + 8 * ii] = a[ii][r];
      }
      if (warp == 0 && i == 0) {   // explicit unit-lower pivot-block reflectors -> side store (the in-place
        double* dst = Vpivl + (int64_t)blockIdx.x * (NB * NB) + lane;   // rows are reused when Q is formed)
#pragma unroll 8
        for (int r = 0; r < 32; r++) dst[r * NB] = (r > lane) ? Rs[r][lane] : ((r == lane) ? 1.0 : 0.0);
      }
    } else {
      double* Vt = Vupl + t * (TB * NB);
      if (warp >= 1 && q >= 0) {   // body block q of the tile (zeros when the block is missing)
        double* dst = Vt + (q * NB + 8 * rg) * NB + cg;
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int ii = 0; ii < 4; ii++) dst[r * NB + 8 * ii] = a[ii][r];
      }
      if (i == 0 && warp == 0) {   // explicit unit-lower pivot block
        double* dst = Vt + lane;
#pragma unroll 8
        for (int r = 0; r < 32; r++) dst[r * NB] = (r > lane) ? Rs[r][lane] : ((r == lane) ? 1.0 : 0.0);
      }
    }
    if (warp == 0 && dense_piv) {
#pragma unroll 8
      for (int r = 1; r < 32; r++) if (r > lane) Rs[r][lane] = 0.0;   // carry only the triangle
    }
    __syncthreads();
  }
  // ---- R of the strip -> upper triangle of its pivot block
  if (warp == 0) {
    double* dst = Vb + (row0 + pivblk * bs) * ld + col0 + lane;
#pragma unroll 8
    for (int r = 0; r < 32; r++) if (r <= lane) dst[(int64_t)r * ld] = Rs[r][lane];
  }
  PT_FLUSH;
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
  const double* Tl; const double* Vupl; const double* Vpivl; int virt;
  double* C0; int64_t ldc0; int coff0; int nchunk0;
  double* C1; int64_t ldc1; int coff1; int nchunk1;
  const double* Csrc;   // forward pass only: C is READ from here (same ld / offsets), written to C0/C1; nullptr = in place
  int dbg;      // timing experiments only (PL_UPD_DBG): 1 no staging, 2 no GEMM1, 4 no T step, 8 no GEMM2, 16 no stores, 32 GEMM1 fragments once, 64 no L2 prefetch
};

// Per-phase cycle counters of one warp per CTA (build with -DPL_UPD_TIMING[=warp]; read with pl_debug_update_read).
#ifdef PL_UPD_TIMING
__device__ unsigned long long g_upd_dbg[8];
extern "C" int pl_debug_update_read(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_upd_dbg, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_upd_dbg, z, sizeof(z));
  return 0;
}
#define UT_DECL long long ut[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ut_prev = clock64();
#define UT_MARK(k) do { long long _t = clock64(); ut[k] += _t - ut_prev; ut_prev = _t; } while (0)
#define UT_FLUSH do { if (lane == 0 && warp == (PL_UPD_TIMING + 0)) { for (int k = 0; k < 8; k++) atomicAdd(&g_upd_dbg[k], (unsigned long long)ut[k]); } } while (0)
#else
#define UT_DECL
#define UT_MARK(k)
#define UT_FLUSH
#endif
// Timing experiments (probes/upd_phase_probe.py): build with -DPL_UPD_EXPERIMENTS and set PL_UPD_DBG to skip phases.
#ifdef PL_UPD_EXPERIMENTS
#define UPD_DBG(bit) (A.dbg & (bit))
#else
#define UPD_DBG(bit) 0
#endif
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

  // ---- per-thread staging geometry, hoisted out of the tile loop: this thread copies the 16-byte
  //      piece (row r_lo [+16], columns c2..c2+1) of every NB x NB slab.
  const int r_lo = tid >> 4, c2 = (tid & 15) * 2;
  const bool upper = A.upper != 0;
  const int64_t v_tile = upper ? (int64_t)TB * NB : (int64_t)G * A.bs * A.ld;
  const int64_t v_q = upper ? (int64_t)NB * NB : A.bs * A.ld;
  const int64_t v_h = upper ? (int64_t)16 * NB : 16 * A.ld;
  const double* vbase = upper ? (A.Vupl + r_lo * NB + c2) : (A.Vb + (A.row0 + r_lo) * A.ld + A.col0 + c2);
  const int64_t c_tile = (int64_t)G * A.bs * ldc, c_q = A.bs * ldc, c_h = 16 * ldc;
  const double* Lb = A.Csrc ? A.Csrc : Cb;
  const double* cbase = Lb + (A.row0 + r_lo) * ldc + coff + c2;
  const double* tbase = A.Tl + r_lo * NB + c2;

  if (!A.forward) {   // backward: carried rows come from memory
    const double* zsrc = Cb + (pivrow + r_lo) * ldc + coff + c2;
    cp_async16(&S.Zs[r_lo][c2], zsrc, true);
    cp_async16(&S.Zs[r_lo + 16][c2], zsrc + c_h, true);
  }

  // GEMM role of this warp
  const int mb = warp >> 2, ng = warp & 3;     // GEMM1 / T-step: (m16 block, n8 group) of the NB x NB W
  const int gq = warp >> 1, gh = warp & 1;     // GEMM2: (slab, column half)

  UT_DECL
  for (int it = 0; it < cnt; it++) {
    const int i = A.forward ? it : (cnt - 1 - it);
    const int64_t t = t0 + i;
    const bool first = (i == 0);
    UT_MARK(7);
    // ---- stage V, C, T of this tile
    if (!UPD_DBG(1) || it == 0) {
      const double* vp = vbase + t * v_tile;
      const double* cp = cbase + t * c_tile;
#pragma unroll
      for (int q = 0; q < G; q++) {
        const bool valid = (t * G + q) < A.nblk;
        const double* vs = valid ? (vp + q * v_q) : vp;
        const double* cs = valid ? (cp + q * c_q) : cp;
        const bool vok = valid || upper;
        if (q == 0 && first && !upper) {   // explicit unit-lower pivot block of a level-0 strip
          const double* ps = A.Vpivl + (int64_t)blockIdx.y * (NB * NB) + r_lo * NB + c2;
          cp_async16(&S.Vs[0][r_lo][c2], ps, true);
          cp_async16(&S.Vs[0][r_lo + 16][c2], ps + 16 * NB, true);
        } else {
          cp_async16(&S.Vs[q][r_lo][c2], vs, vok);
          cp_async16(&S.Vs[q][r_lo + 16][c2], vs + v_h, vok);
        }
        if (q == 0 && first) {
          if (A.forward) { cp_async16(&S.Zs[r_lo][c2], cs, true); cp_async16(&S.Zs[r_lo + 16][c2], cs + c_h, true); }
        } else {
          cp_async16(&S.Cs[q][r_lo][c2], cs, valid && !A.virt);   // virt: the block is known to be zero
          cp_async16(&S.Cs[q][r_lo + 16][c2], cs + c_h, valid && !A.virt);
        }
      }
      const double* tp = tbase + t * (NB * NB);
      cp_async16(&S.Ts[r_lo][c2], tp, true);
      cp_async16(&S.Ts[r_lo + 16][c2], tp + 16 * NB, true);
    }
    cp_async_commit();
    if (it + 1 < cnt && !UPD_DBG(64)) {   // pull the next tile's rows into L2 while this tile computes (128 rows x 256 B each)
      const int64_t tn = A.forward ? t + 1 : t - 1;
      const int pr = tid >> 1, ph = (tid & 1) * 16;   // row of the tile, 128-byte half of the 256-byte row
      const int pq = pr >> 5, prr = pr & 31;
      if ((tn * G + pq) < A.nblk) {
        const double* cpf = Lb + (A.row0 + (tn * G + pq) * A.bs + prr) * ldc + coff + ph;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cpf));
        const double* vpf = upper ? (A.Vupl + (tn * TB + pr) * NB + ph)
                                  : (A.Vb + (A.row0 + (tn * G + pq) * A.bs + prr) * A.ld + A.col0 + ph);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vpf));
      }
    }
    UT_MARK(0);                 // 0: issue of the staging copies + prefetch
    cp_async_wait<0>();
    __syncthreads();
    UT_MARK(1);                 // 1: wait for the data + barrier
    const double (*C0)[SP] = first ? S.Zs : S.Cs[0];   // slab 0 of the first tile is the carried block

    // ---- GEMM1: W = V^T C (+ Z)      four independent accumulator chains (one per slab)
    if (!UPD_DBG(2)) {
      double acc[G][4];
#pragma unroll
      for (int q = 0; q < G; q++) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0; }
      const int qn = A.virt ? (first ? 1 : 0) : G;   // virtual-zero input: only the carried block (slab 0 of the first tile) contributes
      double fa[8], fb[4];
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
#pragma unroll
        for (int q = 0; q < G; q++) {
          if (q >= qn) continue;
          const double (*Cq)[SP] = (q == 0) ? C0 : S.Cs[q];
          if (!UPD_DBG(32) || (q == 0 && kk == 0)) {   // experiment 32: one fragment pair for all 8 MMAs
#pragma unroll
            for (int x = 0; x < 8; x++) fa[x] = S.Vs[q][t4 + 4 * (x >> 1) + 16 * kk][16 * mb + g + 8 * (x & 1)];
#pragma unroll
            for (int x = 0; x < 4; x++) fb[x] = Cq[t4 + 4 * x + 16 * kk][8 * ng + g];
          }
          mma16816(acc[q], fa, fb);
        }
      }
      const int r = 16 * mb + g, c = 8 * ng + 2 * t4;
      double2 z01 = make_double2(0.0, 0.0), z23 = make_double2(0.0, 0.0);
      if (!first) { z01 = *reinterpret_cast<const double2*>(&S.Zs[r][c]); z23 = *reinterpret_cast<const double2*>(&S.Zs[r + 8][c]); }
      *reinterpret_cast<double2*>(&S.Ws[r][c]) = make_double2(((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0])) + z01.x,
                                                               ((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1])) + z01.y);
      *reinterpret_cast<double2*>(&S.Ws[r + 8][c]) = make_double2(((acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2])) + z23.x,
                                                                   ((acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3])) + z23.y);
    }
    UT_MARK(2);                 // 2: GEMM1 + epilogue
    __syncthreads();
    UT_MARK(3);                 // 3: barrier after GEMM1
    // ---- W' = op(T) W     forward: T^T, backward: T
    if (!UPD_DBG(4)) {
      double acc[4] = {0, 0, 0, 0}, accb[4] = {0, 0, 0, 0};
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
        double fa[8], fb[4];
        if (A.forward) {
#pragma unroll
          for (int x = 0; x < 8; x++) fa[x] = S.Ts[t4 + 4 * (x >> 1) + 16 * kk][16 * mb + g + 8 * (x & 1)];
        } else {
#pragma unroll
          for (int x = 0; x < 8; x++) fa[x] = S.Ts[16 * mb + g + 8 * (x & 1)][t4 + 4 * (x >> 1) + 16 * kk];
        }
#pragma unroll
        for (int x = 0; x < 4; x++) fb[x] = S.Ws[t4 + 4 * x + 16 * kk][8 * ng + g];
        if (kk == 0) mma16816(acc, fa, fb); else mma16816(accb, fa, fb);
      }
#pragma unroll
      for (int x = 0; x < 4; x++) acc[x] += accb[x];
      const int r = 16 * mb + g, c = 8 * ng + 2 * t4;
      *reinterpret_cast<double2*>(&S.Wp[r][c]) = make_double2(acc[0], acc[1]);
      *reinterpret_cast<double2*>(&S.Wp[r + 8][c]) = make_double2(acc[2], acc[3]);
      if (!first) {   // carried block: Z -= W' (this warp owns exactly these entries of Z)
        double2 z01 = *reinterpret_cast<const double2*>(&S.Zs[r][c]);
        double2 z23 = *reinterpret_cast<const double2*>(&S.Zs[r + 8][c]);
        z01.x -= acc[0]; z01.y -= acc[1]; z23.x -= acc[2]; z23.y -= acc[3];
        *reinterpret_cast<double2*>(&S.Zs[r][c]) = z01;
        *reinterpret_cast<double2*>(&S.Zs[r + 8][c]) = z23;
      }
    }
    UT_MARK(4);                 // 4: T step
    __syncthreads();
    UT_MARK(5);                 // 5: barrier after the T step
    // ---- GEMM2: C -= V W'     warp -> (slab gq, column half gh)
    if (!UPD_DBG(8)) {
      const bool piv = first && gq == 0;
      double (*Cq)[SP] = piv ? S.Zs : S.Cs[gq];
      const bool valid = (t * G + gq) < A.nblk;
      double* crow = Cb + (A.row0 + (t * G + gq) * A.bs + g) * ldc + coff + 16 * gh + 2 * t4;
      double fa[2][2][8];
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++)
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
#pragma unroll
          for (int x = 0; x < 8; x++) fa[m2][kk][x] = S.Vs[gq][16 * m2 + g + 8 * (x & 1)][t4 + 4 * (x >> 1) + 16 * kk];
#pragma unroll
      for (int nn = 0; nn < 2; nn++) {
        double fb[2][4];
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
#pragma unroll
          for (int x = 0; x < 4; x++) fb[kk][x] = -S.Wp[t4 + 4 * x + 16 * kk][16 * gh + 8 * nn + g];
#pragma unroll
        for (int m2 = 0; m2 < 2; m2++) {
          const int r = 16 * m2 + g, c = 16 * gh + 8 * nn + 2 * t4;
          double2 c01 = *reinterpret_cast<const double2*>(&Cq[r][c]);
          double2 c23 = *reinterpret_cast<const double2*>(&Cq[r + 8][c]);
          double acc[4] = {c01.x, c01.y, c23.x, c23.y};
          mma16816(acc, fa[m2][0], fb[0]);
          mma16816(acc, fa[m2][1], fb[1]);
          if (piv) {
            *reinterpret_cast<double2*>(&S.Zs[r][c]) = make_double2(acc[0], acc[1]);
            *reinterpret_cast<double2*>(&S.Zs[r + 8][c]) = make_double2(acc[2], acc[3]);
          } else if (valid && !UPD_DBG(16)) {
            double* dst = crow + (int64_t)(16 * m2) * ldc + 8 * nn;
            *reinterpret_cast<double2*>(dst) = make_double2(acc[0], acc[1]);
            *reinterpret_cast<double2*>(dst + 8 * ldc) = make_double2(acc[2], acc[3]);
          }
        }
      }
    }
    UT_MARK(6);                 // 6: GEMM2 + stores
    __syncthreads();
  }
  UT_MARK(7);                   // 7: barrier after GEMM2 (+ loop overhead)
  UT_FLUSH;
  // ---- carried rows back to the pivot block rows
  {
    double* zdst = Cb + (pivrow + r_lo) * ldc + coff + c2;
    *reinterpret_cast<double2*>(zdst) = *reinterpret_cast<const double2*>(&S.Zs[r_lo][c2]);
    *reinterpret_cast<double2*>(zdst + c_h) = *reinterpret_cast<const double2*>(&S.Zs[r_lo + 16][c2]);
  }
}

// =============================================================================================
// update kernel, second generation: WARP-PRIVATE COLUMNS
// =============================================================================================
// A CTA still walks one strip of tiles for a group of trailing columns, but every WARP owns 8 columns for itself
// (NW warps = NW * 8 columns per CTA).  The m8n8k4 FP64 MMA is used with the COLUMNS as the M index everywhere:
//     GEMM1   W^T (8 x 32)  = Z^T + C^T V          A = C block from REGISTERS, B = V from shared memory
//     T step  W'^T (8 x 32) = W^T op(T)^T          A = W^T from registers,      B = T from shared memory
//     GEMM2   C^T (8 x rows) -= W'^T V^T           A = W'^T from registers,     B = V from shared memory, D = C block
// The accumulator layout of one product (lane (g, t) holds D[g][2t], D[g][2t+1]) is a legal A-operand layout of the
// next one once the K index is permuted accordingly (the B operand is fetched from shared memory with the same
// permutation), so W, W', the carried block Z and the C blocks never leave the registers, nothing is transposed, and
// the warps of a CTA share only the staged V and T tiles: ONE block barrier per tile (double-buffered cp.async
// staging), no lock-step phases.  C goes from global memory straight into the MMA fragments (read once from HBM for
// GEMM1, once more from L2 for GEMM2, one block ahead of its use) and never through shared memory; every 16-byte
// shared-memory load feeds two MMAs (the M / K permutations pair adjacent reflectors), which is 44 % of the
// shared-memory bandwidth at the full DMMA rate.  V and T are stored dense with an XOR swizzle of their 16-byte
// chunks that makes all three fragment access patterns conflict free.
//
// Index conventions (g = lane >> 2, t = lane & 3):
//   C block (32 rows x 8 columns), cb[rb][e]  = C[row 8 rb + 2 t + e][column g]
//   reflector-indexed blocks (Z^T, W^T, W'^T), x[a][h] with a = 2 pr + e' = X^T[column g][reflector 16 pr + 4 t + 2 h + e']
__device__ __forceinline__ void mma884u(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
// dense NB-column tiles in shared memory, 16-byte chunk c of row r stored at chunk (c ^ swz(r))
__device__ __forceinline__ int vswz(int r) { return r & 7; }
__device__ __forceinline__ int tswz(int r) { return ((r >> 1) & 6) | ((r >> 1) & 1); }
__device__ __forceinline__ const double2* vchunk(const double* Vt, int r, int c) {
  return reinterpret_cast<const double2*>(Vt + r * NB + 2 * (c ^ vswz(r)));
}
__device__ __forceinline__ const double2* tchunk(const double* Tt, int r, int c) {
  return reinterpret_cast<const double2*>(Tt + r * NB + 2 * (c ^ tswz(r)));
}

template <int NW>
struct Upd2Smem {
  double V[2][TB * NB];
  double T[2][NB * NB];
  double2 Z[NW][4][32];      // carried block of every warp between tiles: Z[warp][a][lane] = (z[a][0], z[a][1])
};

// explicit loads: the PTX order of these (volatile) statements relative to the MMAs is the software pipeline
__device__ __forceinline__ double ldg_f64(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void lds_v2(double2& v, unsigned addr) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
}

template <int NW, bool fwd, bool VIRT>
__global__ void __launch_bounds__(NW * 32, 16 / NW) caqr_update2_kernel(UpdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Upd2Smem<NW>& S = *reinterpret_cast<Upd2Smem<NW>*>(smem_raw);
  constexpr int NT = NW * 32;
  constexpr int CPW = NB / 8;                       // warps per 32-column chunk
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  // ---- this warp's 8 columns
  const int chunk = blockIdx.x * (NW / CPW) + warp / CPW;
  const bool active = chunk < A.nchunk0 + A.nchunk1;
  const bool grp0 = chunk < A.nchunk0;
  double* Cb = grp0 ? A.C0 : A.C1;
  const int64_t ldc = grp0 ? A.ldc0 : A.ldc1;
  const int coffw = (grp0 ? A.coff0 + NB * chunk : A.coff1 + NB * (chunk - A.nchunk0)) + 8 * (warp % CPW);
  constexpr bool virt = VIRT;                      // virtual-zero launches have no second column group
  const double* Lb = A.Csrc ? A.Csrc : Cb;
  const bool upper = A.upper != 0;

  const int64_t t0 = (int64_t)blockIdx.y * A.s;
  int cnt = A.s;
  if (t0 + cnt > A.ntiles) cnt = (int)(A.ntiles - t0);
  const int64_t pivrow = A.row0 + t0 * G * A.bs;

  // ---- staging geometry: thread -> 16-byte chunk (row tid / 16 (+ k NT / 16), chunk tid % 16)
  const int sr = tid >> 4, sc = tid & 15;
  auto stage = [&](int64_t tt, bool first_tile, int buf) {
    double* Vs = S.V[buf];
    double* Ts = S.T[buf];
#pragma unroll
    for (int k = 0; k < TB / (NT / 16); k++) {
      const int r = sr + k * (NT / 16);             // row of the tile
      const int q = r >> 5, rr = r & 31;
      const bool valid = (tt * G + q) < A.nblk;
      const double* src;
      bool ok = true;
      if (upper) src = A.Vupl + (tt * TB + r) * NB + 2 * sc;
      else if (q == 0 && first_tile) src = A.Vpivl + (int64_t)blockIdx.y * (NB * NB) + rr * NB + 2 * sc;
      else { src = A.Vb + (A.row0 + ((valid ? tt * G + q : 0)) * A.bs + rr) * A.ld + A.col0 + 2 * sc; ok = valid; }
      cp_async16(Vs + r * NB + 2 * (sc ^ vswz(r)), src, ok);
    }
#pragma unroll
    for (int k = 0; k < NB / (NT / 16); k++) {
      const int r = sr + k * (NT / 16);
      cp_async16(Ts + r * NB + 2 * (sc ^ tswz(r)), A.Tl + tt * (NB * NB) + r * NB + 2 * sc, true);
    }
  };
  // ---- C blocks: (warp-uniform block base) + (lane offset) + (8 rb + e) rows.  The whole tile (4 blocks) of this
  //      warp's columns lives in registers from its load to its store; c[q] is refilled with the NEXT tile's block q
  //      right after it has been stored, and both GEMMs walk the blocks in the order 3, 2, 1, 0, so that every block has
  //      most of a tile's time to arrive.
  const int64_t lane_off = (int64_t)(2 * t) * ldc + g;
  auto blk_base = [&](const double* base, int64_t tt, int q) -> const double* {     // warp uniform
    return base + (A.row0 + (tt * G + q) * A.bs) * ldc + coffw;
  };
  auto load_blk = [&](double (&cb)[4][2], const double* ub, bool ok) {
    if (ok) {
      const double* p = ub + lane_off;
#pragma unroll
      for (int rb = 0; rb < 4; rb++)
#pragma unroll
        for (int e = 0; e < 2; e++) cb[rb][e] = ldg_f64(p + (8 * rb + e) * ldc);
    } else {
#pragma unroll
      for (int rb = 0; rb < 4; rb++) { cb[rb][0] = 0.0; cb[rb][1] = 0.0; }
    }
  };
  auto store_blk = [&](const double (&cb)[4][2], double* ub) {
    double* p = ub + lane_off;
#pragma unroll
    for (int rb = 0; rb < 4; rb++)
#pragma unroll
      for (int e = 0; e < 2; e++) p[(8 * rb + e) * ldc] = cb[rb][e];
  };
  double c[G][4][2];

  // carried block Z^T: z[a][h] = Z[row 16 pr + 4 t + 2 h + e'][column], a = 2 pr + e'; kept in shared memory between
  // tiles (the first tile of a strip, whose block 0 is the pivot block, holds it in the registers of c[0])
  double2* zs = &S.Z[warp][0][lane];               // zs[32 * a]
  auto z_ptr = [&](const double* base) -> const double* { return base + (pivrow + 4 * t) * ldc + coffw + g; };
  if (active) {
    const double* p = z_ptr(fwd ? Lb : Cb);        // forward: the pivot rows of the strip's first tile, read from the source
#pragma unroll
    for (int a = 0; a < 4; a++)
      zs[32 * a] = make_double2(p[(int64_t)(16 * (a >> 1) + (a & 1)) * ldc], p[(int64_t)(16 * (a >> 1) + 2 + (a & 1)) * ldc]);
  }

  // ---- shared-memory fragment addresses (bytes): per-lane parts of the two regular access patterns
  const unsigned sV0 = (unsigned)__cvta_generic_to_shared(&S.V[0][0]);
  const unsigned sT0 = (unsigned)__cvta_generic_to_shared(&S.T[0][0]);
  unsigned o1[2][2], o2[2][2];
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
      o1[e][pr] = (unsigned)((2 * t + e) * 256 + ((8 * pr + g) ^ (2 * t + e)) * 16);   // row 2t+e (+8k), chunk 8pr+g
      o2[pr][e] = (unsigned)(g * 256 + ((8 * pr + 2 * t + e) ^ g) * 16);               // row g (+8k), chunk 8pr+2t+e
    }
  auto voff = [&](int r, int cch) -> unsigned { return (unsigned)(r * 256 + ((cch ^ vswz(r)) * 16)); };
  auto toff = [&](int r, int cch) -> unsigned { return (unsigned)(r * 256 + ((cch ^ tswz(r)) * 16)); };

  const int i0 = fwd ? 0 : cnt - 1;
  stage(t0 + i0, i0 == 0, 0);
  cp_async_commit();
  if (!virt) {                                     // all C blocks of the first tile
#pragma unroll
    for (int qq = 0; qq < G; qq++) {
      const int q = G - 1 - qq;
      if (q == 0 && i0 == 0) continue;             // pivot block: its rows are the carried block
      load_blk(c[q], blk_base(Lb, t0 + i0, q), active && ((t0 + i0) * G + q) < A.nblk);
    }
  }

  // one tile; FIRST (the strip's first tile, whose slab 0 is the pivot block) is a compile-time constant
  auto tile = [&](auto first_c, int it, int i) {
    constexpr bool first = decltype(first_c)::value;
    const int64_t tt = t0 + i;
    const int buf = it & 1;
    cp_async_wait<0>();
    __syncthreads();                      // tile `it` has landed for everybody; everybody is done with tile it-1
    const int in = fwd ? i + 1 : i - 1;   // next tile of the strip (valid when it + 1 < cnt)
    const bool more = it + 1 < cnt;
    if (more) stage(t0 + in, in == 0, buf ^ 1);
    cp_async_commit();
    if (!active) return;
    const unsigned sV = sV0 + (unsigned)buf * (TB * NB * 8);
    const unsigned sT = sT0 + (unsigned)buf * (NB * NB * 8);
    double (&z)[4][2] = c[0];             // first tile only: the carried block in the registers of the (unused) block 0

    // ---- GEMM1: W^T = Z^T + C^T V   (first tile: slab 0 is the pivot block, whose C rows are the carried block)
    double w[4][2];
    if (first) {
#pragma unroll
      for (int a = 0; a < 4; a++) { const double2 v = zs[32 * a]; z[a][0] = v.x; z[a][1] = v.y; w[a][0] = 0.0; w[a][1] = 0.0; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int r = 16 * (a >> 1) + 4 * t + 2 * h + (a & 1);
          double2 b[2];
          lds_v2(b[0], sV + voff(r, g)); lds_v2(b[1], sV + voff(r, 8 + g));
#pragma unroll
          for (int pr = 0; pr < 2; pr++) {
            mma884u(w[2 * pr], z[a][h], b[pr].x);
            mma884u(w[2 * pr + 1], z[a][h], b[pr].y);
          }
        }
    } else {
#pragma unroll
      for (int a = 0; a < 4; a++) { const double2 v = zs[32 * a]; w[a][0] = v.x; w[a][1] = v.y; }
    }
    if (!virt) {
#pragma unroll
      for (int qq = 0; qq < G; qq++) {
        const int q = G - 1 - qq;
        if (first && q == 0) continue;
        double2 b[2][2];                                   // [k-step parity][pr]: fragments one k-step ahead
        lds_v2(b[0][0], sV + o1[0][0] + (32 * q) * 256); lds_v2(b[0][1], sV + o1[0][1] + (32 * q) * 256);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const int rb = ks >> 1, e = ks & 1;
          if (ks + 1 < 8) {
            const int rbn = (ks + 1) >> 1, en = (ks + 1) & 1;
            lds_v2(b[(ks + 1) & 1][0], sV + o1[en][0] + (32 * q + 8 * rbn) * 256);
            lds_v2(b[(ks + 1) & 1][1], sV + o1[en][1] + (32 * q + 8 * rbn) * 256);
          }
#pragma unroll
          for (int pr = 0; pr < 2; pr++) {
            mma884u(w[2 * pr], c[q][rb][e], b[ks & 1][pr].x);
            mma884u(w[2 * pr + 1], c[q][rb][e], b[ks & 1][pr].y);
          }
        }
      }
    }
    // ---- T step: W'^T = W^T op(T)^T   (forward: op(T) = T^T, backward: op(T) = T); wp holds -W'^T
    double wp[4][2];
#pragma unroll
    for (int a = 0; a < 4; a++) { wp[a][0] = 0.0; wp[a][1] = 0.0; }
    if (fwd) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int r = 16 * (a >> 1) + 4 * t + 2 * h + (a & 1);
          // T is upper triangular: its block (reflectors 16..31) x (reflectors 0..15) is zero
          double2 b[2];
          if (a < 2) lds_v2(b[0], sT + toff(r, g));
          lds_v2(b[1], sT + toff(r, 8 + g));
#pragma unroll
          for (int pr = (a < 2 ? 0 : 1); pr < 2; pr++) {
            mma884u(wp[2 * pr], w[a][h], b[pr].x);
            mma884u(wp[2 * pr + 1], w[a][h], b[pr].y);
          }
        }
    } else {
#pragma unroll
      for (int pr = 0; pr < 2; pr++) {
        // rows 16..31 of T have no entries in columns 0..15: the accumulators a2 >= 2 skip pr == 0
        const int na = (pr == 0) ? 2 : 4;
        double2 b0[4], b1[4];
#pragma unroll
        for (int a2 = 0; a2 < 4; a2++) {
          if (a2 >= na) continue;
          const int r = 16 * (a2 >> 1) + 2 * g + (a2 & 1);
          lds_v2(b0[a2], sT + toff(r, 8 * pr + 2 * t)); lds_v2(b1[a2], sT + toff(r, 8 * pr + 2 * t + 1));
        }
#pragma unroll
        for (int a2 = 0; a2 < 4; a2++) if (a2 < na) mma884u(wp[a2], w[2 * pr][0], b0[a2].x);
#pragma unroll
        for (int a2 = 0; a2 < 4; a2++) if (a2 < na) mma884u(wp[a2], w[2 * pr + 1][0], b0[a2].y);
#pragma unroll
        for (int a2 = 0; a2 < 4; a2++) if (a2 < na) mma884u(wp[a2], w[2 * pr][1], b1[a2].x);
#pragma unroll
        for (int a2 = 0; a2 < 4; a2++) if (a2 < na) mma884u(wp[a2], w[2 * pr + 1][1], b1[a2].y);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      if (!first) { const double2 v = zs[32 * a]; zs[32 * a] = make_double2(v.x - wp[a][0], v.y - wp[a][1]); }   // Z -= W'
      wp[a][0] = -wp[a][0]; wp[a][1] = -wp[a][1];
    }
    // ---- GEMM2: C^T -= W'^T V^T, same block order; a stored block's registers take the next tile's block at once
#pragma unroll
    for (int qq = 0; qq < G; qq++) {
      const int q = G - 1 - qq;
      if (first && q == 0) {
        // pivot slab: Z^T -= W'^T V0^T, D[g][2 t + h] <-> row 16 pr + 2 (2 t + h) + e'
#pragma unroll
        for (int pr = 0; pr < 2; pr++) {
          double2 b0[4], b1[4];
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const int r = 16 * (a >> 1) + 2 * g + (a & 1);
            lds_v2(b0[a], sV + voff(r, 8 * pr + 2 * t)); lds_v2(b1[a], sV + voff(r, 8 * pr + 2 * t + 1));
          }
#pragma unroll
          for (int a = 0; a < 4; a++) mma884u(z[a], wp[2 * pr][0], b0[a].x);
#pragma unroll
          for (int a = 0; a < 4; a++) mma884u(z[a], wp[2 * pr + 1][0], b0[a].y);
#pragma unroll
          for (int a = 0; a < 4; a++) mma884u(z[a], wp[2 * pr][1], b1[a].x);
#pragma unroll
          for (int a = 0; a < 4; a++) mma884u(z[a], wp[2 * pr + 1][1], b1[a].y);
        }
#pragma unroll
        for (int a = 0; a < 4; a++) zs[32 * a] = make_double2(z[a][0], z[a][1]);
      } else {
        if (virt) {
#pragma unroll
          for (int rb = 0; rb < 4; rb++) { c[q][rb][0] = 0.0; c[q][rb][1] = 0.0; }
        }
        // units (pr, j): one 16-byte fragment load per 8-row group feeds two MMAs
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int pr = u >> 1, j = u & 1;
          double2 b[4];
#pragma unroll
          for (int rb = 0; rb < 4; rb++) lds_v2(b[rb], sV + o2[pr][j] + (32 * q + 8 * rb) * 256);
#pragma unroll
          for (int rb = 0; rb < 4; rb++) mma884u(c[q][rb], wp[2 * pr][j], b[rb].x);
#pragma unroll
          for (int rb = 0; rb < 4; rb++) mma884u(c[q][rb], wp[2 * pr + 1][j], b[rb].y);
        }
        if ((tt * G + q) < A.nblk) store_blk(c[q], const_cast<double*>(blk_base(Cb, tt, q)));
      }
      if (!virt && more && !(q == 0 && in == 0))     // block q of the next tile (not its pivot block)
        load_blk(c[q], blk_base(Lb, t0 + in, q), ((t0 + in) * G + q) < A.nblk);
    }
    if (more && !virt) {                             // the tile after the next one -> L2 (this warp's 64-byte row pieces)
      const int i2 = fwd ? i + 2 : i - 2;
      if (it + 2 < cnt) {
        const int64_t tn = t0 + i2;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if ((tn * G + k) < A.nblk)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Lb + (A.row0 + (tn * G + k) * A.bs + lane) * ldc + coffw));
      }
    }
  };
  for (int it = 0; it < cnt; it++) {
    const int i = fwd ? it : (cnt - 1 - it);
    if (i == 0) tile(std::true_type{}, it, i); else tile(std::false_type{}, it, i);
  }
  // ---- carried rows back to the pivot block rows
  if (active) {
    double* p = const_cast<double*>(z_ptr(Cb));
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double2 v = zs[32 * a];
      p[(int64_t)(16 * (a >> 1) + (a & 1)) * ldc] = v.x;
      p[(int64_t)(16 * (a >> 1) + 2 + (a & 1)) * ldc] = v.y;
    }
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
// Vb[row0:row0+NB, col0:col0+NB] <- I   (input of the panel's own columns when Q is formed)
__global__ void set_identity_block_kernel(double* Vb, int64_t ld, int64_t row0, int col0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NB * NB) return;
  const int r = e >> 5, c = e & 31;
  Vb[(row0 + r) * ld + col0 + c] = (r == c) ? 1.0 : 0.0;
}

// =============================================================================================
// drivers
// =============================================================================================
static int launch_update(const Plan& P, int p, const Level& L, int li, const double* Vb, const double* Tws,
                         const double* Vup, const double* Vpiv, int virt, double* C0, int64_t ldc0, int coff0, int nchunk0, double* C1, int64_t ldc1,
                         int coff1, int nchunk1, int forward, cudaStream_t st, const double* Csrc = nullptr) {
  if (nchunk0 + nchunk1 <= 0) return 0;
  UpdArgs A;
  A.Vb = Vb; A.ld = P.npad; A.row0 = (int64_t)p * NB; A.col0 = p * NB;
  A.nblk = L.nblk; A.bs = L.bs; A.ntiles = L.ntiles; A.s = L.s; A.upper = li > 0; A.forward = forward;
  A.Tl = Tws + L.t_off * (NB * NB);
  A.Vupl = (li > 0) ? (Vup + L.v_off * (TB * NB)) : nullptr;
  A.Vpivl = (li > 0) ? nullptr : (Vpiv + L.p_off * (NB * NB));
  A.virt = virt;
  static const int upd_dbg = getenv("PL_UPD_DBG") ? atoi(getenv("PL_UPD_DBG")) : 0;
  A.dbg = upd_dbg;
  A.C0 = C0; A.ldc0 = ldc0; A.coff0 = coff0; A.nchunk0 = nchunk0;
  A.C1 = C1; A.ldc1 = ldc1; A.coff1 = coff1; A.nchunk1 = nchunk1;
  A.Csrc = Csrc;
  static DevOnce attr_set;
  if (first_on_device(attr_set)) {
    PL_CUDA(cudaFuncSetAttribute(caqr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UpdSmem)));
    const int sm2 = (int)sizeof(Upd2Smem<8>);
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<8, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
    PL_CUDA(cudaFuncSetAttribute(caqr_update2_kernel<4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
  }
  static const bool upd_old = getenv("PL_UPD_OLD") != nullptr;   // first-generation kernel (A/B timing only)
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
    if (B.Vpivl) B.Vpivl = A.Vpivl + done * (NB * NB);
    const int nch = nchunk0 + nchunk1;
    {
      ProfScope ps(forward ? PROF_UPDATE_F : PROF_UPDATE_Q, st);
      if (upd_old) {
        dim3 grid((unsigned)nch, (unsigned)ny);
        caqr_update_kernel<<<grid, 256, sizeof(UpdSmem), st>>>(B);
      } else {
        const bool one = nch == 1;
        dim3 grid(one ? 1u : (unsigned)((nch + 1) / 2), (unsigned)ny);
        const int nt = one ? 128 : 256;
        const size_t sm = sizeof(Upd2Smem<8>);
        void (*fn)(UpdArgs) = nullptr;
        const bool vt = virt != 0 && nchunk0 > 0;
        if (vt && nchunk1 > 0) { set_error("launch_update: a virtual-zero launch takes one column group"); return -1; }
#define PL_UPD2(NWv) (forward ? (vt ? caqr_update2_kernel<NWv, true, true> : caqr_update2_kernel<NWv, true, false>) \
                              : (vt ? caqr_update2_kernel<NWv, false, true> : caqr_update2_kernel<NWv, false, false>))
        fn = one ? PL_UPD2(4) : PL_UPD2(8);
        fn<<<grid, nt, sm, st>>>(B);
      }
    }
    PL_LAUNCH_CHECK();
    done += ny;
  }
  return 0;
}

static int launch_panel(const Plan& P, int p, const Level& L, int li, double* Vb, double* Tws, double* Vup, double* Vpiv,
                        cudaStream_t st, const double* Asrc = nullptr) {
  const int64_t row0 = (int64_t)p * NB; const int col0 = p * NB;
  ProfScope ps(PROF_PANEL, st);
  static const int occ = getenv("PL_PANEL_OCC") ? atoi(getenv("PL_PANEL_OCC")) : 3;
  double* Tl = Tws + L.t_off * (NB * NB);
  double* Vl = li > 0 ? Vup + L.v_off * (TB * NB) : nullptr;
  double* Pl = li > 0 ? nullptr : Vpiv + L.p_off * (NB * NB);
  if (occ == 4)
    caqr_panel_kernel<4><<<(unsigned)L.nstrips, 160, 0, st>>>(Vb, P.npad, row0, col0, L.nblk, L.bs, L.ntiles, L.s, li > 0, Tl, Vl, Pl, Asrc);
  else if (occ == 2)
    caqr_panel_kernel<2><<<(unsigned)L.nstrips, 160, 0, st>>>(Vb, P.npad, row0, col0, L.nblk, L.bs, L.ntiles, L.s, li > 0, Tl, Vl, Pl, Asrc);
  else
    caqr_panel_kernel<3><<<(unsigned)L.nstrips, 160, 0, st>>>(Vb, P.npad, row0, col0, L.nblk, L.bs, L.ntiles, L.s, li > 0, Tl, Vl, Pl, Asrc);
  PL_LAUNCH_CHECK();
  return 0;
}

// Factorisation driver.  Optional one-panel look-ahead on two streams (PL_LOOKAHEAD=1, read per call): panel p+1 is
// factored on a high-priority side stream while the main stream is still applying panel p to the columns right of it:
//   main:  U(p; columns of panel p+1)  -> E ->  U(p; remaining columns)              -> wait F -> next panel
//   side:                         wait E ->  panel kernels of p+1, all tree levels -> F
// (the panel kernels of the upper tree levels read only the pivot blocks written by the level below, so all levels of
// a panel can run back to back before any of its updates).  Measured at 8 M x 512: no gain (622.5 vs 621.6 ms) --
// two update CTAs fill the register file of an SM (2 x 256 x 128), so a panel CTA only ever replaces an update CTA
// instead of running beside it.  Off by default; kept for a future update kernel with a smaller footprint.
int caqr_factor(const Plan& P, double* Vb, double* Tws, double* Vup, double* Vpiv, cudaStream_t st, const double* Asrc) {
  static cudaStream_t ss_d[MAX_DEV] = {};
  static cudaEvent_t eE_d[MAX_DEV] = {}, eF_d[MAX_DEV] = {};
  const int dv = cur_dev();
  cudaStream_t& ss = ss_d[dv];
  cudaEvent_t &eE = eE_d[dv], &eF = eF_d[dv];
  const bool look = P.K > 1 && P.panels[0][0].ntiles >= 2048 && getenv("PL_LOOKAHEAD") != nullptr;
  if (look && !ss) {
    int lo = 0, hi = 0;
    PL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PL_CUDA(cudaStreamCreateWithPriority(&ss, cudaStreamNonBlocking, hi));
    PL_CUDA(cudaEventCreateWithFlags(&eE, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&eF, cudaEventDisableTiming));
  }
  int rc;
  if (Asrc && look) { set_error("caqr_factor: the fused input read is not available with PL_LOOKAHEAD"); return -1; }
  if (!look) {
    for (int p = 0; p < P.K; p++) {
      const int col0 = p * NB;
      const int ntrail = (int)((P.npad - col0 - NB) / NB);
      for (size_t li = 0; li < P.panels[p].size(); li++) {
        const Level& L = P.panels[p][li];
        // Asrc: the level-0 kernels of the FIRST panel read the caller's matrix and write Vb -- together they touch
        // every entry once, which replaces the separate copy pass over A
        const double* src = (p == 0 && li == 0) ? Asrc : nullptr;
        if ((rc = launch_panel(P, p, L, (int)li, Vb, Tws, Vup, Vpiv, st, src))) return rc;
        rc = launch_update(P, p, L, (int)li, Vb, Tws, Vup, Vpiv, 0, nullptr, 0, 0, 0, Vb, P.npad, col0 + NB, ntrail, 1, st, src);
        if (rc) return rc;
      }
    }
    return 0;
  }
  for (size_t li = 0; li < P.panels[0].size(); li++)
    if ((rc = launch_panel(P, 0, P.panels[0][li], (int)li, Vb, Tws, Vup, Vpiv, st))) return rc;
  for (int p = 0; p + 1 < P.K; p++) {
    const int col0 = p * NB;
    const int ntrail = (int)((P.npad - col0 - NB) / NB);
    const int nl = (int)P.panels[p].size();
    for (int li = 0; li < nl; li++)      // columns of the next panel first
      if ((rc = launch_update(P, p, P.panels[p][li], li, Vb, Tws, Vup, Vpiv, 0, nullptr, 0, 0, 0, Vb, P.npad, col0 + NB, 1, 1, st)))
        return rc;
    PL_CUDA(cudaEventRecord(eE, st));
    PL_CUDA(cudaStreamWaitEvent(ss, eE, 0));
    for (size_t li = 0; li < P.panels[p + 1].size(); li++)
      if ((rc = launch_panel(P, p + 1, P.panels[p + 1][li], (int)li, Vb, Tws, Vup, Vpiv, ss))) return rc;
    PL_CUDA(cudaEventRecord(eF, ss));
    for (int li = 0; li < nl; li++)
      if ((rc = launch_update(P, p, P.panels[p][li], li, Vb, Tws, Vup, Vpiv, 0, nullptr, 0, 0, 0, Vb, P.npad, col0 + 2 * NB, ntrail - 1, 1, st)))
        return rc;
    PL_CUDA(cudaStreamWaitEvent(st, eF, 0));
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
// dorgqr-style backward accumulation: for panel p = K-1 .. 0 the block reflectors are applied (levels top
// down, tiles in reverse order) to the trailing columns, then to the panel's own columns, whose input is
// the identity block on the panel's pivot rows and (virtually) zero elsewhere; the result overwrites the
// reflectors in place -- the pivot-block reflectors needed afterwards live in the Vpiv side store.
int caqr_form_q(const Plan& P, double* Vb, const double* Tws, const double* Vup, const double* Vpiv, cudaStream_t st) {
  {
    int64_t tot = P.npad * P.npad;
    ProfScope ps(PROF_MISC, st);
    zero_r_right_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(Vb, P.npad, (int)P.npad);
    PL_LAUNCH_CHECK();
  }
  for (int p = P.K - 1; p >= 0; p--) {
    const int64_t row0 = (int64_t)p * NB; const int col0 = p * NB;
    const int ntrail = (int)((P.npad - col0 - NB) / NB);
    const int nl = (int)P.panels[p].size();
    for (int li = nl - 1; li >= 0; li--) {
      int rc = launch_update(P, p, P.panels[p][li], li, Vb, Tws, Vup, Vpiv, 0, nullptr, 0, 0, 0, Vb, P.npad, col0 + NB, ntrail, 0, st);
      if (rc) return rc;
    }
    {
      ProfScope ps(PROF_MISC, st);
      set_identity_block_kernel<<<4, 256, 0, st>>>(Vb, P.npad, row0, col0);
      PL_LAUNCH_CHECK();
    }
    for (int li = nl - 1; li >= 0; li--) {
      int rc = launch_update(P, p, P.panels[p][li], li, Vb, Tws, Vup, Vpiv, 1, Vb, P.npad, col0, 1, nullptr, 0, 0, 0, 0, st);
      if (rc) return rc;
    }
  }
  return 0;
}

}  // namespace pl
