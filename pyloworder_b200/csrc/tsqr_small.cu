// Fused two-pass Householder TSQR for SMALL column counts (n <= 64) on one GPU (sm_100a) -- BASELINE config 5
// (1e9 x 64 over 8 GPUs) is this shape.  Replaces, for these shapes, the reference's per-rank
// `dqr` = LAPACKE_dgeqrf + LAPACKE_dorgqr and the two back-multiplies `dmatmul` of dtsqr / dtsqr_svd
// (pyLOM/vmmath/src/svd.c:280-321, 673, 708): same mathematics (Householder reflectors, backward stable, Q orthonormal
// to rounding for any conditioning), executed as ~4 m n^2 flops instead of geqrf + orgqr + 2 GEMMs.
//
//   * the rows are cut into STRIPS (one CTA each, 2 per SM); a strip = a dense HEAD of NP rows (NP = 32 or 64, the
//     padded column count) followed by TALL TILES of 512 rows;
//   * pass 1 (small_factor_kernel): the head gets an ordinary Householder QR; then the strip's NP x NP triangle R stays
//     in shared memory while the tiles stream through: structured QR of [R; tile] with reflectors [e_j; v_j].
//     A tile is processed LEFT-LOOKING in sub-panels of 8 columns that live in REGISTERS (a lane holds one column of
//     32 rows, in the FP64 MMA fragment layout): the sub-panel is loaded, the tile's earlier reflectors (read back
//     from L2, where this CTA has just written them) are applied with their accumulated compact-WY T as DMMA
//     (mma.sync m8n8k4, the native DMMA.8x8x4), then 8 Householder column steps run CTA-wide -- lane-local FMA chains,
//     two shuffles, one block barrier per step, amortised over 512 rows -- and the 8 new reflectors are written to
//     the OUTPUT buffer.  The Gram block V_prev^T V_k (DMMA) extends T by 8 columns: T[:,k] = -T_prev (G T8);
//   * the strips' triangles are stacked and reduced by the generic CAQR path (a few thousand rows);
//   * pass 2 (small_apply_kernel): U = Q [B_s; 0] per strip, tiles in reverse order with the NP x NP block C carried in
//     shared memory:  X = T C (DMMA),  C -= X,  U_tile = -V X  -- the zero block of the target makes this ONE K = NP GEMM
//     per tile (2 m n^2 flops in total); the V fragments go from global memory straight into registers;
//   * small_head_apply_kernel finishes the head rows:  U_head = C - Y T (Y^T C).
//
// HBM traffic: read A, write V, read V, write U = 32 m n bytes (+ 5 % for T); the left-looking re-reads hit L2.
// The executable specification of exactly this decomposition is tests/model_tsqr_small.py.
#include "pl_common.cuh"
#include "caqr.h"
#include <cstdlib>

namespace pl {

constexpr int SNW = 8;             // warps per CTA in pass 1
constexpr int SRW = SRT / SNW;     // rows of the tall tile a warp owns (64)
constexpr int SRB = SRW / 8;       // 8-row blocks per warp

// =============================================================================================
// planner
// =============================================================================================
SmallPlan small_plan(int64_t m, int64_t n) {
  SmallPlan P;
  P.m = m; P.n = (int)n; P.NP = n <= 32 ? 32 : 64;
  // (environment read per call: the tests force many short strips on small inputs)
  const int64_t target = getenv("PL_SMALL_STRIPS") ? atoll(getenv("PL_SMALL_STRIPS")) : 296;
  const int64_t min_tiles = getenv("PL_SMALL_MIN_TILES") ? atoll(getenv("PL_SMALL_MIN_TILES")) : 2;
  int64_t ns = m / (P.NP + min_tiles * SRT);
  if (ns > target) ns = target;
  if (ns < 1) ns = 1;
  const int64_t body = m - ns * P.NP;
  P.ns = ns;
  P.Tt = body / SRT;
  P.part = (int)(body % SRT);
  P.q = P.Tt / ns;
  P.rem = P.Tt % ns;
  P.ntiles = P.Tt + (P.part ? 1 : 0);
  P.tsz = P.NP == 64 ? 3 * 1024 : 1024;
  return P;
}
bool small_eligible(int64_t m, int64_t n) {
  const bool off = getenv("PL_NO_SMALL") != nullptr;
  const int64_t min_rows = getenv("PL_SMALL_MIN_ROWS") ? atoll(getenv("PL_SMALL_MIN_ROWS")) : 32768;
  return !off && n >= 1 && n <= 64 && m >= min_rows;
}

struct StripGeom { int64_t r0; int64_t gt0; int nt; int last; };
__device__ __forceinline__ StripGeom strip_geom(const SmallPlan& P, int64_t i) {
  StripGeom g;
  const int64_t mn = i < P.rem ? i : P.rem;
  g.gt0 = i * P.q + mn;
  g.r0 = i * P.NP + (int64_t)SRT * g.gt0;
  g.nt = (int)(P.q + (i < P.rem ? 1 : 0));
  g.last = SRT;
  if (i == P.ns - 1 && P.part) { g.nt += 1; g.last = P.part; }
  return g;
}

// =============================================================================================
// device helpers
// =============================================================================================
// beta, tau, scale of the reflector that maps [alpha; x] (x^T x = sigma2) to [beta; 0]; same arithmetic as the generic
// panel kernel (caqr.cu): rsqrt / rcp seeds + two Newton steps, LAPACK formulas outside the safe range.
__device__ __forceinline__ void house_scalars(double alpha, double sigma2, double& beta, double& tau, double& scale) {
  beta = alpha; tau = 0.0; scale = 0.0;
  if (sigma2 != 0.0) {
    const double s2 = fma(alpha, alpha, sigma2);
    if (s2 > 1e-280 && s2 < 1e280) {
      double y, rc;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s2));
      const double aa = fabs(alpha);
      const double dd0 = fma(s2, y, aa);
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dd0));
      const double h = 0.5 * s2;
      y = y * fma(-h * y, y, 1.5);
      y = y * fma(-h * y, y, 1.5);
      const double nrm = s2 * y;
      const double dd = aa + nrm;
      rc = rc * fma(-dd, rc, 2.0);
      rc = rc * fma(-dd, rc, 2.0);
      beta = -copysign(nrm, alpha);
      scale = copysign(rc, alpha);
      tau = dd * y;
    } else {
      beta = -copysign(sqrt(s2), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
  }
}

// m8n8k4 FP64 MMA (one DMMA.8x8x4).  g = lane >> 2, t = lane & 3:
//   A (8x4,row): a = A[g][t];   B (4x8,col): b = B[t][g];   C/D (8x8): c[0],c[1] = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void mma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int NP>
struct SmallF {                        // shared memory of pass 1
  static constexpr int NSP = NP / 8;   // sub-panels
  static constexpr int LDA = NP + 4;
  static constexpr int RSZ = NP * (NP - 4 * NSP + 4);   // R packed by 8-row block rows: sum_k 8 (NP - 8k)
  static constexpr int WSZ = SNW * (NP - 8) * 8 + 2 * (NP - 8) * 8;   // Wpart + Ws + Wq
  static constexpr int HSZ = NP * LDA;                  // the head block H aliases Wpart/Ws/Wq (+ tail)
  double Tm[NP][LDA];                  // T of the running tile (head phase: the head's T)
  double Rp[RSZ];
  double Wb[WSZ > HSZ ? WSZ : HSZ];    // Wpart[SNW][NP-8][8] | Ws[NP-8][8] | Wq[NP-8][8];  head phase: H[NP][LDA]
  double red[2][SNW][8];
  double zt[8][8];                     // z (v_l^T v_j) of the running chain, column j; row 7 unused -> tau_j
  double tauv[8];
  double mean[SRT];                    // row means of the tile (centering); head phase: scratch
  double sc[8];
  __device__ double (*Wpart(int w))[8] { return reinterpret_cast<double (*)[8]>(Wb + (size_t)w * (NP - 8) * 8); }
  __device__ double (*Ws())[8] { return reinterpret_cast<double (*)[8]>(Wb + (size_t)SNW * (NP - 8) * 8); }
  __device__ double (*Wq())[8] { return reinterpret_cast<double (*)[8]>(Wb + (size_t)(SNW + 1) * (NP - 8) * 8); }
  __device__ double (*H())[LDA] { return reinterpret_cast<double (*)[LDA]>(Wb); }
};
template <int NP> __device__ __forceinline__ int roff(int k) { return 8 * k * (NP + 4 - 4 * k); }
template <int NP> __device__ __forceinline__ double& Rat(SmallF<NP>& S, int r, int c) {
  const int k = r >> 3;
  return S.Rp[roff<NP>(k) + (r & 7) * (NP - 8 * k) + (c - 8 * k)];
}

// ---- dense Householder QR of the NP x NP head in Hb[0:NP] (R above, unit-lower Y below), T (NP x NP) in Hb[NP:2NP] ----
template <int NP>
__device__ void head_factor(SmallF<NP>& S, int tid, int warp, int lane) {
  double (*H)[NP + 4] = S.H();
  double (*T)[NP + 4] = S.Tm;
  double* vcol = S.mean;                     // v_j (rows > j)
  double* wz = vcol + NP;                    // w_c (c > j) / z_c (c < j)
  for (int e = tid; e < NP * (NP + 4); e += 32 * SNW) (&T[0][0])[e] = 0.0;
  __syncthreads();
  for (int j = 0; j < NP; j++) {
    if (warp == 0) {
      double s = 0.0;
      for (int r = j + 1 + lane; r < NP; r += 32) { const double v = H[r][j]; s = fma(v, v, s); }
      s = warp_sum(s);
      double beta, tau, scale;
      house_scalars(H[j][j], s, beta, tau, scale);
      if (lane == 0) { S.sc[0] = beta; S.sc[1] = tau; S.sc[2] = scale; }
      for (int r = lane; r < NP; r += 32) vcol[r] = (r > j) ? H[r][j] * scale : 0.0;
    }
    __syncthreads();
    const double beta = S.sc[0], tau = S.sc[1];
    if (tid < NP && tid != j) {              // dot of v_j with column tid (rows > j) + the pivot-row entry
      double s0 = 0.0, s1 = 0.0;
      int r = j + 1;
      for (; r + 1 < NP; r += 2) { s0 = fma(vcol[r], H[r][tid], s0); s1 = fma(vcol[r + 1], H[r + 1][tid], s1); }
      if (r < NP) s0 = fma(vcol[r], H[r][tid], s0);
      const double dot = H[j][tid] + (s0 + s1);
      wz[tid] = (tid > j) ? tau * dot : dot;
    }
    __syncthreads();
    // trailing update (rows >= j, columns > j), reflector column j, T column j
    for (int e = tid; e < (NP - j) * NP; e += 32 * SNW) {
      const int r = j + e / NP, c = e % NP;
      if (c > j) H[r][c] -= ((r == j) ? 1.0 : vcol[r]) * wz[c];
      else if (c == j) H[r][c] = (r == j) ? beta : vcol[r];
    }
    if (tid < j) {
      double s = 0.0;
      for (int l = tid; l < j; l++) s = fma(T[tid][l], wz[l], s);
      T[tid][j] = -tau * s;
    } else if (tid == j) {
      T[j][j] = tau;
    }
    __syncthreads();
  }
}

// =============================================================================================
// pass 1
// =============================================================================================
// Per-phase cycle counters of thread 0 (build with -DPL_SMALL_TIMING; read with pl_debug_small_read):
// 0 load / means, 1 left-looking GEMM1, 2 its small products, 3 GEMM2, 4 chain, 5 store + Gram + T column, 6 head, 7 T store.
#ifdef PL_SMALL_TIMING
__device__ unsigned long long g_small_dbg[8];
extern "C" int pl_debug_small_read(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_small_dbg, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_small_dbg, z, sizeof(z));
  return 0;
}
#define ST_DECL long long stt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long st_prev = clock64();
#define ST_MARK(k) do { long long _t = clock64(); stt[k] += _t - st_prev; st_prev = _t; } while (0)
#define ST_FLUSH do { if (tid == 0) { for (int k = 0; k < 8; k++) atomicAdd(&g_small_dbg[k], (unsigned long long)stt[k]); } } while (0)
#else
#define ST_DECL
#define ST_MARK(k)
#define ST_FLUSH
#endif

struct SmallFArgs {
  SmallPlan P;
  const double* A; int64_t lda;        // input (n valid columns)
  double* V; int64_t ldv;              // reflector store (NP columns, ldv >= NP); may alias the input and the later output
  double* Tst;                         // ntiles * tsz
  double* Thst;                        // ns * NP * NP
  double* Rstack; int64_t ldr;         // (ns NP) x n
  double* mean; int center;
};

// Partial products of this warp's SRW rows:  P (8k x 8) = V[:, 0:8k]^T C,  C = the register sub-panel (lane (g, t) holds
// column g, rows 8 rb + 2 t + e).  The K index of an MMA runs over the rows (2 t + s), s fixed.  One 16-byte load feeds
// two MMAs: accumulator 2 pr + e holds the reflectors 16 pr + 2 g + e (a permutation of the M index).  The loads of a row
// block are issued ahead of its MMAs (the fragments come from L2; 16 warps per SM hide the rest of the latency).
template <int NSP>
__device__ __forceinline__ void lgemm1(double (*Wp)[8], const double (&c)[SRB][2], int k, const double* Vw, int64_t ldv,
                                       int rows_w, int g, int t) {
  constexpr int NPR = NSP / 2;
  double acc[2 * NPR][2];
#pragma unroll
  for (int q = 0; q < 2 * NPR; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
#pragma unroll
  for (int rb = 0; rb < SRB; rb++) {
    double2 av[2][NPR];
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int row = 8 * rb + 2 * t + s;
      const bool ok = row < rows_w;
      const double* vp = Vw + (int64_t)row * ldv + 2 * g;
#pragma unroll
      for (int pr = 0; pr < NPR; pr++)
        av[s][pr] = (ok && 2 * pr < k) ? *reinterpret_cast<const double2*>(vp + 16 * pr) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int pr = 0; pr < NPR; pr++) {
        if (2 * pr < k) {
          mma884(acc[2 * pr], av[s][pr].x, c[rb][s]);
          mma884(acc[2 * pr + 1], av[s][pr].y, c[rb][s]);
        }
      }
  }
#pragma unroll
  for (int pr = 0; pr < NPR; pr++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int refl = 16 * pr + 2 * g + e;
      if (refl < 8 * k) *reinterpret_cast<double2*>(&Wp[refl][2 * t]) = make_double2(acc[2 * pr + e][0], acc[2 * pr + e][1]);
    }
}

// C^T (8 cols x rows) -= W'^T (8 x 8k) V_prev^T (8k x rows): the accumulator fragment IS the register sub-panel.  K runs
// over the reflectors; one 16-byte load of V[row][8 pr + 2 t .. +1] feeds two MMAs (K order permuted accordingly).
template <int NSP>
__device__ __forceinline__ void lgemm2(double (&c)[SRB][2], const double (*Wq)[8], int k, const double* Vw, int64_t ldv,
                                       int rows_w, int g, int t) {
  double wa[NSP - 1][2];
#pragma unroll
  for (int pr = 0; pr < NSP - 1; pr++) {
    wa[pr][0] = (pr < k) ? -Wq[8 * pr + 2 * t][g] : 0.0;
    wa[pr][1] = (pr < k) ? -Wq[8 * pr + 2 * t + 1][g] : 0.0;
  }
#pragma unroll
  for (int rb0 = 0; rb0 < SRB; rb0 += 2) {
    double2 bv[2][NSP - 1];
#pragma unroll
    for (int ri = 0; ri < 2; ri++) {
      const int row = 8 * (rb0 + ri) + g;
      const bool ok = row < rows_w;
      const double* vp = Vw + (int64_t)row * ldv + 2 * t;
#pragma unroll
      for (int pr = 0; pr < NSP - 1; pr++)
        bv[ri][pr] = (ok && pr < k) ? *reinterpret_cast<const double2*>(vp + 8 * pr) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int pr = 0; pr < NSP - 1; pr++) {
      if (pr < k) {
#pragma unroll
        for (int ri = 0; ri < 2; ri++) mma884(c[rb0 + ri], wa[pr][0], bv[ri][pr].x);
#pragma unroll
        for (int ri = 0; ri < 2; ri++) mma884(c[rb0 + ri], wa[pr][1], bv[ri][pr].y);
      }
    }
  }
}

// sum of the SNW per-warp partial blocks at [r][cc]
template <int NP>
__device__ __forceinline__ double wpart_sum(SmallF<NP>& S, int r, int cc) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int w = 0; w < SNW; w += 2) { s0 += S.Wpart(w)[r][cc]; s1 += S.Wpart(w + 1)[r][cc]; }
  return s0 + s1;
}

template <int NP>
__global__ void __launch_bounds__(32 * SNW, 2) small_factor_kernel(SmallFArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallF<NP>& S = *reinterpret_cast<SmallF<NP>*>(smem_raw);
  constexpr int NSP = NP / 8, LDA = NP + 4, NT = 32 * SNW;
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n = A.P.n;
  const StripGeom sg = strip_geom(A.P, blockIdx.x);
  double* const V = A.V;                         // written and read back in this kernel: no __restrict__ / read-only path
  double (*Tm)[LDA] = S.Tm;
  ST_DECL

  // ---- head: rows [r0, r0 + NP)
  {
    double (*H)[LDA] = S.H();
    const double inv_n = 1.0 / (double)n;
    for (int r = warp; r < NP; r += SNW) {
      const double* src = A.A + (sg.r0 + r) * A.lda;
      double v0 = (lane < n) ? src[lane] : 0.0;
      double v1 = (NP == 64 && lane + 32 < n) ? src[lane + 32] : 0.0;
      if (A.center) {
        const double mu = warp_sum(v0 + v1) * inv_n;
        if (lane < n) v0 -= mu;
        if (NP == 64 && lane + 32 < n) v1 -= mu;
        if (lane == 0) A.mean[sg.r0 + r] = mu;
      }
      H[r][lane] = v0;
      if (NP == 64) H[r][lane + 32] = v1;
    }
    __syncthreads();
    head_factor<NP>(S, tid, warp, lane);
    for (int r = warp; r < NP; r += SNW) {               // Y (and R above it) -> reflector store, T -> head store
      double* dst = V + (sg.r0 + r) * A.ldv;
      dst[lane] = H[r][lane];
      if (NP == 64) dst[lane + 32] = H[r][lane + 32];
      double* tdst = A.Thst + (int64_t)blockIdx.x * (NP * NP) + r * NP;
      tdst[lane] = Tm[r][lane];
      if (NP == 64) tdst[lane + 32] = Tm[r][lane + 32];
    }
    for (int e = tid; e < NP * NP; e += NT) {            // R = triu(H) -> packed triangle
      const int r = e / NP, c = e % NP;
      if (c >= 8 * (r >> 3)) Rat<NP>(S, r, c) = (c >= r) ? H[r][c] : 0.0;
    }
    __syncthreads();
  }
  ST_MARK(6);

  // ---- tall tiles
  for (int tt = 0; tt < sg.nt; tt++) {
    const int rows = (tt == sg.nt - 1) ? sg.last : SRT;
    const int64_t t0 = sg.r0 + NP + (int64_t)tt * SRT;
    const int wrow0 = SRW * warp;                      // this warp's first row inside the tile
    const int rows_w = rows - wrow0;                   // its valid rows (<= 0: none)
    const double* Aw = A.A + (t0 + wrow0) * A.lda;
    double* Vw = V + (t0 + wrow0) * A.ldv;
    if (A.center) {                                    // row means of the warp's rows (this read also pulls the tile into L2)
      const double inv_n = 1.0 / (double)n;
#pragma unroll 4
      for (int r = 0; r < SRW; r++) {
        double mu = 0.0;
        if (r < rows_w) {
          const double* src = Aw + (int64_t)r * A.lda;
          const double v0 = (lane < n) ? src[lane] : 0.0;
          const double v1 = (NP == 64 && lane + 32 < n) ? src[lane + 32] : 0.0;
          mu = warp_sum(v0 + v1) * inv_n;
          if (lane == 0) A.mean[t0 + wrow0 + r] = mu;
        }
        if (lane == 0) S.mean[wrow0 + r] = mu;
      }
      __syncwarp();
    } else {                                           // pull the warp's rows into L2 (the sub-panel loads are 64-byte pieces)
      for (int r = lane; r < SRW && r < rows_w; r += 32) {
        const char* p = reinterpret_cast<const char*>(Aw + (int64_t)r * A.lda);
        for (int b = 0; b < n * 8; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
      }
    }
    ST_MARK(0);

    for (int k = 0; k < NSP; k++) {
      const int c0 = 8 * k, wk = NP - 8 * k;
      double* Rk = S.Rp + roff<NP>(k);                 // rows 8k..8k+7, columns 8k..NP-1:  Rk[jj * wk + c]
      // ---- (1) sub-panel -> registers: c[rb][e] = tile[8 rb + 2 t + e][8 k + g]
      double c[SRB][2];
      {
        const bool colok = c0 + g < n;
#pragma unroll
        for (int rb = 0; rb < SRB; rb++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int row = 8 * rb + 2 * t + e;
            double v = 0.0;
            if (colok && row < rows_w) {
              v = Aw[(int64_t)row * A.lda + c0 + g];
              if (A.center) v -= S.mean[wrow0 + row];
            }
            c[rb][e] = v;
          }
      }
      ST_MARK(0);
      // ---- (2) left-looking: apply the tile's reflectors 0 .. 8k-1 (block reflector with the accumulated T)
      if (k > 0) {
        lgemm1<NSP>(S.Wpart(warp), c, k, Vw, A.ldv, rows_w, g, t);
        __syncthreads();
        ST_MARK(1);
        for (int e = tid; e < 64 * k; e += NT) {       // W = sum of the partials + R[0:8k, Jk]
          const int r = e >> 3, cc = e & 7;
          S.Ws()[r][cc] = wpart_sum<NP>(S, r, cc) + Rat<NP>(S, r, c0 + cc);
        }
        __syncthreads();
        for (int e = tid; e < 64 * k; e += NT) {       // W' = T_prev^T W;  R[0:8k, Jk] -= W'
          const int r = e >> 3, cc = e & 7;
          double s = 0.0;
          for (int l = 0; l <= r; l++) s = fma(Tm[l][r], S.Ws()[l][cc], s);
          S.Wq()[r][cc] = s;
          Rat<NP>(S, r, c0 + cc) -= s;
        }
        __syncthreads();
        ST_MARK(2);
        lgemm2<NSP>(c, S.Wq(), k, Vw, A.ldv, rows_w, g, t);
        ST_MARK(3);
      }
      // ---- (3) the chain: 8 structured Householder steps on [R; C], CTA-wide, one barrier per step.  Reflector
      //      columns stay unscaled (u = x) during the chain and are scaled by 1/(alpha - beta) at the end.
      double mysc = 0.0;
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        const int src = (jj << 2) | t;
        double x[2 * SRB];
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
        for (int rb = 0; rb < SRB; rb += 2) {
          x[2 * rb] = __shfl_sync(FULL, c[rb][0], src);         x[2 * rb + 1] = __shfl_sync(FULL, c[rb][1], src);
          x[2 * rb + 2] = __shfl_sync(FULL, c[rb + 1][0], src); x[2 * rb + 3] = __shfl_sync(FULL, c[rb + 1][1], src);
          p0 = fma(x[2 * rb], c[rb][0], p0);         p1 = fma(x[2 * rb + 1], c[rb][1], p1);
          p2 = fma(x[2 * rb + 2], c[rb + 1][0], p2); p3 = fma(x[2 * rb + 3], c[rb + 1][1], p3);
        }
        double p = (p0 + p1) + (p2 + p3);
        p += __shfl_xor_sync(FULL, p, 1);
        p += __shfl_xor_sync(FULL, p, 2);
        if (t == 0) S.red[jj & 1][warp][g] = p;
        const double alpha = Rk[jj * wk + jj];                   // row jj of R is read BEFORE the barrier, written after it
        const double rjc = Rk[jj * wk + g];
        __syncthreads();
        double tot;
        {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int w = 0; w < SNW; w += 2) { s0 += S.red[jj & 1][w][g]; s1 += S.red[jj & 1][w + 1][g]; }
          tot = s0 + s1;
        }
        const double sigma2 = __shfl_sync(FULL, tot, jj << 2);
        double beta, tau, scale;
        house_scalars(alpha, sigma2, beta, tau, scale);
        const double w = (g > jj) ? tau * fma(scale, tot, rjc) : 0.0;
        if (warp == 0 && t == 0) {
          if (g > jj) Rk[jj * wk + g] = rjc - w;
          else if (g == jj) { Rk[jj * wk + jj] = beta; S.tauv[jj] = tau; }
          else S.zt[g][jj] = mysc * scale * tot;                 // z_l = v_l^T v_jj (the columns are still unscaled)
        }
        const double sw = scale * w;
        if (g == jj) mysc = scale;
#pragma unroll
        for (int rb = 0; rb < SRB; rb++) {
          c[rb][0] = fma(-x[2 * rb], sw, c[rb][0]);
          c[rb][1] = fma(-x[2 * rb + 1], sw, c[rb][1]);
        }
      }
      // ---- (4) scale, reflectors -> output buffer; T8 (compact-WY recurrence on the stored z, tau) -> diagonal block of T
#pragma unroll
      for (int rb = 0; rb < SRB; rb++) {
        c[rb][0] *= mysc; c[rb][1] *= mysc;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int row = 8 * rb + 2 * t + e;
          if (row < rows_w) Vw[(int64_t)row * A.ldv + c0 + g] = c[rb][e];
        }
      }
      if (warp == 0) {
        __syncwarp();
        if (lane < 8) {                                          // lane = row of T8
          double trow[8];
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < jj; l++) acc = fma(trow[l], S.zt[l][jj], acc);
            const double tj = S.tauv[jj];
            trow[jj] = (lane < jj) ? -tj * acc : ((lane == jj) ? tj : 0.0);
          }
#pragma unroll
          for (int q = 0; q < 8; q++) Tm[c0 + lane][c0 + q] = trow[q];
        }
      }
      ST_MARK(4);
      // ---- (5) T[0:8k, Jk] = -T_prev (G T8),  G = V_prev^T V_k
      if (k > 0) {
        lgemm1<NSP>(S.Wpart(warp), c, k, Vw, A.ldv, rows_w, g, t);
        __syncthreads();
        for (int e = tid; e < 64 * k; e += NT) {
          const int r = e >> 3, cc = e & 7;
          S.Ws()[r][cc] = wpart_sum<NP>(S, r, cc);
        }
        __syncthreads();
        for (int e = tid; e < 64 * k; e += NT) {       // G T8
          const int r = e >> 3, cc = e & 7;
          double s = 0.0;
          for (int l = 0; l <= cc; l++) s = fma(S.Ws()[r][l], Tm[c0 + l][c0 + cc], s);
          S.Wq()[r][cc] = s;
        }
        __syncthreads();
        for (int e = tid; e < 64 * k; e += NT) {
          const int r = e >> 3, cc = e & 7;
          double s = 0.0;
          for (int l = r; l < c0; l++) s = fma(Tm[r][l], S.Wq()[l][cc], s);
          Tm[r][c0 + cc] = -s;
        }
      }
      __syncthreads();
      ST_MARK(5);
    }
    {
      double* Tt = A.Tst + (sg.gt0 + tt) * (int64_t)A.P.tsz;
      for (int e = tid; e < A.P.tsz; e += NT) {
        const int blk = e >> 10, rr = (e >> 5) & 31, cc = e & 31;
        const int r = rr + (blk == 2 ? 32 : 0), cl = cc + (blk >= 1 ? 32 : 0);
        Tt[e] = (cl >= r) ? Tm[r][cl] : 0.0;
      }
    }
    __syncthreads();
    ST_MARK(7);
  }
  // ---- the strip's triangle -> stack
  for (int e = tid; e < NP * n; e += NT) {
    const int r = e / n, c = e % n;
    A.Rstack[((int64_t)blockIdx.x * NP + r) * A.ldr + c] = (c >= r) ? Rat<NP>(S, r, c) : 0.0;
  }
  ST_FLUSH;
}

// =============================================================================================
// pass 2
// =============================================================================================
template <int NP>
struct SmallA {
  static constexpr int LDC = NP + 4, LDX = NP + 2, NBLK = NP == 64 ? 3 : 1;
  double Cs[NP][LDC];                  // C[k][col]
  double Xt[NP][LDX];                  // X transposed: Xt[col][k]
  double Tb[NBLK][32][36];
};
struct SmallAArgs {
  SmallPlan P;
  const double* V; int64_t ldv;
  const double* Tst;
  double* B; int64_t ldb;              // (ns NP) x nw: in = Q_stack W, out = the block left for the head
  double* U; int64_t ldu; int nw;
};

template <int NP>
__global__ void __launch_bounds__(256, 2) small_apply_kernel(SmallAArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallA<NP>& S = *reinterpret_cast<SmallA<NP>*>(smem_raw);
  constexpr int NKB = NP / 16, NNB = NP / 8, NBLK = SmallA<NP>::NBLK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const StripGeom sg = strip_geom(A.P, blockIdx.x);
  const int nw = A.nw;
  double* Bs = A.B + (int64_t)blockIdx.x * NP * A.ldb;
  for (int e = tid; e < NP * NP; e += 256) {
    const int k = e / NP, c = e % NP;
    S.Cs[k][c] = (c < nw) ? Bs[(int64_t)k * A.ldb + c] : 0.0;
  }
  const bool vec_store = (A.ldu % 2 == 0) && ((reinterpret_cast<uintptr_t>(A.U) & 15) == 0);
  for (int tt = sg.nt - 1; tt >= 0; tt--) {
    const int rows = (tt == sg.nt - 1) ? sg.last : SRT;
    const int64_t t0 = sg.r0 + NP + (int64_t)tt * SRT;
    // T of this tile -> shared memory (cp.async)
    {
      const double* Tt = A.Tst + (sg.gt0 + tt) * (int64_t)A.P.tsz;
      for (int e = tid; e < NBLK * 512; e += 256) {      // 16-byte pieces
        const int blk = e >> 9, rr = (e >> 4) & 31, c2 = (e & 15) * 2;
        cp_async16(&S.Tb[blk][rr][c2], Tt + blk * 1024 + rr * 32 + c2, true);
      }
      cp_async_commit();
    }
    {   // the first 128-row slab of this tile -> L2
      const char* nx = reinterpret_cast<const char*>(A.V + (t0 + (tid >> 1)) * A.ldv) + (tid & 1) * (NP * 4);
      if ((tid >> 1) < rows)
        for (int b = 0; b < NP * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + b));
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- step A: X = T C  (upper block triangle of T only)
    {
      int mbs[2], nbs[2], nm, nn;
      if (NP == 64) { mbs[0] = (warp & 1) ? 1 : 0; mbs[1] = (warp & 1) ? 2 : 3; nm = 2; nbs[0] = 2 * (warp >> 1); nbs[1] = nbs[0] + 1; nn = 2; }
      else { mbs[0] = warp >> 2; mbs[1] = 0; nm = 1; nbs[0] = warp & 3; nbs[1] = 0; nn = 1; }
      for (int im = 0; im < nm; im++) {
        const int mb = mbs[im];
        for (int in = 0; in < nn; in++) {
          const int nb = nbs[in];
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          for (int kb = mb; kb < NKB; kb++) {
            double fa[8], fb[4];
#pragma unroll
            for (int x = 0; x < 8; x++) {
              const int r = 16 * mb + g + 8 * (x & 1), c = 16 * kb + t4 + 4 * (x >> 1);
              fa[x] = S.Tb[(r >> 5) + (c >> 5)][r & 31][c & 31];
            }
#pragma unroll
            for (int x = 0; x < 4; x++) fb[x] = S.Cs[16 * kb + t4 + 4 * x][8 * nb + g];
            mma16816(acc, fa, fb);
          }
          const int r = 16 * mb + g, c = 8 * nb + 2 * t4;
          S.Xt[c][r] = acc[0]; S.Xt[c + 1][r] = acc[1];
          S.Xt[c][r + 8] = acc[2]; S.Xt[c + 1][r + 8] = acc[3];
        }
      }
    }
    __syncthreads();
    // ---- C -= X
    for (int e = tid; e < NP * NP; e += 256) {
      const int k = e / NP, c = e % NP;
      S.Cs[k][c] -= S.Xt[c][k];
    }
    // ---- step C: U_tile = -V X, 128 rows at a time: warp -> rows 16 warp .. 16 warp + 15, K order (kb, 4 t4 + j)
    for (int sub = 0; sub * STB < rows; sub++) {
      const int rows_s = rows - sub * STB;               // valid rows from this 128-row slab on
      const int64_t s0 = t0 + (int64_t)sub * STB;
      double va[NKB][8];
      {
        const int r0 = 16 * warp + g, r1 = r0 + 8;
        const double* p0 = A.V + (s0 + r0) * A.ldv + 4 * t4;
        const double* p1 = A.V + (s0 + r1) * A.ldv + 4 * t4;
#pragma unroll
        for (int kb = 0; kb < NKB; kb++) {
          double2 x0 = make_double2(0.0, 0.0), x1 = x0, y0 = x0, y1 = x0;
          if (r0 < rows_s) { x0 = *reinterpret_cast<const double2*>(p0 + 16 * kb); x1 = *reinterpret_cast<const double2*>(p0 + 16 * kb + 2); }
          if (r1 < rows_s) { y0 = *reinterpret_cast<const double2*>(p1 + 16 * kb); y1 = *reinterpret_cast<const double2*>(p1 + 16 * kb + 2); }
          // a[2 j + h] = V[row g + 8 h][16 kb + 4 t4 + j]
          va[kb][0] = x0.x; va[kb][2] = x0.y; va[kb][4] = x1.x; va[kb][6] = x1.y;
          va[kb][1] = y0.x; va[kb][3] = y0.y; va[kb][5] = y1.x; va[kb][7] = y1.y;
        }
      }
      {   // the slab after this one (or the first slab of the next tile) -> L2
        int64_t nrow = s0 + STB + (tid >> 1);
        bool ok = (sub + 1) * STB + (tid >> 1) < rows;
        if ((sub + 1) * STB >= rows) { nrow = t0 - SRT + (tid >> 1); ok = tt > 0; }
        if (ok) {
          const char* nx = reinterpret_cast<const char*>(A.V + nrow * A.ldv) + (tid & 1) * (NP * 4);
          for (int b = 0; b < NP * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + b));
        }
      }
#pragma unroll
      for (int nh = 0; nh < NNB / 4; nh++) {
        double acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0; }
#pragma unroll
        for (int kb = 0; kb < NKB; kb++) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double* xp = &S.Xt[8 * (4 * nh + q) + g][16 * kb + 4 * t4];
            const double2 b01 = *reinterpret_cast<const double2*>(xp), b23 = *reinterpret_cast<const double2*>(xp + 2);
            const double fb[4] = {b01.x, b01.y, b23.x, b23.y};
            mma16816(acc[q], va[kb], fb);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int c = 8 * (4 * nh + q) + 2 * t4;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int r = 16 * warp + g + 8 * h;
            if (r < rows_s) {
              double* dst = A.U + (s0 + r) * A.ldu + c;
              if (vec_store && c + 1 < nw) *reinterpret_cast<double2*>(dst) = make_double2(-acc[q][2 * h], -acc[q][2 * h + 1]);
              else { if (c < nw) dst[0] = -acc[q][2 * h]; if (c + 1 < nw) dst[1] = -acc[q][2 * h + 1]; }
            }
          }
        }
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < NP * NP; e += 256) {          // what is left of C belongs to the head
    const int k = e / NP, c = e % NP;
    if (c < nw) Bs[(int64_t)k * A.ldb + c] = S.Cs[k][c];
  }
}

// U_head = C - Y T (Y^T C): one CTA per strip, everything in shared memory (NP <= 64: 3 x 33 KB).
template <int NP>
__global__ void __launch_bounds__(256) small_head_apply_kernel(SmallAArgs A, const double* __restrict__ Thst) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double (*Y)[NP + 1] = reinterpret_cast<double (*)[NP + 1]>(smem_raw);
  double (*C)[NP + 1] = Y + NP;
  double (*P)[NP + 1] = C + NP;
  const int tid = threadIdx.x;
  const StripGeom sg = strip_geom(A.P, blockIdx.x);
  const int nw = A.nw;
  const double* Bs = A.B + (int64_t)blockIdx.x * NP * A.ldb;
  const double* Th = Thst + (int64_t)blockIdx.x * (NP * NP);
  for (int e = tid; e < NP * NP; e += 256) {
    const int r = e / NP, c = e % NP;
    const double v = A.V[(sg.r0 + r) * A.ldv + c];
    Y[r][c] = (r > c) ? v : ((r == c) ? 1.0 : 0.0);
    C[r][c] = (c < nw) ? Bs[(int64_t)r * A.ldb + c] : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < NP * NP; e += 256) {          // P = Y^T C
    const int i = e / NP, c = e % NP;
    double s = 0.0;
    for (int r = i; r < NP; r++) s = fma(Y[r][i], C[r][c], s);
    P[i][c] = s;
  }
  __syncthreads();
  double z[(NP * NP) / 256];
#pragma unroll
  for (int q = 0; q < (NP * NP) / 256; q++) {         // Z = T P  (T upper triangular)
    const int e = tid + 256 * q, i = e / NP, c = e % NP;
    double s = 0.0;
    for (int l = i; l < NP; l++) s = fma(Th[i * NP + l], P[l][c], s);
    z[q] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < (NP * NP) / 256; q++) { const int e = tid + 256 * q; P[e / NP][e % NP] = z[q]; }
  __syncthreads();
  for (int e = tid; e < NP * NP; e += 256) {          // U_head = C - Y Z
    const int r = e / NP, c = e % NP;
    if (c >= nw) continue;
    double s = C[r][c];
    for (int l = 0; l <= r; l++) s = fma(-Y[r][l], P[l][c], s);
    A.U[(sg.r0 + r) * A.ldu + c] = s;
  }
}

// =============================================================================================
// drivers
// =============================================================================================
template <int NP>
static int small_factor_t(const SmallFArgs& A, cudaStream_t st) {
  static DevOnce attr;
  if (first_on_device(attr))
    PL_CUDA(cudaFuncSetAttribute(small_factor_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallF<NP>)));
  ProfScope ps(PROF_SMALL, st);
  small_factor_kernel<NP><<<(unsigned)A.P.ns, 32 * SNW, sizeof(SmallF<NP>), st>>>(A);
  PL_LAUNCH_CHECK();
  return 0;
}
int small_factor(const SmallPlan& P, const double* A, int64_t lda, double* V, int64_t ldv, double* Tst, double* Thst,
                 double* Rstack, int64_t ldr, double* mean, int center, cudaStream_t st) {
  SmallFArgs a;
  a.P = P; a.A = A; a.lda = lda; a.V = V; a.ldv = ldv; a.Tst = Tst; a.Thst = Thst; a.Rstack = Rstack; a.ldr = ldr;
  a.mean = mean; a.center = center;
  return P.NP == 64 ? small_factor_t<64>(a, st) : small_factor_t<32>(a, st);
}

template <int NP>
static int small_apply_t(const SmallAArgs& A, const double* Thst, cudaStream_t st) {
  static DevOnce attr;
  constexpr int head_smem = 3 * NP * (NP + 1) * 8;
  if (first_on_device(attr)) {
    PL_CUDA(cudaFuncSetAttribute(small_apply_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallA<NP>)));
    PL_CUDA(cudaFuncSetAttribute(small_head_apply_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, head_smem));
  }
  ProfScope ps(PROF_SMALL, st);
  small_apply_kernel<NP><<<(unsigned)A.P.ns, 256, sizeof(SmallA<NP>), st>>>(A);
  PL_LAUNCH_CHECK();
  small_head_apply_kernel<NP><<<(unsigned)A.P.ns, 256, head_smem, st>>>(A, Thst);
  PL_LAUNCH_CHECK();
  return 0;
}
int small_apply(const SmallPlan& P, const double* V, int64_t ldv, const double* Tst, const double* Thst, double* B, int64_t ldb,
                double* U, int64_t ldu, int nw, cudaStream_t st) {
  SmallAArgs a;
  a.P = P; a.V = V; a.ldv = ldv; a.Tst = Tst; a.B = B; a.ldb = ldb; a.U = U; a.ldu = ldu; a.nw = nw;
  return P.NP == 64 ? small_apply_t<64>(a, Thst, st) : small_apply_t<32>(a, Thst, st);
}

}  // namespace pl
