// Internal C++ interface between the translation units of libpylom_b200.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace pl {

struct Level {
  int64_t nblk;     // NB-row blocks at this level
  int64_t bs;       // row distance between consecutive blocks
  int64_t ntiles;   // tiles of G blocks
  int s;            // tiles per strip (flat tree inside a CTA)
  int64_t nstrips;
  int64_t t_off;    // tile offset of this level in the T store
  int64_t v_off;    // tile offset in the upper-level reflector store (-1 at level 0: in place)
  int64_t p_off;    // strip offset in the pivot-block reflector store (level 0 only, else -1)
};

struct Plan {
  int64_t m, n, npad, mrows;
  int K;                                   // panels
  std::vector<std::vector<Level>> panels;  // [panel][level]
  int64_t t_tiles, vup_tiles, vpiv_strips; // totals
};

Plan make_plan(int64_t m, int64_t n);
// Asrc (optional): read the input from the caller's row-major matrix with leading dimension P.npad instead of from Vb
// (needs n == npad and m % NB == 0); Vb then needs no copy of A, only its padding rows zeroed.
int caqr_factor(const Plan& P, double* Vb, double* Tws, double* Vup, double* Vpiv, cudaStream_t st, const double* Asrc = nullptr);
int caqr_extract_r(const Plan& P, const double* Vb, double* R, int64_t ldr, cudaStream_t st);
int caqr_form_q(const Plan& P, double* Vb, const double* Tws, const double* Vup, const double* Vpiv, cudaStream_t st);

// tsqr_small.cu : fused two-pass TSQR for n <= 64 (strips of a dense NP-row head + 512-row tall tiles, NP = 32 / 64)
constexpr int STB = 128;   // rows per warp in pass 1 / per slab in pass 2
constexpr int SRT = 512;   // tall tile rows
struct SmallPlan {
  int64_t m; int n, NP;
  int64_t ns;              // strips
  int64_t Tt, q, rem;      // whole tiles in total, per strip, strips with one more
  int64_t ntiles;          // Tt + (part != 0)
  int part;                // rows of the partial tile at the end of the last strip (0: none; < SRT)
  int tsz;                 // doubles of T per tile (1024 / 3072)
};
SmallPlan small_plan(int64_t m, int64_t n);
bool small_eligible(int64_t m, int64_t n);
// pass 1: A (lda, n columns, optional row centering -> mean) -> reflectors V (ldv >= NP; may alias the later output),
// per-tile T (ntiles * tsz), per-strip head T (ns * NP * NP), stacked strip triangles Rstack ((ns NP) x n, ldr)
int small_factor(const SmallPlan& P, const double* A, int64_t lda, double* V, int64_t ldv, double* Tst, double* Thst,
                 double* Rstack, int64_t ldr, double* mean, int center, cudaStream_t st);
// pass 2: U (m x nw, ldu) = Q [B_s; 0], B ((ns NP) x nw, ldb; destroyed).  U may alias V.
int small_apply(const SmallPlan& P, const double* V, int64_t ldv, const double* Tst, const double* Thst, double* B, int64_t ldb,
                double* U, int64_t ldu, int nw, cudaStream_t st);

// center.cu
int temporal_mean(double* out, const double* X, int64_t m, int64_t n, cudaStream_t st);
int subtract_mean(double* out, int64_t ldo, const double* X, const double* mean, int64_t m, int64_t n, int64_t pad_to, cudaStream_t st);
int center_rows(double* Y, int64_t ldy, double* mean, const double* X, int64_t m, int64_t n, int64_t pad_to, cudaStream_t st);
int temporal_variance(double* out, const double* X, const double* mean, int64_t m, int64_t n, cudaStream_t st);
int norm_variance(double* out, int64_t ldo, const double* X, const double* mean, const double* var, int64_t m, int64_t n, int64_t pad_to, cudaStream_t st);
int center_var_rows(double* Y, int64_t ldy, double* mean, double* var, const double* X, int64_t m, int64_t n, int64_t pad_to, cudaStream_t st);
int copy_pad(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t m, int64_t n, int64_t pad_to, cudaStream_t st);
int widen_f32(double* dst, const float* src, int64_t count, cudaStream_t st);
int narrow_f64(float* dst, const double* src, int64_t count, cudaStream_t st);
int complex_embed(double* Ah, const double* A, int64_t m, int64_t n, cudaStream_t st);
int complex_pack(double* Uc, const double* P, const double* Q, int64_t m, int64_t n, cudaStream_t st);
int vecmat(double* C, int64_t ldc, const double* v, const double* A, int64_t lda, int64_t m, int64_t n, cudaStream_t st);
int sumsq_diff(double* out2, double* scratch, const double* A, const double* B, int64_t cnt, cudaStream_t st);
constexpr int SUMSQ_SCRATCH_DOUBLES = 2 * 148 * 4;

// gemm.cu :  C[m x n] (ldc) = A[m x k] (lda) * Bp[kp x np] (ldb = np), Bp zero padded to kp%16==0, np%64==0
int gemm_tall(double* C, int64_t ldc, const double* A, int64_t lda, const double* Bp, int64_t ldb, int64_t m, int64_t n, int64_t k, cudaStream_t st);
// gemm_tn.cu : C (a x b, ldc) = X^T Y, X (m x a, ldx), Y (m x b, ldy); ws >= gemm_tn_workspace_bytes(a, b)
size_t gemm_tn_workspace_bytes(int64_t a, int64_t b);
int gemm_tn(double* C, int64_t ldc, const double* X, int64_t ldx, int64_t a, const double* Y, int64_t ldy, int64_t b, int64_t m,
            double* ws, cudaStream_t st);
int pad_small(double* dst, int64_t rows_p, int64_t cols_p, const double* src, int64_t lds, int64_t rows, int64_t cols, const double* rowscale, cudaStream_t st);

// svd_small.cu : R (n x n, ldr) = Ur diag(S) VT ; Ur, VT n x n with given ld; scratch >= 2*n*n + 4*n + 64 doubles
int svd_small(double* Ur, int64_t ldu, double* S, double* VT, int64_t ldvt, const double* R, int64_t ldr, int64_t n,
              double* scratch, int* sweeps_out, cudaStream_t st);
int64_t svd_small_scratch_doubles(int64_t n);

}  // namespace pl

// comm.cu : the NCCL communicator behind the collective entry points
struct pl_comm;
namespace pl {
int comm_allgather_inplace(pl_comm* c, double* buf, size_t count_per_rank, cudaStream_t st);   // rank r's block already at buf + r*count
int comm_side(pl_comm* c, cudaStream_t* side, cudaEvent_t* eR, cudaEvent_t* eS);
int comm_rank(const pl_comm* c);
int comm_size(const pl_comm* c);
int comm_scratch(pl_comm* c, size_t bytes, void** out);
}  // namespace pl
