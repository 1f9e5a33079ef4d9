// C ABI of libpylom_b200 (declared in include/pylom_b200.h) -- thin composition of the kernels.
#include "pl_common.cuh"
#include "caqr.h"
#include "../../include/pylom_b200.h"
#include <atomic>
#include <memory>
#include <cstring>
#include <mutex>
#include <vector>
#include <cstdlib>
#include <thread>

namespace pl {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
const char* last_error() { return g_err; }
void count_launches(long long k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

// ---- profiling ------------------------------------------------------------------------------
struct ProfRec { int cls; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
bool prof_enabled() { return g_prof_on; }
void prof_begin(int cls, cudaStream_t st) {
  ProfRec r; r.cls = cls; cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); cudaEventRecord(r.e0, st); g_prof.push_back(r);
}
void prof_end(cudaStream_t st) { cudaEventRecord(g_prof.back().e1, st); }

// ---- workspace layout for one tall matrix ---------------------------------------------------
struct WsLayout {
  Plan plan;
  size_t vb, tws, vup, vpiv, r, bp, ur, svd, vt, s, tmp, total;
  bool ext_vb; int64_t tmp_rows;
  // small-n path (tsqr_small.cu): reflector store (vb, ld NP), per-tile / per-head T, stacked strip triangles, the
  // block that seeds pass 2, and the generic layout of the (ns NP) x n stack
  bool small = false;
  SmallPlan sp;
  size_t s_tst = 0, s_thst = 0, s_rst = 0, s_bst = 0, s_inner = 0;
  std::shared_ptr<WsLayout> inner;
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static WsLayout make_layout(int64_t m, int64_t n, bool ext_vb = false, bool allow_small = true) {
  WsLayout L;
  if (allow_small && small_eligible(m, n)) {
    L.small = true;
    L.sp = small_plan(m, n);
    const SmallPlan& S = L.sp;
    L.plan.m = m; L.plan.n = n; L.plan.npad = S.NP; L.plan.K = 0; L.plan.mrows = m;
    L.plan.t_tiles = L.plan.vup_tiles = L.plan.vpiv_strips = 0;
    L.ext_vb = ext_vb; L.tmp = 0; L.tmp_rows = 0; L.tws = L.vup = L.vpiv = 0;
    size_t off = 0;
    L.vb = off;     if (!ext_vb) off += al((size_t)m * S.NP * 8);
    L.s_tst = off;  off += al((size_t)S.ntiles * S.tsz * 8);
    L.s_thst = off; off += al((size_t)S.ns * S.NP * S.NP * 8);
    L.s_rst = off;  off += al((size_t)S.ns * S.NP * n * 8);
    L.s_bst = off;  off += al((size_t)S.ns * S.NP * n * 8);
    L.r = off;      off += al((size_t)n * n * 8);
    L.bp = off;
    L.ur = off;     off += al((size_t)n * n * 8);
    L.vt = off;     off += al((size_t)n * n * 8);
    L.s = off;      off += al((size_t)n * 8);
    L.svd = off;    off += al((size_t)svd_small_scratch_doubles(n) * 8);
    L.inner = std::make_shared<WsLayout>(make_layout(S.ns * S.NP, n, false, false));
    L.s_inner = off; off += L.inner->total;
    L.total = off;
    return L;
  }
  L.plan = make_plan(m, n);
  const Plan& P = L.plan;
  size_t off = 0;
  L.ext_vb = ext_vb; L.tmp = 0; L.tmp_rows = 0;
  L.vb = off;   if (!ext_vb) off += al((size_t)P.mrows * P.npad * 8);
  L.tws = off;  off += al((size_t)P.t_tiles * NB * NB * 8);
  L.vup = off;  off += al((size_t)(P.vup_tiles > 0 ? P.vup_tiles : 1) * TB * NB * 8);
  L.vpiv = off; off += al((size_t)P.vpiv_strips * NB * NB * 8);
  L.r = off;    off += al((size_t)n * n * 8);
  const int64_t kp = round_up(n, 16), np = round_up(n, 64);
  L.bp = off;   off += al((size_t)kp * np * 8);
  L.ur = off;   off += al((size_t)n * n * 8);
  L.vt = off;   off += al((size_t)n * n * 8);
  L.s = off;    off += al((size_t)n * 8);
  L.svd = off;  off += al((size_t)svd_small_scratch_doubles(n) * 8);
  if (ext_vb) {   // row-chunk buffer of the in-place back-multiply (<= 2 GiB)
    int64_t rows = (int64_t)(2147483648LL / 8) / P.npad;
    rows = rows < 4096 ? 4096 : rows; if (rows > m) rows = m;
    L.tmp_rows = rows; L.tmp = off; off += al((size_t)rows * P.npad * 8);
  }
  L.total = off;
  return L;
}
static inline double* at(void* ws, size_t off) { return reinterpret_cast<double*>(static_cast<char*>(ws) + off); }

static int check_ws(const WsLayout& L, void* ws, size_t ws_bytes, int argpos) {
  if (!ws || ws_bytes < L.total) { set_error("workspace too small: need %zu bytes, got %zu", L.total, ws_bytes); return -argpos; }
  if (reinterpret_cast<uintptr_t>(ws) & 255) { set_error("workspace must be 256-byte aligned"); return -argpos; }
  return 0;
}

static int qr_factor(double* R, double* X_mean, const double* A, int64_t m, int64_t n, int center, void* ws,
                     const WsLayout& L, cudaStream_t st, double* X_var = nullptr, double* Vb_ext = nullptr) {
  const Plan& P = L.plan;
  double* Vb = Vb_ext ? Vb_ext : at(ws, L.vb);
  int rc;
  if (L.small) {   // n <= 64: fused tile TSQR (pass 1) -> stacked strip triangles -> generic CAQR of the small stack
    const SmallPlan& S = L.sp;
    const double* src = A; int64_t lda = n; int cen = center == 1;
    if (center == 2) {   // variance normalisation: materialise (A - mean) / var in the reflector store, factor it in place
      ProfScope ps(PROF_COPY, st);
      if ((rc = center_var_rows(Vb, S.NP, X_mean, X_var, A, m, n, S.NP, st))) return rc;
      src = Vb; lda = S.NP; cen = 0;
    }
    double* Rst = at(ws, L.s_rst);
    if ((rc = small_factor(S, src, lda, Vb, S.NP, at(ws, L.s_tst), at(ws, L.s_thst), Rst, n, X_mean, cen, st))) return rc;
    return qr_factor(R, nullptr, Rst, S.ns * S.NP, n, 0, static_cast<char*>(ws) + L.s_inner, *L.inner, st);
  }
  // No centering, n a multiple of the panel width, whole row blocks: the first panel's kernels read A directly and the
  // copy pass (8 + 8 bytes per entry) disappears.  PL_NO_FUSED_INPUT / PL_LOOKAHEAD keep the copy.
  const bool fused = !center && P.npad == n && (m % NB) == 0 && P.K > 1 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                     getenv("PL_NO_FUSED_INPUT") == nullptr && getenv("PL_LOOKAHEAD") == nullptr;
  {
    ProfScope ps(PROF_COPY, st);
    if (center == 2) rc = center_var_rows(Vb, P.npad, X_mean, X_var, A, m, n, P.npad, st);
    else if (center) rc = center_rows(Vb, P.npad, X_mean, A, m, n, P.npad, st);
    else if (!fused) rc = copy_pad(Vb, P.npad, A, n, m, n, P.npad, st);
    else rc = 0;
    if (rc) return rc;
    PL_CUDA(cudaMemsetAsync(Vb + (size_t)m * P.npad, 0, (size_t)(P.mrows - m) * P.npad * 8, st));
  }
  rc = caqr_factor(P, Vb, at(ws, L.tws), at(ws, L.vup), at(ws, L.vpiv), st, fused ? A : nullptr);
  if (rc) return rc;
  if (R) rc = caqr_extract_r(P, Vb, R, n, st);
  return rc;
}

static int qr_apply_q(double* U, int64_t ldu, const double* W, int64_t ldw, int64_t nw, int64_t m, int64_t n, int flags,
                      void* ws, const WsLayout& L, cudaStream_t st, double* Vb_ext = nullptr) {
  const Plan& P = L.plan;
  double* Vb = Vb_ext ? Vb_ext : at(ws, L.vb);
  int rc;
  if (L.small) {   // pass 2 seeded with Q_stack W (explicit Q_stack when W == NULL)
    const SmallPlan& S = L.sp;
    if (!W && nw != n) { set_error("apply_q: W == NULL needs nw == n"); return -5; }
    if (nw > n) { set_error("apply_q: nw > n"); return -5; }
    double* Bst = at(ws, L.s_bst);
    if ((rc = qr_apply_q(Bst, nw, W, ldw, nw, S.ns * S.NP, n, flags, static_cast<char*>(ws) + L.s_inner, *L.inner, st))) return rc;
    if (flags & 2) return 0;
    return small_apply(S, Vb, S.NP, at(ws, L.s_tst), at(ws, L.s_thst), Bst, nw, U, ldu, (int)nw, st);
  }
  if (!(flags & 1)) {
    rc = caqr_form_q(P, Vb, at(ws, L.tws), at(ws, L.vup), at(ws, L.vpiv), st);
    if (rc) return rc;
  }
  if (flags & 2) return 0;      // form Q1 only; a later call with bit 0 multiplies
  if (!W) {
    if (nw != n) { set_error("apply_q: W == NULL needs nw == n"); return -5; }
    if (U == Vb) return 0;   // explicit Q already in place
    return copy_pad(U, ldu, Vb, P.npad, m, n, n, st);
  }
  if (nw > n) { set_error("apply_q: nw > n"); return -5; }
  const int64_t kp = round_up(n, 16), np = round_up(nw, 64);
  double* Bp = at(ws, L.bp);
  rc = pad_small(Bp, kp, np, W, ldw, n, nw, nullptr, st);
  if (rc) return rc;
  ProfScope ps(PROF_GEMM, st);
  // k is rounded up to the packed height: the extra columns of Q1 (finite padding) meet zero rows of Bp,
  // and an even k keeps the 16-byte cp.async path for odd n
  if (U != Vb) return gemm_tall(U, ldu, Vb, P.npad, Bp, np, m, nw, kp, st);
  // in place (U aliases the factorisation buffer): row chunks through a bounded temporary
  if (!L.ext_vb || ldu != P.npad || nw != n) { set_error("apply_q: in-place product needs an external Vb with ld == n_pad == n"); return -1; }
  double* tmp = at(ws, L.tmp);
  for (int64_t r0 = 0; r0 < m; r0 += L.tmp_rows) {
    const int64_t rows = (m - r0 < L.tmp_rows) ? (m - r0) : L.tmp_rows;
    rc = gemm_tall(tmp, P.npad, Vb + r0 * P.npad, P.npad, Bp, np, rows, nw, kp, st);
    if (rc) return rc;
    PL_CUDA(cudaMemcpyAsync(Vb + r0 * P.npad, tmp, (size_t)rows * P.npad * 8, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}
}  // namespace pl

using namespace pl;

#define PL_ARG(cond, pos, msg) do { if (!(cond)) { set_error("bad argument %d: %s", pos, msg); return -(pos); } } while (0)

extern "C" {

static bool svd_overlap_enabled(int64_t m);
static int svd_and_apply_overlapped(double* Ui, double* S, double* VT, const double* R, double* Ur, int64_t m, int64_t n,
                                    void* ws, const WsLayout& L, cudaStream_t st, double* Vb_ext);

int pl_version(void) { return 100; }
const char* pl_last_error(void) { return last_error(); }
int64_t pl_launch_count(void) { return (int64_t)g_launches.load(); }
void pl_profile_enable(int on) { g_prof_on = on != 0; }
int pl_profile_read(double* ms_by_class, int64_t* launches_by_class, int ncls) {
  for (int i = 0; i < ncls; i++) { ms_by_class[i] = 0.0; launches_by_class[i] = 0; }
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.cls < ncls) { ms_by_class[r.cls] += ms; launches_by_class[r.cls] += 1; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return 0;
}

int pl_temporal_mean_f64(double* out, const double* X, int64_t m, int64_t n, void* stream) {
  PL_ARG(m >= 0 && n > 0, 3, "m >= 0, n > 0");
  return temporal_mean(out, X, m, n, (cudaStream_t)stream);
}
int pl_subtract_mean_f64(double* out, const double* X, const double* X_mean, int64_t m, int64_t n, void* stream) {
  PL_ARG(m >= 0 && n > 0, 4, "m >= 0, n > 0");
  return subtract_mean(out, n, X, X_mean, m, n, n, (cudaStream_t)stream);
}
int pl_center_f64(double* Y, double* X_mean, const double* X, int64_t m, int64_t n, void* stream) {
  PL_ARG(m >= 0 && n > 0, 4, "m >= 0, n > 0");
  return center_rows(Y, n, X_mean, X, m, n, n, (cudaStream_t)stream);
}
int pl_temporal_variance_f64(double* out, const double* X, const double* X_mean, int64_t m, int64_t n, void* stream) {
  PL_ARG(m >= 0 && n > 0, 4, "m >= 0, n > 0");
  return temporal_variance(out, X, X_mean, m, n, (cudaStream_t)stream);
}
int pl_norm_variance_f64(double* out, const double* X, const double* X_mean, const double* X_var, int64_t m, int64_t n, void* stream) {
  PL_ARG(m >= 0 && n > 0, 5, "m >= 0, n > 0");
  return norm_variance(out, n, X, X_mean, X_var, m, n, n, (cudaStream_t)stream);
}
size_t pl_matmul_tn_workspace_bytes(int64_t a, int64_t b) { return (a > 0 && b > 0) ? gemm_tn_workspace_bytes(a, b) : 256; }
int pl_matmul_tn_f64(double* C, int64_t ldc, const double* X, int64_t ldx, int64_t a, const double* Y, int64_t ldy, int64_t b,
                     int64_t m, void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(a >= 0 && b >= 0 && m >= 0, 5, "negative size");
  if (a == 0 || b == 0) return 0;
  PL_ARG(C && ldc >= b, 2, "C / ldc");
  PL_ARG(m == 0 || (X && ldx >= a), 4, "X / ldx");
  PL_ARG(m == 0 || (Y && ldy >= b), 7, "Y / ldy");
  if (!ws || ws_bytes < gemm_tn_workspace_bytes(a, b) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("matmul_tn: workspace too small or misaligned (need %zu bytes)", gemm_tn_workspace_bytes(a, b)); return -10;
  }
  return gemm_tn(C, ldc, X, ldx, a, Y, ldy, b, m, static_cast<double*>(ws), (cudaStream_t)stream);
}
int pl_widen_f32_f64(double* dst, const float* src, int64_t count, void* stream) {
  if (!dst || !src || count < 0) { set_error("pl_widen_f32_f64: bad arguments"); return -1; }
  return widen_f32(dst, src, count, (cudaStream_t)stream);
}
int pl_narrow_f64_f32(float* dst, const double* src, int64_t count, void* stream) {
  if (!dst || !src || count < 0) { set_error("pl_narrow_f64_f32: bad arguments"); return -1; }
  return narrow_f64(dst, src, count, (cudaStream_t)stream);
}
int pl_complex_embed_f64(double* Ahat, const double* A, int64_t m, int64_t n, void* stream) {
  if (!Ahat || !A || m < 0 || n < 0) { set_error("pl_complex_embed_f64: bad arguments"); return -1; }
  return complex_embed(Ahat, A, m, n, (cudaStream_t)stream);
}
int pl_complex_pack_f64(double* Uc, const double* P, const double* Q, int64_t m, int64_t n, void* stream) {
  if (!Uc || !P || m < 0 || n < 0) { set_error("pl_complex_pack_f64: bad arguments"); return -1; }
  return complex_pack(Uc, P, Q, m, n, (cudaStream_t)stream);
}
int pl_vecmat_f64(double* C, const double* v, const double* A, int64_t m, int64_t n, void* stream) {
  return vecmat(C, n, v, A, n, m, n, (cudaStream_t)stream);
}

size_t pl_matmul_workspace_bytes(int64_t n, int64_t k) { return (size_t)round_up(k, 16) * round_up(n, 64) * 8 + 256; }
int pl_matmul_f64(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m,
                  int64_t n, int64_t k, void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(m >= 0 && n > 0 && k > 0, 7, "sizes");
  PL_ARG(ws && ws_bytes >= pl_matmul_workspace_bytes(n, k), 10, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t kp = round_up(k, 16), np = round_up(n, 64);
  int rc = pad_small((double*)ws, kp, np, B, ldb, k, n, nullptr, st);
  if (rc) return rc;
  return gemm_tall(C, ldc, A, lda, (double*)ws, np, m, n, k, st);
}

size_t pl_rmse_workspace_bytes(void) { return SUMSQ_SCRATCH_DOUBLES * 8; }
int pl_rmse_sums_f64(double* out2, const double* A, const double* B, int64_t count, void* ws, void* stream) {
  return sumsq_diff(out2, (double*)ws, A, B, count, (cudaStream_t)stream);
}

size_t pl_qr_workspace_bytes(int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  return make_layout(m, n).total;
}

int pl_qr_factor_f64(double* R, double* X_mean, const double* A, int64_t m, int64_t n, int center, void* ws,
                     size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n, 4, "need m >= n > 0 (the reference has the same precondition, svd.py:69)");
  PL_ARG(!center || X_mean, 2, "X_mean required when center != 0");
  WsLayout L = make_layout(m, n);
  int rc = check_ws(L, ws, ws_bytes, 7);
  if (rc) return rc;
  return qr_factor(R, X_mean, A, m, n, center, ws, L, (cudaStream_t)stream);
}

int pl_qr_factor_var_f64(double* R, double* X_mean, double* X_var, const double* A, int64_t m, int64_t n, void* ws,
                         size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n, 5, "need m >= n > 0");
  PL_ARG(X_mean && X_var, 2, "X_mean and X_var required");
  WsLayout L = make_layout(m, n);
  int rc = check_ws(L, ws, ws_bytes, 7);
  if (rc) return rc;
  return qr_factor(R, X_mean, A, m, n, 2, ws, L, (cudaStream_t)stream, X_var);
}

int pl_qr_apply_q_f64(double* U, int64_t ldu, const double* W, int64_t ldw, int64_t nw, int64_t m, int64_t n, int flags,
                      void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n, 6, "need m >= n > 0");
  WsLayout L = make_layout(m, n);
  int rc = check_ws(L, ws, ws_bytes, 9);
  if (rc) return rc;
  return qr_apply_q(U, ldu, W, ldw, nw, m, n, flags, ws, L, (cudaStream_t)stream);
}

/* ---- in-place variants: the caller's U buffer ((m + n + 32) x n doubles, n % 32 == 0) is the factorisation buffer */
int64_t pl_qr_inplace_rows(int64_t m, int64_t n) { return (n > 0 && n % NB == 0) ? m + n + NB : 0; }
size_t pl_qr_workspace_bytes_inplace(int64_t m, int64_t n) {
  if (m <= 0 || n <= 0 || n % NB) return 0;
  return make_layout(m, n, true).total;
}
int pl_qr_factor_inplace_f64(double* R, double* X_mean, double* Ubuf, const double* A, int64_t m, int64_t n, int center,
                             void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n && n % NB == 0, 5, "need m >= n > 0 and n % 32 == 0");
  PL_ARG(!center || X_mean, 2, "X_mean required when center != 0");
  WsLayout L = make_layout(m, n, true);
  int rc = check_ws(L, ws, ws_bytes, 8);
  if (rc) return rc;
  return qr_factor(R, X_mean, A, m, n, center, ws, L, (cudaStream_t)stream, nullptr, Ubuf);
}
int pl_qr_apply_q_inplace_f64(double* Ubuf, const double* W, int64_t ldw, int64_t m, int64_t n, int flags, void* ws,
                              size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n && n % NB == 0, 5, "need m >= n > 0 and n % 32 == 0");
  WsLayout L = make_layout(m, n, true);
  int rc = check_ws(L, ws, ws_bytes, 7);
  if (rc) return rc;
  return qr_apply_q(Ubuf, n, W, ldw, n, m, n, flags, ws, L, (cudaStream_t)stream, Ubuf);
}
int pl_pod_run_inplace_f64(double* Ubuf, double* S, double* VT, double* X_mean, const double* X, int64_t m, int64_t n,
                           int remove_mean, void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n && n % NB == 0, 7, "need m >= n > 0 and n % 32 == 0");
  PL_ARG(!remove_mean || X_mean, 4, "X_mean required when remove_mean != 0");
  cudaStream_t st = (cudaStream_t)stream;
  WsLayout L = make_layout(m, n, true);
  int rc = check_ws(L, ws, ws_bytes, 9);
  if (rc) return rc;
  double* R = at(ws, L.r);
  rc = qr_factor(R, X_mean, X, m, n, remove_mean ? 1 : 0, ws, L, st, nullptr, Ubuf);
  if (rc) return rc;
  double* Ur = at(ws, L.ur);
  if (!L.small && svd_overlap_enabled(m)) return svd_and_apply_overlapped(Ubuf, S, VT, R, Ur, m, n, ws, L, st, Ubuf);
  {
    ProfScope ps(PROF_SVD, st);
    rc = svd_small(Ur, n, S, VT, n, R, n, n, at(ws, L.svd), nullptr, st);
  }
  if (rc) return rc;
  return qr_apply_q(Ubuf, n, Ur, n, n, m, n, 0, ws, L, st, Ubuf);
}

size_t pl_svd_workspace_bytes(int64_t n) { return (size_t)svd_small_scratch_doubles(n) * 8 + 256; }
int pl_svd_f64(double* U, double* S, double* VT, const double* Y, int64_t n, void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(n > 0, 5, "n > 0");
  PL_ARG(ws && ws_bytes >= pl_svd_workspace_bytes(n), 6, "workspace too small");
  return svd_small(U, n, S, VT, n, Y, n, n, (double*)ws, nullptr, (cudaStream_t)stream);
}


// The Jacobi SVD of R (32 latency-bound CTAs) runs on a high-priority side stream while the main stream forms the
// explicit Q (which needs the reflectors only); the back-multiply waits for both.  Worth ~1.5 % at 8 M x 512.
static bool svd_overlap_enabled(int64_t m) {
  static const bool off = getenv("PL_NO_SVD_OVERLAP") != nullptr;
  static const int64_t min_rows = getenv("PL_SVD_OVERLAP_MIN_ROWS") ? atoll(getenv("PL_SVD_OVERLAP_MIN_ROWS")) : 1000000;
  return !off && m >= min_rows;
}
static int svd_and_apply_overlapped(double* Ui, double* S, double* VT, const double* R, double* Ur, int64_t m, int64_t n,
                                    void* ws, const WsLayout& L, cudaStream_t st, double* Vb_ext) {
  static cudaStream_t ss_d[MAX_DEV] = {};
  static cudaEvent_t eR_d[MAX_DEV] = {}, eS_d[MAX_DEV] = {};
  const int dv = cur_dev();
  cudaStream_t& ss = ss_d[dv];
  cudaEvent_t &eR = eR_d[dv], &eS = eS_d[dv];
  if (!ss) {
    int lo = 0, hi = 0;
    PL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PL_CUDA(cudaStreamCreateWithPriority(&ss, cudaStreamNonBlocking, hi));
    PL_CUDA(cudaEventCreateWithFlags(&eR, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&eS, cudaEventDisableTiming));
  }
  double* Vb = Vb_ext ? Vb_ext : at(ws, L.vb);
  PL_CUDA(cudaEventRecord(eR, st));
  int rc = caqr_form_q(L.plan, Vb, at(ws, L.tws), at(ws, L.vup), at(ws, L.vpiv), st);   // enqueued, asynchronous
  if (rc) return rc;
  PL_CUDA(cudaStreamWaitEvent(ss, eR, 0));
  {
    ProfScope ps(PROF_SVD, ss);
    rc = svd_small(Ur, n, S, VT, n, R, n, n, at(ws, L.svd), nullptr, ss);
  }
  if (rc) return rc;
  PL_CUDA(cudaEventRecord(eS, ss));
  PL_CUDA(cudaStreamWaitEvent(st, eS, 0));
  return qr_apply_q(Ui, n, Ur, n, n, m, n, 1, ws, L, st, Vb_ext);      // flag 1: Q already formed
}

static int tsqr_svd_impl(double* Ui, double* S, double* VT, double* X_mean, const double* Ai, int64_t m, int64_t n,
                         int center, void* ws, size_t ws_bytes, cudaStream_t st) {
  WsLayout L = make_layout(m, n);
  int rc = check_ws(L, ws, ws_bytes, 9);
  if (rc) return rc;
  double* R = at(ws, L.r);
  // small-n path with n == NP: the reflectors are written straight into the output (no A-sized buffer at all)
  double* vext = (L.small && n == L.sp.NP) ? Ui : nullptr;
  rc = qr_factor(R, X_mean, Ai, m, n, center, ws, L, st, nullptr, vext);
  if (rc) return rc;
  double* Ur = at(ws, L.ur);
  if (!L.small && svd_overlap_enabled(m)) return svd_and_apply_overlapped(Ui, S, VT, R, Ur, m, n, ws, L, st, nullptr);
  {
    ProfScope ps(PROF_SVD, st);
    rc = svd_small(Ur, n, S, VT, n, R, n, n, at(ws, L.svd), nullptr, st);
  }
  if (rc) return rc;
  return qr_apply_q(Ui, n, Ur, n, n, m, n, 0, ws, L, st, vext);
}

}  // extern "C" (the collective composition below is C++ linkage, called from comm.cu)

// ---- P ranks: local factor -> ncclAllGather of R (in place in the gather buffer) -> stack QR + Jacobi -> apply ----------
namespace pl {
struct DistLayout { WsLayout L, Ls; size_t o_rst, o_w, o_stack, total; };
static DistLayout make_dist_layout(int64_t m, int64_t n, int P, bool inplace) {
  DistLayout D;
  D.L = make_layout(m, n, inplace);
  D.Ls = make_layout((int64_t)P * n, n);
  size_t off = D.L.total;
  D.o_rst = off;   off += al((size_t)P * n * n * 8);
  D.o_w = off;     off += al((size_t)P * n * n * 8);
  D.o_stack = off; off += D.Ls.total;
  D.total = off;
  return D;
}
size_t dist_ws_bytes(int64_t m, int64_t n, int P, int flags) {
  if ((flags & 1) && (n % NB)) return 0;
  return make_dist_layout(m, n, P, (flags & 1) != 0).total;
}
int dist_tsqr_svd(pl_comm* c, double* Ui, double* S, double* VT, double* X_mean, const double* Ai, int64_t m, int64_t n,
                  int center, int flags, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int P = comm_size(c), rank = comm_rank(c);
  const bool inplace = (flags & 1) != 0;
  if (inplace && (n % NB)) { set_error("in-place variant needs n %% 32 == 0"); return -10; }
  DistLayout D = make_dist_layout(m, n, P, inplace);
  if (!ws || ws_bytes < D.total) { set_error("workspace too small: need %zu bytes, got %zu", D.total, ws_bytes); return -11; }
  if (reinterpret_cast<uintptr_t>(ws) & 255) { set_error("workspace must be 256-byte aligned"); return -11; }
  double* Rst = at(ws, D.o_rst);
  double* Wst = at(ws, D.o_w);
  void* ws_stack = static_cast<char*>(ws) + D.o_stack;
  double* Vb_ext = (inplace || (D.L.small && n == D.L.sp.NP)) ? Ui : nullptr;   // small-n path: reflectors go straight into Ui
  // local Householder QR; R_i lands directly in this rank's slot of the gather buffer
  int rc = qr_factor(Rst + (size_t)rank * n * n, X_mean, Ai, m, n, center, ws, D.L, st, nullptr, Vb_ext);
  if (rc) return rc;
  const bool overlap = svd_overlap_enabled(m);
  cudaStream_t wk = st, side = nullptr;
  cudaEvent_t eR = nullptr, eS = nullptr;
  if (overlap) {
    if ((rc = comm_side(c, &side, &eR, &eS))) return rc;
    PL_CUDA(cudaEventRecord(eR, st));
    if ((rc = qr_apply_q(nullptr, n, nullptr, 0, n, m, n, 2, ws, D.L, st, Vb_ext))) return rc;   // explicit Q1 on the main stream
    PL_CUDA(cudaStreamWaitEvent(side, eR, 0));
    wk = side;
  }
  if ((rc = comm_allgather_inplace(c, Rst, (size_t)n * n, wk))) return rc;
  if ((rc = tsqr_svd_impl(Wst, S, VT, nullptr, Rst, (int64_t)P * n, n, 0, ws_stack, D.Ls.total, wk))) return rc;   // Wst = Q2 Ur
  if (overlap) {
    PL_CUDA(cudaEventRecord(eS, side));
    PL_CUDA(cudaStreamWaitEvent(st, eS, 0));
  }
  return qr_apply_q(Ui, n, Wst + (size_t)rank * n * n, n, n, m, n, overlap ? 1 : 0, ws, D.L, st, Vb_ext);
}
}  // namespace pl

extern "C" {

int pl_tsqr_svd_f64(double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n, void* ws, size_t ws_bytes,
                    void* stream) {
  PL_ARG(n > 0 && m >= n, 5, "need m >= n > 0");
  return tsqr_svd_impl(Ui, S, VT, nullptr, Ai, m, n, 0, ws, ws_bytes, (cudaStream_t)stream);
}

int pl_pod_run_f64(double* U, double* S, double* VT, double* X_mean, const double* X, int64_t m, int64_t n, int remove_mean,
                   void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(n > 0 && m >= n, 6, "need m >= n > 0");
  PL_ARG(!remove_mean || X_mean, 4, "X_mean required when remove_mean != 0");
  return tsqr_svd_impl(U, S, VT, X_mean, X, m, n, remove_mean ? 1 : 0, ws, ws_bytes, (cudaStream_t)stream);
}

int pl_reconstruct_f64(double* X, const double* U, int64_t ldu, const double* S, const double* VT, int64_t ldvt, int64_t m,
                       int64_t N, int64_t n, void* ws, size_t ws_bytes, void* stream) {
  PL_ARG(m >= 0 && N > 0 && n > 0, 7, "sizes");
  PL_ARG(ws && ws_bytes >= pl_matmul_workspace_bytes(n, N), 10, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t kp = round_up(N, 16), np = round_up(n, 64);
  int rc = pad_small((double*)ws, kp, np, VT, ldvt, N, n, S, st);   // diag(S) V folded into the packing
  if (rc) return rc;
  return gemm_tall(X, n, U, ldu, (double*)ws, np, m, n, N, st);
}

// Device buffers of the host-pointer entry point are cached across calls (grow-only): a multi-GB cudaMalloc /
// cudaFree pair per call costs several hundred milliseconds.  pl_host_cache_free() releases them.
constexpr int HC_SLOTS = 6;
static struct HostCache { void* p[HC_SLOTS] = {}; size_t cap[HC_SLOTS] = {}; } g_hc;
static int hc_get(int slot, size_t bytes, void** out) {
  if (g_hc.cap[slot] < bytes) {
    if (g_hc.p[slot]) cudaFree(g_hc.p[slot]);
    g_hc.p[slot] = nullptr; g_hc.cap[slot] = 0;
    cudaError_t e = cudaMalloc(&g_hc.p[slot], bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return 1000 + (int)e; }
    g_hc.cap[slot] = bytes;
  }
  *out = g_hc.p[slot];
  return 0;
}
// ---- pageable host memory ------------------------------------------------------------------------------------------
// A numpy caller (what the reference's Cython binding passes) hands over PAGEABLE memory; cudaMemcpyAsync then goes
// through the driver's single-threaded bounce buffer at ~12 GB/s (measured: 2 M x 512 took as long as 8 M x 512 from
// pinned buffers).  For such pointers the pipeline stages the rows itself: a ring of pinned 64 MB slots (allocated
// once, cached like the device buffers), filled / drained by a few host threads while the DMA of the previous slot
// and the GPU work of the previous chunk are in flight.
constexpr int PG_SLOTS = 4;
constexpr size_t PG_BYTES = (size_t)64 << 20;
static struct PageRing {
  void* p[PG_SLOTS] = {};
  cudaEvent_t ev[PG_SLOTS] = {};
  int next = 0;
} g_pg;
static int pg_init() {
  for (int i = 0; i < PG_SLOTS; i++) {
    if (!g_pg.p[i]) PL_CUDA(cudaHostAlloc(&g_pg.p[i], PG_BYTES, cudaHostAllocDefault));
    if (!g_pg.ev[i]) PL_CUDA(cudaEventCreateWithFlags(&g_pg.ev[i], cudaEventDisableTiming));
  }
  return 0;
}
static void pg_free() {
  for (int i = 0; i < PG_SLOTS; i++) {
    if (g_pg.p[i]) cudaFreeHost(g_pg.p[i]);
    if (g_pg.ev[i]) cudaEventDestroy(g_pg.ev[i]);
    g_pg.p[i] = nullptr; g_pg.ev[i] = nullptr;
  }
}
static bool is_pageable(const void* ptr) {
  if (getenv("PL_HOST_NO_STAGING")) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}
static void par_memcpy(void* dst, const void* src, size_t bytes) {
  static const int nt = []() { unsigned h = std::thread::hardware_concurrency(); int t = h ? (int)h : 4; if (const char* e = getenv("PL_HOST_COPY_THREADS")) t = atoi(e); return t < 1 ? 1 : (t > 16 ? 16 : t); }();
  if (bytes < ((size_t)4 << 20) || nt == 1) { memcpy(dst, src, bytes); return; }
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (int i = 1; i < nt; i++) {
    const size_t o = (size_t)i * per;
    if (o >= bytes) break;
    const size_t len = (o + per > bytes) ? bytes - o : per;
    th.emplace_back([=]() { memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, len); });
  }
  memcpy(dst, src, per < bytes ? per : bytes);
  for (auto& t : th) t.join();
}
// pageable host -> device, enqueued on `cs` piece by piece through the pinned ring (blocks the calling thread only for the
// host-side copies; the DMA of a piece overlaps the host copy of the next one)
static int pg_h2d(void* dst, const void* src, size_t bytes, cudaStream_t cs) {
  int rc = pg_init();
  if (rc) return rc;
  for (size_t o = 0; o < bytes; o += PG_BYTES) {
    const size_t len = bytes - o < PG_BYTES ? bytes - o : PG_BYTES;
    const int k = g_pg.next; g_pg.next = (k + 1) % PG_SLOTS;
    PL_CUDA(cudaEventSynchronize(g_pg.ev[k]));                 // the slot's previous DMA is done
    par_memcpy(g_pg.p[k], static_cast<const char*>(src) + o, len);
    PL_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + o, g_pg.p[k], len, cudaMemcpyHostToDevice, cs));
    PL_CUDA(cudaEventRecord(g_pg.ev[k], cs));
  }
  return 0;
}
// device -> pageable host: the DMA of piece i+1 runs while piece i is copied out of its pinned slot; returns when the data
// is in `dst` (the caller has already ordered `cs` behind the producer of `src`)
static int pg_d2h(void* dst, const void* src, size_t bytes, cudaStream_t cs) {
  int rc = pg_init();
  if (rc) return rc;
  int pend_k = -1; size_t pend_o = 0, pend_len = 0;
  for (size_t o = 0; o < bytes; o += PG_BYTES) {
    const size_t len = bytes - o < PG_BYTES ? bytes - o : PG_BYTES;
    const int k = g_pg.next; g_pg.next = (k + 1) % PG_SLOTS;
    PL_CUDA(cudaEventSynchronize(g_pg.ev[k]));
    PL_CUDA(cudaMemcpyAsync(g_pg.p[k], static_cast<const char*>(src) + o, len, cudaMemcpyDeviceToHost, cs));
    PL_CUDA(cudaEventRecord(g_pg.ev[k], cs));
    if (pend_k >= 0) {
      PL_CUDA(cudaEventSynchronize(g_pg.ev[pend_k]));
      par_memcpy(static_cast<char*>(dst) + pend_o, g_pg.p[pend_k], pend_len);
    }
    pend_k = k; pend_o = o; pend_len = len;
  }
  if (pend_k >= 0) {
    PL_CUDA(cudaEventSynchronize(g_pg.ev[pend_k]));
    par_memcpy(static_cast<char*>(dst) + pend_o, g_pg.p[pend_k], pend_len);
  }
  return 0;
}

void pl_host_cache_free(void) {
  for (int i = 0; i < HC_SLOTS; i++) { if (g_hc.p[i]) cudaFree(g_hc.p[i]); g_hc.p[i] = nullptr; g_hc.cap[i] = 0; }
  pg_free();
}

// Row-chunk count of the host pipeline: ~2 GiB of snapshots per chunk, at least 4 chunks above 512 MiB, and every
// chunk at least 4n rows tall (PL_HOST_CHUNKS overrides).
static int host_chunks(int64_t m, int64_t n) {
  const double bytes = (double)m * n * 8;
  int64_t c = (int64_t)ceil(bytes / 2147483648.0);
  if (bytes > 536870912.0 && c < 4) c = 4;
  if (const char* e = getenv("PL_HOST_CHUNKS")) c = atoi(e);
  if (c > 64) c = 64;
  while (c > 1 && m / c < 4 * n) c--;
  return c < 1 ? 1 : (int)c;
}

// Row partition of the host pipeline: C chunks, the first C-1 of `mc` rows (whole 128-row tiles), the last one takes
// the remainder and is never shorter than 4n rows (it absorbs a short tail).
static int host_chunk_plan(int64_t m, int64_t n, int64_t* mc_out) {
  int C = host_chunks(m, n);
  int64_t mc = round_up((m + C - 1) / C, TB);
  C = (int)((m + mc - 1) / mc);
  if (C > 1 && m - (int64_t)(C - 1) * mc < 4 * n) C--;
  if (C <= 1) { C = 1; mc = m; }
  *mc_out = mc;
  return C;
}
// exported for the host-logic tests (pure host arithmetic, no device access): rows of every chunk
int pl_host_chunk_rows(int64_t m, int64_t n, int64_t* rows, int max_chunks) {
  if (n <= 0 || m < n) return -1;
  int64_t mc = 0;
  const int C = host_chunk_plan(m, n, &mc);
  for (int c = 0; c < C && c < max_chunks; c++) rows[c] = (c == C - 1) ? m - (int64_t)c * mc : mc;
  return C;
}

// Host-pointer TSQR-SVD (replaces dtsqr_svd, pyLOM/vmmath/src/svd.c:678-712).  The rows are processed as C chunks,
// i.e. as a two-level TSQR on one device, so that PCIe and the GPU work at the same time:
//   factor:  H2D(c+1)   ||  factor(c), R_c, explicit Q_c             (copy stream / compute stream)
//            QR of the stacked R_c -> this rank's R                   (side stream, beside the last chunk's Q_c)
//   [single rank: Jacobi SVD of R on the side stream; P ranks: the caller exchanges the R's and provides W]
//   apply:   B = Q_stack W;  U_c = Q_c B_c (GEMM)  ||  D2H(c-1)       (ping-pong output buffers)
// The state between the two phases (chunk plans, device buffers) lives in g_hs.
struct HostChunk { int64_t r0, rows; Plan P; size_t vb, tws, vup, vpiv; };
static struct HostState {
  bool valid = false;
  int64_t m = 0, n = 0, kp = 0, np = 0, m2 = 0;
  int C = 0;
  size_t chunk_out = 0, o_rs = 0, o_bs = 0, o_r2 = 0, o_ur = 0, o_bp = 0, o_svd = 0, o_ws2 = 0, o_w = 0;
  std::vector<HostChunk> ch;
  WsLayout L2;
  void *dVb = nullptr, *dOut = nullptr, *dS = nullptr, *dV = nullptr, *aux = nullptr;
  cudaStream_t st = nullptr, cs = nullptr, s2 = nullptr;   // compute / copy / small-factor streams
  cudaEvent_t eR = nullptr, eB = nullptr;
} g_hs;

static int host_streams() {
  HostState& H = g_hs;
  if (H.st) return 0;
  int lo = 0, hi = 0;
  PL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  PL_CUDA(cudaStreamCreateWithFlags(&H.st, cudaStreamNonBlocking));
  PL_CUDA(cudaStreamCreateWithFlags(&H.cs, cudaStreamNonBlocking));
  PL_CUDA(cudaStreamCreateWithPriority(&H.s2, cudaStreamNonBlocking, hi));
  PL_CUDA(cudaEventCreateWithFlags(&H.eR, cudaEventDisableTiming));
  PL_CUDA(cudaEventCreateWithFlags(&H.eB, cudaEventDisableTiming));
  return 0;
}

struct EventList {
  std::vector<cudaEvent_t> e;
  int init(int k) {
    e.resize(k);
    for (auto& x : e) PL_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    return 0;
  }
  ~EventList() { for (auto x : e) if (x) cudaEventDestroy(x); }
};

// Phase 1.  Leaves this rank's R (n x n) in the device buffer at o_r2, enqueued on stream s2 (not synchronised).
static int host_factor(const double* Ai, int64_t m, int64_t n) {
  HostState& H = g_hs;
  H.valid = false;
  int64_t mc = 0;
  const int C = host_chunk_plan(m, n, &mc);
  const int64_t npad = round_up(n, NB);
  const bool direct = (npad == n);                       // rows land in the factorisation buffer as they are
  H.m = m; H.n = n; H.C = C;
  H.ch.assign(C, HostChunk());
  size_t vb_bytes = 0, aux_bytes = 0;
  int64_t max_rows = 0;
  for (int c = 0; c < C; c++) {
    HostChunk& k = H.ch[c];
    k.r0 = (int64_t)c * mc; k.rows = (c == C - 1) ? m - k.r0 : mc;
    if (k.rows < n) { set_error("host pipeline: chunk %d has %lld rows < n", c, (long long)k.rows); return -5; }
    k.P = make_plan(k.rows, n);
    k.vb = vb_bytes;   vb_bytes += al((size_t)k.P.mrows * k.P.npad * 8);
    k.tws = aux_bytes;  aux_bytes += al((size_t)k.P.t_tiles * NB * NB * 8);
    k.vup = aux_bytes;  aux_bytes += al((size_t)(k.P.vup_tiles > 0 ? k.P.vup_tiles : 1) * TB * NB * 8);
    k.vpiv = aux_bytes; aux_bytes += al((size_t)k.P.vpiv_strips * NB * NB * 8);
    if (k.rows > max_rows) max_rows = k.rows;
  }
  // small buffers: stacked R (C n x n), B = Q_stack W (C n x n), R2, Ur, W, packed B chunk, Jacobi scratch, level-2 workspace
  H.kp = round_up(n, 16); H.np = round_up(n, 64); H.m2 = (int64_t)C * n;
  H.L2 = make_layout(H.m2, n);
  size_t off = aux_bytes;
  H.o_rs = off;  off += al((size_t)H.m2 * n * 8);
  H.o_bs = off;  off += al((size_t)H.m2 * n * 8);
  H.o_r2 = off;  off += al((size_t)n * n * 8);
  H.o_ur = off;  off += al((size_t)n * n * 8);
  H.o_w = off;   off += al((size_t)n * n * 8);
  H.o_bp = off;  off += al((size_t)H.kp * H.np * 8);
  H.o_svd = off; off += al((size_t)svd_small_scratch_doubles(n) * 8);
  H.o_ws2 = off; off += C > 1 ? H.L2.total : 0;
  H.chunk_out = al((size_t)max_rows * n * 8);
  void* stage = nullptr;
  int rc;
  if ((rc = hc_get(0, vb_bytes, &H.dVb)) || (rc = hc_get(1, 2 * H.chunk_out, &H.dOut)) || (rc = hc_get(2, (size_t)n * 8, &H.dS)) ||
      (rc = hc_get(3, (size_t)n * n * 8, &H.dV)) || (rc = hc_get(4, off, &H.aux)) ||
      (!direct && (rc = hc_get(5, 2 * H.chunk_out, &stage)))) { pl_host_cache_free(); return rc; }
  if ((rc = host_streams())) return rc;
  EventList ev, evs;
  if ((rc = ev.init(C)) || (rc = evs.init(2))) return rc;
  cudaStream_t st = H.st, cs = H.cs, s2 = H.s2;
  const bool pageable_in = is_pageable(Ai);
  double* Rs = at(H.aux, H.o_rs);
  double* R2 = at(H.aux, H.o_r2);
  // chunks arrive, get factored and turned into explicit Q_c while the next chunk is on the wire
  for (int c = 0; c < C; c++) {
    HostChunk& k = H.ch[c];
    double* Vb = at(H.dVb, k.vb);
    const size_t bytes = (size_t)k.rows * n * 8;
    if (direct) {
      if (pageable_in) { if ((rc = pg_h2d(Vb, Ai + k.r0 * n, bytes, cs))) return rc; }
      else PL_CUDA(cudaMemcpyAsync(Vb, Ai + k.r0 * n, bytes, cudaMemcpyHostToDevice, cs));
      PL_CUDA(cudaEventRecord(ev.e[c], cs));
      PL_CUDA(cudaStreamWaitEvent(st, ev.e[c], 0));
    } else {
      double* sg = reinterpret_cast<double*>(static_cast<char*>(stage) + (size_t)(c & 1) * H.chunk_out);
      if (c >= 2) PL_CUDA(cudaStreamWaitEvent(cs, evs.e[c & 1], 0));         // staging buffer consumed
      if (pageable_in) { if ((rc = pg_h2d(sg, Ai + k.r0 * n, bytes, cs))) return rc; }
      else PL_CUDA(cudaMemcpyAsync(sg, Ai + k.r0 * n, bytes, cudaMemcpyHostToDevice, cs));
      PL_CUDA(cudaEventRecord(ev.e[c], cs));
      PL_CUDA(cudaStreamWaitEvent(st, ev.e[c], 0));
      if ((rc = copy_pad(Vb, k.P.npad, sg, n, k.rows, n, k.P.npad, st))) return rc;
      PL_CUDA(cudaEventRecord(evs.e[c & 1], st));
    }
    PL_CUDA(cudaMemsetAsync(Vb + (size_t)k.rows * k.P.npad, 0, (size_t)(k.P.mrows - k.rows) * k.P.npad * 8, st));
    if ((rc = caqr_factor(k.P, Vb, at(H.aux, k.tws), at(H.aux, k.vup), at(H.aux, k.vpiv), st))) return rc;
    if ((rc = caqr_extract_r(k.P, Vb, C > 1 ? Rs + (size_t)c * n * n : R2, n, st))) return rc;
    if (c == C - 1) PL_CUDA(cudaEventRecord(H.eR, st));
    if ((rc = caqr_form_q(k.P, Vb, at(H.aux, k.tws), at(H.aux, k.vup), at(H.aux, k.vpiv), st))) return rc;
  }
  // side stream, beside the last chunk's Q formation: QR of the stacked R_c
  PL_CUDA(cudaStreamWaitEvent(s2, H.eR, 0));
  if (C > 1) {
    void* ws2 = static_cast<char*>(H.aux) + H.o_ws2;
    if ((rc = qr_factor(R2, nullptr, Rs, H.m2, n, 0, ws2, H.L2, s2))) return rc;
  }
  H.valid = true;      // (events still pending are released by the runtime when they complete)
  return 0;
}

// Phase 2.  Wd: device n x n matrix, ready on stream s2.  Ui (host) = Q_local Wd.
static int host_apply(double* Ui, const double* Wd) {
  HostState& H = g_hs;
  if (!H.valid) { set_error("host pipeline: apply without a preceding factor call"); return -1; }
  H.valid = false;
  const int64_t n = H.n;
  const int C = H.C;
  cudaStream_t st = H.st, cs = H.cs, s2 = H.s2;
  int rc;
  EventList ev, evd;
  if ((rc = ev.init(C)) || (rc = evd.init(2))) return rc;
  const bool pageable_out = is_pageable(Ui);
  const double* B = Wd;
  if (C > 1) {
    void* ws2 = static_cast<char*>(H.aux) + H.o_ws2;
    double* Bs = at(H.aux, H.o_bs);
    if ((rc = qr_apply_q(Bs, n, Wd, n, n, H.m2, n, 0, ws2, H.L2, s2))) return rc;
    B = Bs;
  }
  PL_CUDA(cudaEventRecord(H.eB, s2));
  PL_CUDA(cudaStreamWaitEvent(st, H.eB, 0));
  double* Bp = at(H.aux, H.o_bp);
  // U_c = Q_c B_c; the D2H of a chunk overlaps the GEMM of the next one
  for (int c = 0; c < C; c++) {
    HostChunk& k = H.ch[c];
    double* out = reinterpret_cast<double*>(static_cast<char*>(H.dOut) + (size_t)(c & 1) * H.chunk_out);
    if (c >= 2) PL_CUDA(cudaStreamWaitEvent(st, evd.e[c & 1], 0));           // output buffer drained
    if ((rc = pad_small(Bp, H.kp, H.np, B + (size_t)c * n * n, n, n, n, nullptr, st))) return rc;
    if ((rc = gemm_tall(out, n, at(H.dVb, k.vb), k.P.npad, Bp, H.np, k.rows, n, H.kp, st))) return rc;
    PL_CUDA(cudaEventRecord(ev.e[c], st));
    if (!pageable_out) {
      PL_CUDA(cudaStreamWaitEvent(cs, ev.e[c], 0));
      PL_CUDA(cudaMemcpyAsync(Ui + k.r0 * n, out, (size_t)k.rows * n * 8, cudaMemcpyDeviceToHost, cs));
      PL_CUDA(cudaEventRecord(evd.e[c & 1], cs));
    } else {
      // pageable destination: the previous chunk is drained through the pinned ring while this chunk's GEMM runs
      // (the copy stream is still ordered behind the previous chunk's GEMM only)
      if (c >= 1) {
        HostChunk& kp = H.ch[c - 1];
        double* outp = reinterpret_cast<double*>(static_cast<char*>(H.dOut) + (size_t)((c - 1) & 1) * H.chunk_out);
        if ((rc = pg_d2h(Ui + kp.r0 * n, outp, (size_t)kp.rows * n * 8, cs))) return rc;
        PL_CUDA(cudaEventRecord(evd.e[(c - 1) & 1], cs));
      }
      PL_CUDA(cudaStreamWaitEvent(cs, ev.e[c], 0));
      if (c == C - 1) {
        if ((rc = pg_d2h(Ui + k.r0 * n, out, (size_t)k.rows * n * 8, cs))) return rc;
        PL_CUDA(cudaEventRecord(evd.e[c & 1], cs));
      }
    }
  }
  PL_CUDA(cudaStreamSynchronize(s2));
  PL_CUDA(cudaStreamSynchronize(st));
  PL_CUDA(cudaStreamSynchronize(cs));
  return 0;
}

int pl_tsqr_svd_host_f64(double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n) {
  PL_ARG(n > 0 && m >= n, 5, "need m >= n > 0");
  int rc = host_factor(Ai, m, n);
  if (rc) return rc;
  HostState& H = g_hs;
  double* Ur = at(H.aux, H.o_ur);
  if ((rc = svd_small(Ur, n, (double*)H.dS, (double*)H.dV, n, at(H.aux, H.o_r2), n, n, at(H.aux, H.o_svd), nullptr, H.s2))) return rc;
  PL_CUDA(cudaMemcpyAsync(S, H.dS, (size_t)n * 8, cudaMemcpyDeviceToHost, H.s2));
  PL_CUDA(cudaMemcpyAsync(VT, H.dV, (size_t)n * n * 8, cudaMemcpyDeviceToHost, H.s2));
  return host_apply(Ui, Ur);
}

// P ranks: the reference's dtsqr_svd is collective (MPI inside, svd.c:602-669).  Here the exchange stays with the
// caller:  factor (local)  ->  all-gather of the n x n R's (MPI / NCCL)  ->  SVD of the (P n) x n stack, e.g. with
// pl_tsqr_svd_host_f64 / pl_tsqr_svd_f64, giving W_stack = Q2 Ur, S, VT  ->  apply with W = rows [rank n, (rank+1) n).
// R and W may be host or device pointers (unified addressing).
int pl_tsqr_host_factor_f64(double* R, const double* Ai, int64_t m, int64_t n) {
  PL_ARG(n > 0 && m >= n, 4, "need m >= n > 0");
  PL_ARG(R != nullptr, 1, "R is NULL");
  int rc = host_factor(Ai, m, n);
  if (rc) return rc;
  HostState& H = g_hs;
  PL_CUDA(cudaMemcpyAsync(R, at(H.aux, H.o_r2), (size_t)n * n * 8, cudaMemcpyDefault, H.s2));
  PL_CUDA(cudaStreamSynchronize(H.s2));        // R is ready; the last chunk's Q formation keeps running on the compute stream
  return 0;
}
// SVD of the gathered stack between the two phases: Wstack (P n x n) = Q2 Ur, S, VT.  Uses its own small device
// buffers (not the cached pipeline state, which holds the factorisation in flight).  Host or device pointers.
int pl_tsqr_host_stack_f64(double* Wstack, double* S, double* VT, const double* Rstack, int64_t P, int64_t n) {
  PL_ARG(P >= 1 && n > 0, 5, "need P >= 1, n > 0");
  PL_ARG(Wstack && S && VT && Rstack, 1, "NULL pointer");
  int rc = host_streams();
  if (rc) return rc;
  const int64_t m2 = P * n;
  const size_t mat = al((size_t)m2 * n * 8), sq = al((size_t)n * n * 8), wsb = pl_qr_workspace_bytes(m2, n);
  char* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, 2 * mat + sq + al((size_t)n * 8) + wsb + 256);
  if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return 1000 + (int)e; }
  double* dR = reinterpret_cast<double*>(buf);
  double* dW = reinterpret_cast<double*>(buf + mat);
  double* dVt = reinterpret_cast<double*>(buf + 2 * mat);
  double* dS = reinterpret_cast<double*>(buf + 2 * mat + sq);
  void* ws = buf + 2 * mat + sq + al((size_t)n * 8);
  cudaStream_t s = g_hs.s2;
  rc = (int)cudaMemcpyAsync(dR, Rstack, (size_t)m2 * n * 8, cudaMemcpyDefault, s);
  if (!rc) rc = tsqr_svd_impl(dW, dS, dVt, nullptr, dR, m2, n, 0, ws, wsb, s);
  if (!rc) rc = (int)cudaMemcpyAsync(Wstack, dW, (size_t)m2 * n * 8, cudaMemcpyDefault, s);
  if (!rc) rc = (int)cudaMemcpyAsync(S, dS, (size_t)n * 8, cudaMemcpyDefault, s);
  if (!rc) rc = (int)cudaMemcpyAsync(VT, dVt, (size_t)n * n * 8, cudaMemcpyDefault, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  cudaFree(buf);
  if (!rc && e2 != cudaSuccess) { set_error("CUDA error: %s", cudaGetErrorString(e2)); rc = 1000 + (int)e2; }
  return rc;
}
}  // extern "C"
namespace pl {
// Host-pointer collective (pl_tsqr_svd_host_dist_f64): the chunked host pipeline around one ncclAllGather.
int dist_tsqr_svd_host(pl_comm* c, double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n) {
  const int P = comm_size(c), rank = comm_rank(c);
  if (P == 1) return pl_tsqr_svd_host_f64(Ui, S, VT, Ai, m, n);
  int rc = host_factor(Ai, m, n);
  if (rc) return rc;
  HostState& H = g_hs;
  const int64_t m2 = (int64_t)P * n;
  const WsLayout Ls = make_layout(m2, n);
  const size_t mat = al((size_t)m2 * n * 8), sq = al((size_t)n * n * 8);
  char* buf = nullptr;
  if ((rc = comm_scratch(c, 2 * mat + sq + al((size_t)n * 8) + Ls.total + 256, (void**)&buf))) return rc;
  buf = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(buf) + 255) & ~(uintptr_t)255);
  double* Rst = reinterpret_cast<double*>(buf);
  double* Wst = reinterpret_cast<double*>(buf + mat);
  double* dVt = reinterpret_cast<double*>(buf + 2 * mat);
  double* dS = reinterpret_cast<double*>(buf + 2 * mat + sq);
  void* ws2 = buf + 2 * mat + sq + al((size_t)n * 8);
  cudaStream_t s2 = H.s2;
  PL_CUDA(cudaMemcpyAsync(Rst + (size_t)rank * n * n, at(H.aux, H.o_r2), (size_t)n * n * 8, cudaMemcpyDeviceToDevice, s2));
  if ((rc = comm_allgather_inplace(c, Rst, (size_t)n * n, s2))) return rc;
  if ((rc = tsqr_svd_impl(Wst, dS, dVt, nullptr, Rst, m2, n, 0, ws2, Ls.total, s2))) return rc;
  PL_CUDA(cudaMemcpyAsync(S, dS, (size_t)n * 8, cudaMemcpyDeviceToHost, s2));
  PL_CUDA(cudaMemcpyAsync(VT, dVt, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s2));
  return host_apply(Ui, Wst + (size_t)rank * n * n);
}
}  // namespace pl
extern "C" {

int pl_tsqr_host_apply_f64(double* Ui, const double* W, int64_t m, int64_t n) {
  HostState& H = g_hs;
  PL_ARG(H.valid && m == H.m && n == H.n, 3, "apply must follow pl_tsqr_host_factor_f64 with the same m, n");
  PL_ARG(Ui != nullptr && W != nullptr, 1, "NULL pointer");
  double* Wd = at(H.aux, H.o_w);
  PL_CUDA(cudaMemcpyAsync(Wd, W, (size_t)n * n * 8, cudaMemcpyDefault, H.s2));
  return host_apply(Ui, Wd);
}

}  // extern "C"
