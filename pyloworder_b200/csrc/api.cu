// C ABI of libpylom_b200 (declared in include/pylom_b200.h) -- thin composition of the kernels.
#include "pl_common.cuh"
#include "caqr.h"
#include "../../include/pylom_b200.h"
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>
#include <cstdlib>

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
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static WsLayout make_layout(int64_t m, int64_t n, bool ext_vb = false) {
  WsLayout L;
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
  {
    ProfScope ps(PROF_COPY, st);
    if (center == 2) rc = center_var_rows(Vb, P.npad, X_mean, X_var, A, m, n, P.npad, st);
    else if (center) rc = center_rows(Vb, P.npad, X_mean, A, m, n, P.npad, st);
    else rc = copy_pad(Vb, P.npad, A, n, m, n, P.npad, st);
    if (rc) return rc;
    PL_CUDA(cudaMemsetAsync(Vb + (size_t)m * P.npad, 0, (size_t)(P.mrows - m) * P.npad * 8, st));
  }
  rc = caqr_factor(P, Vb, at(ws, L.tws), at(ws, L.vup), at(ws, L.vpiv), st);
  if (rc) return rc;
  if (R) rc = caqr_extract_r(P, Vb, R, n, st);
  return rc;
}

static int qr_apply_q(double* U, int64_t ldu, const double* W, int64_t ldw, int64_t nw, int64_t m, int64_t n, int flags,
                      void* ws, const WsLayout& L, cudaStream_t st, double* Vb_ext = nullptr) {
  const Plan& P = L.plan;
  double* Vb = Vb_ext ? Vb_ext : at(ws, L.vb);
  int rc;
  if (!(flags & 1)) {
    rc = caqr_form_q(P, Vb, at(ws, L.tws), at(ws, L.vup), at(ws, L.vpiv), st);
    if (rc) return rc;
  }
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
  if (svd_overlap_enabled(m)) return svd_and_apply_overlapped(Ubuf, S, VT, R, Ur, m, n, ws, L, st, Ubuf);
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
  return !off && m >= 1000000;
}
static int svd_and_apply_overlapped(double* Ui, double* S, double* VT, const double* R, double* Ur, int64_t m, int64_t n,
                                    void* ws, const WsLayout& L, cudaStream_t st, double* Vb_ext) {
  static cudaStream_t ss = nullptr;
  static cudaEvent_t eR = nullptr, eS = nullptr;
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
  rc = qr_factor(R, X_mean, Ai, m, n, center, ws, L, st);
  if (rc) return rc;
  double* Ur = at(ws, L.ur);
  if (svd_overlap_enabled(m)) return svd_and_apply_overlapped(Ui, S, VT, R, Ur, m, n, ws, L, st, nullptr);
  {
    ProfScope ps(PROF_SVD, st);
    rc = svd_small(Ur, n, S, VT, n, R, n, n, at(ws, L.svd), nullptr, st);
  }
  if (rc) return rc;
  return qr_apply_q(Ui, n, Ur, n, n, m, n, 0, ws, L, st);
}

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
void pl_host_cache_free(void) {
  for (int i = 0; i < HC_SLOTS; i++) { if (g_hc.p[i]) cudaFree(g_hc.p[i]); g_hc.p[i] = nullptr; g_hc.cap[i] = 0; }
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

// Host-pointer TSQR-SVD (replaces dtsqr_svd, pyLOM/vmmath/src/svd.c:~1390, for one rank).  The rows are processed as
// C chunks, i.e. as a two-level TSQR on one device, so that PCIe and the GPU work at the same time:
//   H2D(c+1)            ||  factor(c), R_c, explicit Q_c            (copy stream / compute stream)
//   QR of the stacked R_c, Jacobi SVD, B = Q_stack Ur               (side stream, next to the last chunk's Q_c)
//   U_c = Q_c B_c (GEMM) ||  D2H(c-1)                               (ping-pong output buffers)
int pl_tsqr_svd_host_f64(double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n) {
  PL_ARG(n > 0 && m >= n, 5, "need m >= n > 0");
  int C = host_chunks(m, n);
  const int64_t npad = round_up(n, NB);
  const bool direct = (npad == n);                       // rows land in the factorisation buffer as they are
  int64_t mc = round_up((m + C - 1) / C, TB);            // whole tiles per chunk; the last chunk takes the remainder
  C = (int)((m + mc - 1) / mc);
  if (C > 1 && m - (int64_t)(C - 1) * mc < 4 * n) C--;   // ... and absorbs a remainder that would be too short
  if (C == 1) mc = m;
  struct Chunk { int64_t r0, rows; Plan P; size_t vb, tws, vup, vpiv; };
  std::vector<Chunk> ch(C);
  size_t vb_bytes = 0, aux_bytes = 0;
  int64_t max_rows = 0;
  for (int c = 0; c < C; c++) {
    Chunk& k = ch[c];
    k.r0 = (int64_t)c * mc; k.rows = (c == C - 1) ? m - k.r0 : mc;
    if (k.rows < n) { set_error("host pipeline: chunk %d has %lld rows < n", c, (long long)k.rows); return -5; }
    k.P = make_plan(k.rows, n);
    k.vb = vb_bytes;   vb_bytes += al((size_t)k.P.mrows * k.P.npad * 8);
    k.tws = aux_bytes;  aux_bytes += al((size_t)k.P.t_tiles * NB * NB * 8);
    k.vup = aux_bytes;  aux_bytes += al((size_t)(k.P.vup_tiles > 0 ? k.P.vup_tiles : 1) * TB * NB * 8);
    k.vpiv = aux_bytes; aux_bytes += al((size_t)k.P.vpiv_strips * NB * NB * 8);
    if (k.rows > max_rows) max_rows = k.rows;
  }
  // small buffers: stacked R (C n x n), B = Q_stack Ur (C n x n), R2, Ur, packed B chunk, Jacobi scratch, level-2 workspace
  const int64_t kp = round_up(n, 16), np = round_up(n, 64);
  const int64_t m2 = (int64_t)C * n;
  WsLayout L2 = make_layout(m2, n);
  size_t off = aux_bytes;
  const size_t o_rs = off;  off += al((size_t)m2 * n * 8);
  const size_t o_bs = off;  off += al((size_t)m2 * n * 8);
  const size_t o_r2 = off;  off += al((size_t)n * n * 8);
  const size_t o_ur = off;  off += al((size_t)n * n * 8);
  const size_t o_bp = off;  off += al((size_t)kp * np * 8);
  const size_t o_svd = off; off += al((size_t)svd_small_scratch_doubles(n) * 8);
  const size_t o_ws2 = off; off += C > 1 ? L2.total : 0;
  void *dVb = nullptr, *dOut = nullptr, *dS = nullptr, *dV = nullptr, *aux = nullptr, *stage = nullptr;
  const size_t chunk_out = al((size_t)max_rows * n * 8);
  int rc;
  if ((rc = hc_get(0, vb_bytes, &dVb)) || (rc = hc_get(1, 2 * chunk_out, &dOut)) || (rc = hc_get(2, (size_t)n * 8, &dS)) ||
      (rc = hc_get(3, (size_t)n * n * 8, &dV)) || (rc = hc_get(4, off, &aux)) ||
      (!direct && (rc = hc_get(5, 2 * chunk_out, &stage)))) { pl_host_cache_free(); return rc; }
  static cudaStream_t st = nullptr, cs = nullptr, s2 = nullptr;   // compute / copy / small-factor streams
  static cudaEvent_t eR = nullptr, eB = nullptr;
  if (!st) {
    int lo = 0, hi = 0;
    PL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PL_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    PL_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    PL_CUDA(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi));
    PL_CUDA(cudaEventCreateWithFlags(&eR, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&eB, cudaEventDisableTiming));
  }
  std::vector<cudaEvent_t> ev(C), evs(2), evd(2);
  for (auto* v : {&ev, &evs, &evd}) for (auto& e : *v) PL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  struct EvGuard { std::vector<cudaEvent_t>*a, *b, *c; ~EvGuard() { for (auto* v : {a, b, c}) for (auto e : *v) cudaEventDestroy(e); } }
      guard{&ev, &evs, &evd};
  double* Rs = at(aux, o_rs);
  double* Bs = at(aux, o_bs);
  double* R2 = at(aux, o_r2);
  double* Ur = at(aux, o_ur);
  double* Bp = at(aux, o_bp);
  // ---- phase 1: chunks arrive, get factored and turned into explicit Q_c while the next chunk is on the wire
  for (int c = 0; c < C; c++) {
    Chunk& k = ch[c];
    double* Vb = at(dVb, k.vb);
    const size_t bytes = (size_t)k.rows * n * 8;
    if (direct) {
      PL_CUDA(cudaMemcpyAsync(Vb, Ai + k.r0 * n, bytes, cudaMemcpyHostToDevice, cs));
      PL_CUDA(cudaEventRecord(ev[c], cs));
      PL_CUDA(cudaStreamWaitEvent(st, ev[c], 0));
    } else {
      double* sg = reinterpret_cast<double*>(static_cast<char*>(stage) + (size_t)(c & 1) * chunk_out);
      if (c >= 2) PL_CUDA(cudaStreamWaitEvent(cs, evs[c & 1], 0));           // staging buffer consumed
      PL_CUDA(cudaMemcpyAsync(sg, Ai + k.r0 * n, bytes, cudaMemcpyHostToDevice, cs));
      PL_CUDA(cudaEventRecord(ev[c], cs));
      PL_CUDA(cudaStreamWaitEvent(st, ev[c], 0));
      if ((rc = copy_pad(Vb, k.P.npad, sg, n, k.rows, n, k.P.npad, st))) return rc;
      PL_CUDA(cudaEventRecord(evs[c & 1], st));
    }
    PL_CUDA(cudaMemsetAsync(Vb + (size_t)k.rows * k.P.npad, 0, (size_t)(k.P.mrows - k.rows) * k.P.npad * 8, st));
    if ((rc = caqr_factor(k.P, Vb, at(aux, k.tws), at(aux, k.vup), at(aux, k.vpiv), st))) return rc;
    if ((rc = caqr_extract_r(k.P, Vb, C > 1 ? Rs + (size_t)c * n * n : R2, n, st))) return rc;
    if (c == C - 1) PL_CUDA(cudaEventRecord(eR, st));
    if ((rc = caqr_form_q(k.P, Vb, at(aux, k.tws), at(aux, k.vup), at(aux, k.vpiv), st))) return rc;
  }
  // ---- phase 2 (side stream, beside the last chunk's Q formation): stacked-R QR, Jacobi SVD, B = Q_stack Ur
  PL_CUDA(cudaStreamWaitEvent(s2, eR, 0));
  const double* B = Ur;
  if (C > 1) {
    void* ws2 = static_cast<char*>(aux) + o_ws2;
    if ((rc = qr_factor(R2, nullptr, Rs, m2, n, 0, ws2, L2, s2))) return rc;
  }
  if ((rc = svd_small(Ur, n, (double*)dS, (double*)dV, n, R2, n, n, at(aux, o_svd), nullptr, s2))) return rc;
  if (C > 1) {
    void* ws2 = static_cast<char*>(aux) + o_ws2;
    if ((rc = qr_apply_q(Bs, n, Ur, n, n, m2, n, 0, ws2, L2, s2))) return rc;
    B = Bs;
  }
  PL_CUDA(cudaMemcpyAsync(S, dS, (size_t)n * 8, cudaMemcpyDeviceToHost, s2));
  PL_CUDA(cudaMemcpyAsync(VT, dV, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s2));
  PL_CUDA(cudaEventRecord(eB, s2));
  PL_CUDA(cudaStreamWaitEvent(st, eB, 0));
  // ---- phase 3: U_c = Q_c B_c; the D2H of a chunk overlaps the GEMM of the next one
  for (int c = 0; c < C; c++) {
    Chunk& k = ch[c];
    double* out = reinterpret_cast<double*>(static_cast<char*>(dOut) + (size_t)(c & 1) * chunk_out);
    if (c >= 2) PL_CUDA(cudaStreamWaitEvent(st, evd[c & 1], 0));             // output buffer drained
    if ((rc = pad_small(Bp, kp, np, B + (size_t)c * n * n, n, n, n, nullptr, st))) return rc;
    if ((rc = gemm_tall(out, n, at(dVb, k.vb), k.P.npad, Bp, np, k.rows, n, kp, st))) return rc;
    PL_CUDA(cudaEventRecord(ev[c], st));
    PL_CUDA(cudaStreamWaitEvent(cs, ev[c], 0));
    PL_CUDA(cudaMemcpyAsync(Ui + k.r0 * n, out, (size_t)k.rows * n * 8, cudaMemcpyDeviceToHost, cs));
    PL_CUDA(cudaEventRecord(evd[c & 1], cs));
  }
  PL_CUDA(cudaStreamSynchronize(s2));
  PL_CUDA(cudaStreamSynchronize(st));
  PL_CUDA(cudaStreamSynchronize(cs));
  return 0;
}

}  // extern "C"
