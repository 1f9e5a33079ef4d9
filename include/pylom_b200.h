/* libpylom_b200 -- C ABI of the B200-native POD / TSQR-SVD hot path.
 *
 * Drop-in boundary for pyLOM's compiled math layer (pyLOM/vmmath/src/{averaging,svd,vector_matrix,
 * stats}.h as declared to Cython in pyLOM/vmmath/cfuncs.pxd).  Every entry point names the
 * reference function it replaces.  Differences from the reference ABI, all deliberate:
 *   - sizes are int64_t (the reference's `int m*n` overflows at 2^31 elements, svd.c:586,697);
 *   - array arguments are DEVICE pointers unless the name ends in `_host`;
 *   - no hidden malloc on the device: the caller passes a workspace sized by the matching
 *     *_workspace_bytes() query (the reference mallocs/frees scratch inside each call);
 *   - explicit stream (a cudaStream_t passed as void*); calls are stream ordered and asynchronous (the Jacobi SVD
 *     tracks its convergence on the device; only the `_host` entry points and n > 1024 SVDs synchronise);
 *   - one process may drive several GPUs: internal streams, events and cached buffers are kept per CUDA device, the
 *     device current at the call must be the one the pointers live on;
 *   - return value 0 = ok, < 0 = bad argument (minus its position), > 0 = CUDA / convergence error;
 *     pl_last_error() returns the message (the reference returns the LAPACK info, svd.pyx:399).
 * Matrices are row-major, contiguous, fp64, exactly as in the reference (C order numpy arrays).
 */
#ifndef PYLOM_B200_H
#define PYLOM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int pl_version(void);
const char* pl_last_error(void);

/* ---- averaging:  pyLOM/vmmath/src/averaging.h ------------------------------------------------ */
/* replaces dtemporal_mean(double *out, double *X, const int m, const int n)   averaging.c:29-46  */
int pl_temporal_mean_f64(double* out, const double* X, int64_t m, int64_t n, void* stream);
/* replaces dsubtract_mean(double *out, double *X, double *X_mean, m, n)       averaging.c:109-124 */
int pl_subtract_mean_f64(double* out, const double* X, const double* X_mean, int64_t m, int64_t n, void* stream);
/* fused temporal_mean + subtract_mean (what POD.run does back to back, POD/wrapper.pyx:127-133) */
int pl_center_f64(double* Y, double* X_mean, const double* X, int64_t m, int64_t n, void* stream);

/* replaces dtemporal_variance(double *out, double *X, double *Xmean, m, n): population variance  averaging.c:70-90 */
int pl_temporal_variance_f64(double* out, const double* X, const double* X_mean, int64_t m, int64_t n, void* stream);
/* replaces dsubtract_mean + dnorm_variance (averaging.c:143-158): out = (X - X_mean) / X_var, the body of
 * norm_variance(X, X_mean, X_var) in pyLOM/vmmath/averaging.py:61-74 */
int pl_norm_variance_f64(double* out, const double* X, const double* X_mean, const double* X_var, int64_t m, int64_t n, void* stream);

/* ---- dense helpers:  pyLOM/vmmath/src/vector_matrix.h ---------------------------------------- */
/* replaces dmatmul(double *C, double *A, double *B, m, n, k): C(m,n) = A(m,k) B(k,n)  vector_matrix.c:234-242.
 * lda/ldc allow the strided views POD.truncate returns (POD/wrapper.py:78-80). */
size_t pl_matmul_workspace_bytes(int64_t n, int64_t k);
int pl_matmul_f64(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                  int64_t m, int64_t n, int64_t k, void* ws, size_t ws_bytes, void* stream);
/* rank-local part of dmatmulp(double *C, double *A, double *B, m, n, k) (vector_matrix.c:344-356: cblas_dgemm +
 * MPI_Allreduce) for the shapes pyLOM calls it with -- matmulp(Ai.T, Qi), matmulp(Qi.T, Ai) in randomized_qr
 * (vmmath/svd.py:139,143) and matmulp(U.T, Y) in DMD: C(a,b) = X^T Y with X (m,a), Y (m,b) row-major tall operands,
 * i.e. a reduction over the distributed rows.  The caller adds C over the ranks (one all-reduce). */
size_t pl_matmul_tn_workspace_bytes(int64_t a, int64_t b);
int pl_matmul_tn_f64(double* C, int64_t ldc, const double* X, int64_t ldx, int64_t a, const double* Y, int64_t ldy,
                     int64_t b, int64_t m, void* ws, size_t ws_bytes, void* stream);
/* fp32 callers (the reference's float variants stsqr_svd / stemporal_mean ..., pyLOM/vmmath/src/svd.c:416-563,
 * averaging.c:18-27): the Python layer widens fp32 inputs to fp64 on the device with these two streaming kernels, runs
 * the fp64 path and narrows the results, so fp32 callers get fp32 arrays back (computed more accurately than sgeqrf would). */
int pl_widen_f32_f64(double* dst, const float* src, int64_t count, void* stream);
int pl_narrow_f64_f32(float* dst, const double* src, int64_t count, void* stream);

/* complex128 callers (ztsqr_svd, pyLOM/vmmath/src/svd.c:714-1010; SPOD, pyLOM/SPOD/wrapper.py:80-86): the Python layer
 * factors the real embedding Ahat = [[Ar, -Ai], [Ai, Ar]] (2m x 2n) with the fp64 path and keeps one member of every
 * pair of equal singular values.  pl_complex_embed_f64 builds Ahat from the interleaved complex matrix A (m x n);
 * pl_complex_pack_f64 writes Uc (m x n complex) = (P_top - Q_bot) + i (Q_top + P_bot) for 2m x n real P, Q (Q may be NULL). */
int pl_complex_embed_f64(double* Ahat, const double* A, int64_t m, int64_t n, void* stream);
int pl_complex_pack_f64(double* Uc, const double* P, const double* Q, int64_t m, int64_t n, void* stream);

/* replaces dvecmat(double *v, double *A, m, n): C[i,:] = v[i] A[i,:] (out of place)  vector_matrix.c:401-414 */
int pl_vecmat_f64(double* C, const double* v, const double* A, int64_t m, int64_t n, void* stream);

/* ---- statistics:  pyLOM/vmmath/src/stats.h --------------------------------------------------- */
/* the two local sums of dRMSE_relative (stats.c:44-72): out2[0] = sum (A-B)^2, out2[1] = sum A^2.
 * ws: >= pl_rmse_workspace_bytes() */
size_t pl_rmse_workspace_bytes(void);
int pl_rmse_sums_f64(double* out2, const double* A, const double* B, int64_t count, void* ws, void* stream);

/* ---- QR / SVD:  pyLOM/vmmath/src/svd.h:12-39 -------------------------------------------------- */
/* Workspace for one tall matrix (m x n): padded factorisation buffer + T factors + scratch. */
size_t pl_qr_workspace_bytes(int64_t m, int64_t n);

/* First half of dqr (svd.c:280-321, LAPACKE_dgeqrf): copy A (optionally minus its row means, the
 * POD.run centering) into the workspace, Householder-factor it there, return R (n x n, zeros below
 * the diagonal, svd.c:304-307).  A is not modified.  X_mean (m) may be NULL when center == 0. */
int pl_qr_factor_f64(double* R, double* X_mean, const double* A, int64_t m, int64_t n, int center,
                     void* ws, size_t ws_bytes, void* stream);
/* Same with the POD.run(divide_variance=True) preprocessing fused in: the factored matrix is
 * (A - rowmean) / rowvariance (POD/wrapper.py:36-38); X_mean and X_var (m each) are outputs. */
int pl_qr_factor_var_f64(double* R, double* X_mean, double* X_var, const double* A, int64_t m, int64_t n,
                         void* ws, size_t ws_bytes, void* stream);
/* Second half of dqr (LAPACKE_dorgqr) fused with the back-multiplies of dtsqr/dtsqr_svd
 * (dmatmul at svd.c:673 and svd.c:708):  U(m, nw) = Q1 * W, W (n x nw, ldw) on the device.
 * W == NULL gives U = Q1 (nw must equal n).  Must follow pl_qr_factor_f64 on the same workspace.
 * flags bit 0: Q1 has already been formed in this workspace by an earlier call.
 * flags bit 1: only form Q1 in the workspace and return (U, W unused) -- lets the caller overlap the Q formation with
 *              the exchange of the R factors and the small SVD on another stream, then call again with bit 0. */
int pl_qr_apply_q_f64(double* U, int64_t ldu, const double* W, int64_t ldw, int64_t nw, int64_t m, int64_t n,
                      int flags, void* ws, size_t ws_bytes, void* stream);

/* replaces dsvd(double *U, double *S, double *VT, double *Y, m, n) for the square n x n case it is
 * used for on this path (svd.c:83-139, call site svd.c:706).  S descending, VT = V^T. */
size_t pl_svd_workspace_bytes(int64_t n);
int pl_svd_f64(double* U, double* S, double* VT, const double* Y, int64_t n, void* ws, size_t ws_bytes, void* stream);

/* replaces dtsqr_svd(double *Ui, double *S, double *VT, double *Ai, m, n) on ONE rank  svd.c:678-712.
 * ws: >= pl_qr_workspace_bytes(m, n).  With several ranks the host layer composes
 * pl_qr_factor_f64 -> NCCL all-gather of R -> pl_qr_factor_f64/apply on the stack -> pl_svd_f64 ->
 * pl_qr_apply_q_f64 (see pyloworder_b200/vmmath/svd.py). */
int pl_tsqr_svd_f64(double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n,
                    void* ws, size_t ws_bytes, void* stream);
/* POD.run(remove_mean) on one rank: _drun in pyLOM/POD/wrapper.pyx:95-151 (centering fused into the
 * copy that feeds the factorisation).  X_mean (m) receives the row means when remove_mean != 0. */
int pl_pod_run_f64(double* U, double* S, double* VT, double* X_mean, const double* X, int64_t m, int64_t n,
                   int remove_mean, void* ws, size_t ws_bytes, void* stream);
/* POD.reconstruct: X(m,n) = U(m,N) diag(S) VT(N,n)   _dreconstruct, POD/wrapper.pyx:324-351 */
int pl_reconstruct_f64(double* X, const double* U, int64_t ldu, const double* S, const double* VT, int64_t ldvt,
                       int64_t m, int64_t N, int64_t n, void* ws, size_t ws_bytes, void* stream);

/* In-place variants (n % 32 == 0): the output buffer doubles as the factorisation buffer, so the footprint is
 * A + U + T-factors (~2.3 x A instead of ~3.4 x A; what lets a 1.25e8 x 64 shard of BASELINE config 5 fit in 180 GB).
 * Ubuf must hold pl_qr_inplace_rows(m, n) x n doubles; rows [0, m) are the result U (or Q), the tail is scratch.
 * Same semantics as pl_qr_factor_f64 / pl_qr_apply_q_f64 / pl_pod_run_f64 otherwise. */
int64_t pl_qr_inplace_rows(int64_t m, int64_t n);
size_t pl_qr_workspace_bytes_inplace(int64_t m, int64_t n);
int pl_qr_factor_inplace_f64(double* R, double* X_mean, double* Ubuf, const double* A, int64_t m, int64_t n, int center,
                             void* ws, size_t ws_bytes, void* stream);
int pl_qr_apply_q_inplace_f64(double* Ubuf, const double* W, int64_t ldw, int64_t m, int64_t n, int flags,
                              void* ws, size_t ws_bytes, void* stream);
int pl_pod_run_inplace_f64(double* Ubuf, double* S, double* VT, double* X_mean, const double* X, int64_t m, int64_t n,
                           int remove_mean, void* ws, size_t ws_bytes, void* stream);

/* Same signature and meaning as the reference's dtsqr_svd but HOST pointers (what a ctypes / Cython
 * binding of the reference would pass): allocates device memory, copies in, computes, copies back. */
int pl_tsqr_svd_host_f64(double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n);
/* P ranks from host memory.  The reference's dtsqr_svd is collective (MPI inside, svd.c:602-669); here the exchange
 * stays with the caller so that any transport works:
 *   pl_tsqr_host_factor_f64(R, Ai, m_i, n)      local phase: chunked H2D || factorisation, R (n x n) out
 *   all-gather the P matrices R (MPI_Allgather / ncclAllGather)  ->  Rstack (P n x n)
 *   pl_tsqr_host_stack_f64(Wstack, S, VT, Rstack, P, n)   SVD of the stack: Wstack (P n x n) = Q2 Ur, S, VT
 *   pl_tsqr_host_apply_f64(Ui, Wstack + rank n n, m_i, n)   Ui = Q1_i W, chunked GEMM || D2H
 * R, Rstack, Wstack, S, VT and W may be host or device pointers; Ai / Ui are host pointers.  State is kept between the two calls
 * (one factorisation in flight per process). */
int pl_tsqr_host_factor_f64(double* R, const double* Ai, int64_t m, int64_t n);
int pl_tsqr_host_stack_f64(double* Wstack, double* S, double* VT, const double* Rstack, int64_t P, int64_t n);
int pl_tsqr_host_apply_f64(double* Ui, const double* W, int64_t m, int64_t n);
/* row partition the host pipeline uses for an m x n shard (pure host arithmetic; returns the chunk count and fills
 * rows[0 .. min(count, max_chunks))).  PL_HOST_CHUNKS overrides the target count. */
int pl_host_chunk_rows(int64_t m, int64_t n, int64_t* rows, int max_chunks);
/* the host entry point keeps its device buffers between calls (grow-only); this releases them */
void pl_host_cache_free(void);

/* ---- P ranks, ONE collective call per rank (what dtsqr_svd is in the reference: MPI inside, svd.c:565-712) ---------
 * The communicator wraps an NCCL communicator (libnccl.so.2 is loaded at run time with dlopen; NCCL is only needed
 * when these entry points are used).  Bootstrap exactly like MPI + NCCL programs do: rank 0 calls
 * pl_get_unique_id(), the 128 bytes travel to the other ranks by any means (MPI_Bcast in the reference's world, a
 * file, torch.distributed), every rank calls pl_comm_init_rank() with the CUDA device it drives current.
 * Replaces MPI_COMM_WORLD + the 2 ceil(log2 P) blocking MPI_Send/MPI_Recv rounds of the butterfly (svd.c:602-669) by
 * a single ncclAllGather of the n x n R factors. */
typedef struct pl_comm* pl_comm_t;
#define PL_UNIQUE_ID_BYTES 128
int pl_get_unique_id(void* id128);
int pl_comm_init_rank(pl_comm_t* comm, const void* id128, int rank, int size);
int pl_comm_rank(pl_comm_t comm);
int pl_comm_size(pl_comm_t comm);
int pl_comm_destroy(pl_comm_t comm);
/* replaces dtsqr_svd (svd.c:678-712) / _drun (POD/wrapper.pyx:95-151) on P ranks, DEVICE pointers: local Householder
 * QR (R written straight into this rank's slot of the gather buffer), ncclAllGather, QR of the (P n) x n stack and
 * Jacobi SVD redundantly on every rank (bit-identical S, VT), Ui = Q1_i (Q2_i Ur).  For shards of >= 1 M rows the
 * explicit Q1 is formed on `stream` while exchange + small factorisations run on an internal high-priority stream.
 * center != 0: the row means are removed first (X_mean out).  flags bit 0: in-place variant, Ui is the
 * pl_qr_inplace_rows(m, n) x n buffer.  ws >= pl_tsqr_svd_dist_workspace_bytes(comm, m, n, flags). */
size_t pl_tsqr_svd_dist_workspace_bytes(pl_comm_t comm, int64_t m, int64_t n, int flags);
int pl_tsqr_svd_dist_f64(pl_comm_t comm, double* Ui, double* S, double* VT, double* X_mean, const double* Ai, int64_t m,
                         int64_t n, int center, int flags, void* ws, size_t ws_bytes, void* stream);
/* the same collective with HOST pointers and the argument list of the reference's dtsqr_svd (+ the communicator):
 * chunked H2D || factorisation, ncclAllGather of R inside the library, stack SVD, chunked GEMM || D2H. */
int pl_tsqr_svd_host_dist_f64(pl_comm_t comm, double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n);

/* instrumentation: number of kernel launches issued by this library since load */
int64_t pl_launch_count(void);
/* optional CUDA-event timing per kernel class (0 copy/center, 1 panel, 2 update(factor), 3 update(form Q),
 * 4 tall GEMM, 5 small SVD, 6 misc, 7 small-n fused TSQR kernels); pl_profile_read synchronises, fills ms / launch counts and resets. */
void pl_profile_enable(int on);
int pl_profile_read(double* ms_by_class, int64_t* launches_by_class, int ncls);

#ifdef __cplusplus
}
#endif
#endif
