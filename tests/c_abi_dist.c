/* Two ranks through the C ABI, no Python and no torch: one PROCESS per GPU (fork before any CUDA call), exactly what a
 * Cython / C caller of the reference's collective dtsqr_svd (pyLOM/vmmath/src/svd.c:678-712) does with MPI ranks.
 *   rank 0: pl_get_unique_id -> (a pipe here, MPI_Bcast in the reference's world)
 *   every rank: cudaSetDevice(rank), pl_comm_init_rank, pl_tsqr_svd_host_dist_f64 on its row shard
 * Checks: S identical on both ranks (bitwise), S equal to the single-rank pl_tsqr_svd_host_f64 of the whole matrix to
 * 1e-12 relative, ||U^T U - I||_max <= 1e-12 over both shards, ||A - U S VT||_max <= 1e-12 ||A||.
 * Build + run: see tests/test_gpu_parity.py::test_c_abi_two_ranks_no_torch.  Test infrastructure, not product code. */
#include <math.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pylom_b200.h"

extern int cudaSetDevice(int);
extern int cudaGetDeviceCount(int*);

#define P 2
static int64_t M = 60000, N = 96;
static double *A, *U, *S[P], *VT[P];
static int* rcs;

static void split(int64_t m, int r, int p, int64_t* r0, int64_t* r1) {   /* worksplit, pyLOM/utils/parall.py:24-48 */
  int64_t base = m / p, rem = m % p;
  *r0 = r * base + (r < rem ? r : rem);
  *r1 = *r0 + base + (r < rem ? 1 : 0);
}
static void* shared(size_t bytes) { return mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0); }

static int rank_main(int rank, int fd_id_read, int fd_id_write) {
  unsigned char uid[PL_UNIQUE_ID_BYTES];
  int64_t r0, r1;
  split(M, rank, P, &r0, &r1);
  if (cudaSetDevice(rank) != 0) return 1;
  if (rank == 0) {
    if (pl_get_unique_id(uid) != 0) { fprintf(stderr, "unique id: %s\n", pl_last_error()); return 1; }
    if (write(fd_id_write, uid, sizeof(uid)) != (ssize_t)sizeof(uid)) return 1;          /* "MPI_Bcast" of the id */
  } else {
    if (read(fd_id_read, uid, sizeof(uid)) != (ssize_t)sizeof(uid)) return 1;
  }
  pl_comm_t comm;
  if (pl_comm_init_rank(&comm, uid, rank, P) != 0) { fprintf(stderr, "init: %s\n", pl_last_error()); return 1; }
  int rc = pl_tsqr_svd_host_dist_f64(comm, U + r0 * N, S[rank], VT[rank], A + r0 * N, r1 - r0, N);
  if (rc != 0) fprintf(stderr, "rank %d: %s\n", rank, pl_last_error());
  pl_comm_destroy(comm);
  return rc;
}

int main(void) {
  A = shared(sizeof(double) * M * N); U = shared(sizeof(double) * M * N); rcs = shared(sizeof(int) * P);
  for (int r = 0; r < P; r++) { S[r] = shared(sizeof(double) * N); VT[r] = shared(sizeof(double) * N * N); rcs[r] = 1; }
  uint64_t z = 88172645463325252ULL;                            /* xorshift: reproducible dense matrix with a graded column scaling */
  for (int64_t i = 0; i < M; i++)
    for (int64_t j = 0; j < N; j++) {
      z ^= z << 13; z ^= z >> 7; z ^= z << 17;
      A[i * N + j] = ((double)(z >> 11) / 9007199254740992.0 - 0.5) * pow(10.0, -4.0 * (double)j / (double)N) + sin(1e-3 * (double)i * (double)(j + 1));
    }
  int fds[2];
  if (pipe(fds) != 0) return 1;
  pid_t child = fork();                                          /* before the first CUDA call: each process owns one GPU */
  if (child == 0) { rcs[1] = rank_main(1, fds[0], fds[1]); _exit(rcs[1] ? 1 : 0); }
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev < P) { printf("C_ABI_DIST SKIP (needs %d GPUs)\n", P); return 0; }
  rcs[0] = rank_main(0, fds[0], fds[1]);
  int status = 0;
  waitpid(child, &status, 0);
  int ok = (rcs[0] == 0) && (rcs[1] == 0);
  if (!ok) { printf("C_ABI_DIST FAIL (return codes %d %d)\n", rcs[0], rcs[1]); return 1; }
  ok &= memcmp(S[0], S[1], sizeof(double) * N) == 0 && memcmp(VT[0], VT[1], sizeof(double) * N * N) == 0;
  /* single-rank reference through the same ABI */
  double *U1 = malloc(sizeof(double) * M * N), *S1 = malloc(sizeof(double) * N), *V1 = malloc(sizeof(double) * N * N);
  cudaSetDevice(0);
  if (pl_tsqr_svd_host_f64(U1, S1, V1, A, M, N) != 0) { printf("C_ABI_DIST FAIL (single rank: %s)\n", pl_last_error()); return 1; }
  double ds = 0.0;
  for (int64_t j = 0; j < N; j++) ds = fmax(ds, fabs(S1[j] - S[0][j]) / S1[0]);
  /* orthogonality and reconstruction of the distributed result */
  double orth = 0.0, rec = 0.0, amax = 0.0;
  for (int64_t a = 0; a < N; a++)
    for (int64_t b = a; b < N; b++) {
      double s = 0.0;
      for (int64_t i = 0; i < M; i++) s += U[i * N + a] * U[i * N + b];
      orth = fmax(orth, fabs(s - (a == b ? 1.0 : 0.0)));
    }
  for (int64_t i = 0; i < M; i += 97)
    for (int64_t j = 0; j < N; j++) {
      double s = 0.0;
      for (int64_t k = 0; k < N; k++) s += U[i * N + k] * S[0][k] * VT[0][k * N + j];
      rec = fmax(rec, fabs(s - A[i * N + j])); amax = fmax(amax, fabs(A[i * N + j]));
    }
  printf("S identical across ranks: %d, |S - S_1rank|/S_0 = %.2e, ||UtU - I|| = %.2e, recon = %.2e (|A| = %.2e)\n", ok, ds, orth, rec, amax);
  ok &= ds <= 1e-12 && orth <= 1e-12 && rec <= 1e-12 * amax;
  printf(ok ? "C_ABI_DIST PASS\n" : "C_ABI_DIST FAIL\n");
  return ok ? 0 : 1;
}
