// Collective entry points: the NCCL communicator behind the C ABI (include/pylom_b200.h).
//
// The reference's dtsqr / dtsqr_svd are collective C calls on MPI_COMM_WORLD (pyLOM/vmmath/src/svd.c:565-712): log2(P)
// blocking MPI_Send/MPI_Recv rounds up the butterfly and the same number down.  Here the exchange is ONE ncclAllGather of
// the n x n R factors; the kernel that extracts R_i writes it straight into this rank's slot of the gather buffer
// (in-place all-gather, no staging copy), every rank then factors the bit-identical (P n) x n stack redundantly.
// libnccl.so.2 is resolved with dlopen at first use, so the library has no link-time NCCL dependency (inside a torch
// process this binds to the NCCL torch already loaded; stand-alone C callers get the system libnccl).
#include "pl_common.cuh"
#include "caqr.h"
#include "../../include/pylom_b200.h"
#include <dlfcn.h>
#include <cstring>

namespace pl {

// minimal NCCL ABI (stable since NCCL 2.0): opaque comm, 128-byte unique id, enum values of ncclDataType_t
typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
constexpr int NCCL_FLOAT64 = 8;     // ncclDouble
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) { set_error("NCCL not found: dlopen(libnccl.so.2) failed: %s", dlerror()); return 2001; }
  NcclApi a; a.h = h;
  a.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(ncclComm_t*, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
  a.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
  a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather) { set_error("libnccl lacks a required symbol"); return 2002; }
  g_nccl = a;
  return 0;
}
#define PL_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    int _r = (expr);                                                                                    \
    if (_r != 0) {                                                                                      \
      pl::set_error("%s failed: %s", #expr, pl::g_nccl.GetErrorString ? pl::g_nccl.GetErrorString(_r) : "NCCL error"); \
      return 2100 + _r;                                                                                 \
    }                                                                                                   \
  } while (0)

}  // namespace pl

struct pl_comm {
  pl::ncclComm_t nccl = nullptr;
  int rank = 0, size = 1, device = 0;
  cudaStream_t side = nullptr;              // high-priority stream for exchange + small factorisations
  cudaEvent_t eR = nullptr, eS = nullptr;
  void* host_buf = nullptr; size_t host_cap = 0;   // device scratch of the host-pointer collective (grow-only)
};

using namespace pl;

// internal composition helpers implemented in api.cu
namespace pl {
size_t dist_ws_bytes(int64_t m, int64_t n, int P, int flags);
int dist_tsqr_svd(pl_comm* c, double* Ui, double* S, double* VT, double* X_mean, const double* Ai, int64_t m, int64_t n,
                  int center, int flags, void* ws, size_t ws_bytes, cudaStream_t st);
int dist_tsqr_svd_host(pl_comm* c, double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n);
int comm_allgather_inplace(pl_comm* c, double* buf, size_t count_per_rank, cudaStream_t st) {
  if (c->size == 1) return 0;
  PL_NCCL(g_nccl.AllGather(buf + (size_t)c->rank * count_per_rank, buf, count_per_rank, NCCL_FLOAT64, c->nccl, st));
  return 0;
}
int comm_side(pl_comm* c, cudaStream_t* side, cudaEvent_t* eR, cudaEvent_t* eS) {
  if (!c->side) {
    int lo = 0, hi = 0;
    PL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PL_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
    PL_CUDA(cudaEventCreateWithFlags(&c->eR, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&c->eS, cudaEventDisableTiming));
  }
  *side = c->side; *eR = c->eR; *eS = c->eS;
  return 0;
}
int comm_rank(const pl_comm* c) { return c->rank; }
int comm_size(const pl_comm* c) { return c->size; }
int comm_scratch(pl_comm* c, size_t bytes, void** out) {
  if (c->host_cap < bytes) {
    if (c->host_buf) cudaFree(c->host_buf);
    c->host_buf = nullptr; c->host_cap = 0;
    cudaError_t e = cudaMalloc(&c->host_buf, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return 1000 + (int)e; }
    c->host_cap = bytes;
  }
  *out = c->host_buf;
  return 0;
}
}  // namespace pl

extern "C" {

int pl_get_unique_id(void* id128) {
  if (!id128) { set_error("bad argument 1: id128 is NULL"); return -1; }
  int rc = nccl_load();
  if (rc) return rc;
  NcclId id;
  PL_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int pl_comm_init_rank(pl_comm_t* comm, const void* id128, int rank, int size) {
  if (!comm) { set_error("bad argument 1: comm is NULL"); return -1; }
  if (size < 1 || rank < 0 || rank >= size) { set_error("bad argument 3: need 0 <= rank < size"); return -3; }
  pl_comm* c = new pl_comm();
  c->rank = rank; c->size = size;
  PL_CUDA(cudaGetDevice(&c->device));
  if (size > 1) {
    if (!id128) { delete c; set_error("bad argument 2: id128 is NULL"); return -2; }
    int rc = nccl_load();
    if (rc) { delete c; return rc; }
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    int r = g_nccl.CommInitRank(&c->nccl, size, id, rank);
    if (r != 0) { delete c; set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return 2100 + r; }
  }
  *comm = c;
  return 0;
}

int pl_comm_rank(pl_comm_t c) { return c ? c->rank : -1; }
int pl_comm_size(pl_comm_t c) { return c ? c->size : -1; }

int pl_comm_destroy(pl_comm_t c) {
  if (!c) return 0;
  if (c->nccl) g_nccl.CommDestroy(c->nccl);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->eR) cudaEventDestroy(c->eR);
  if (c->eS) cudaEventDestroy(c->eS);
  if (c->host_buf) cudaFree(c->host_buf);
  delete c;
  return 0;
}

size_t pl_tsqr_svd_dist_workspace_bytes(pl_comm_t c, int64_t m, int64_t n, int flags) {
  if (!c || m <= 0 || n <= 0) return 0;
  return dist_ws_bytes(m, n, c->size, flags);
}

int pl_tsqr_svd_dist_f64(pl_comm_t c, double* Ui, double* S, double* VT, double* X_mean, const double* Ai, int64_t m,
                         int64_t n, int center, int flags, void* ws, size_t ws_bytes, void* stream) {
  if (!c) { set_error("bad argument 1: comm is NULL"); return -1; }
  if (!(n > 0 && m >= n)) { set_error("bad argument 7: need m >= n > 0 on every rank (svd.py:69)"); return -7; }
  if (center && !X_mean) { set_error("bad argument 5: X_mean required when center != 0"); return -5; }
  return dist_tsqr_svd(c, Ui, S, VT, X_mean, Ai, m, n, center, flags, ws, ws_bytes, (cudaStream_t)stream);
}

int pl_tsqr_svd_host_dist_f64(pl_comm_t c, double* Ui, double* S, double* VT, const double* Ai, int64_t m, int64_t n) {
  if (!c) { set_error("bad argument 1: comm is NULL"); return -1; }
  if (!(n > 0 && m >= n)) { set_error("bad argument 6: need m >= n > 0 on every rank"); return -6; }
  return dist_tsqr_svd_host(c, Ui, S, VT, Ai, m, n);
}

}  // extern "C"
