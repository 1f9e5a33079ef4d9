// Shared device/host helpers for libpylom_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

namespace pl {

constexpr int NB = 32;        // panel width (columns per block reflector)
constexpr int G  = 4;         // NB-row blocks per tile
constexpr int TB = NB * G;    // tile rows (128)
constexpr int SMAX = 16;      // max tiles per strip (flat tree inside one CTA)

// ---- error reporting (C ABI returns int, message via pl_last_error) ----------------------
void set_error(const char* fmt, ...);
const char* last_error();
void count_launches(long long k);   // instrumentation: kernels launched by this library

#define PL_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1000 + (int)_e;                                                               \
    }                                                                                      \
  } while (0)

#define PL_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    pl::count_launches(1);                                                                 \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      pl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1000 + (int)_e;                                                               \
    }                                                                                      \
  } while (0)

// ---- optional per-kernel-class event timing (bench.py roofline leg; off by default) ---------
enum ProfClass { PROF_COPY = 0, PROF_PANEL = 1, PROF_UPDATE_F = 2, PROF_UPDATE_Q = 3, PROF_GEMM = 4, PROF_SVD = 5, PROF_MISC = 6, PROF_SMALL = 7, PROF_NCLS = 8 };
bool prof_enabled();
void prof_begin(int cls, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(int cls, cudaStream_t s) : st(s), on(prof_enabled()) { if (on) prof_begin(cls, st); }
  ~ProfScope() { if (on) prof_end(st); }
};

// ---- per-device state: one process may drive several GPUs (the header's threading note allows it), so streams, events,
// function attributes and cached device buffers are keyed by the CUDA device that is current at the call.
constexpr int MAX_DEV = 64;
static inline int cur_dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < MAX_DEV) ? d : 0; }
// One-shot flag per (call site, device): `if (once_per_device(flags)) cudaFuncSetAttribute(...)`.
struct DevOnce { bool done[MAX_DEV] = {}; };
static inline bool first_on_device(DevOnce& f) { const int d = cur_dev(); if (f.done[d]) return false; f.done[d] = true; return true; }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

#ifdef __CUDACC__
// ---- FP64 tensor-core MMA (DMMA).  Fragment layout verified on B200 by probes/dmma_probe.cu:
//   g = lane>>2, t = lane&3
//   A (16x16,row): a[i] = A[g + 8*(i&1)][t + 4*(i>>1)]
//   B (16x8, col): b[i] = B[k = t + 4*i][n = g]
//   C/D (16x8)   : c[0],c[1] = C[g][2t],C[g][2t+1];  c[2],c[3] = C[g+8][2t],C[g+8][2t+1]
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
      "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- cp.async (LDGSTS) 16-byte copies with zero fill --------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__

}  // namespace pl
