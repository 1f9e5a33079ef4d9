/* Single-rank stand-in for <mpi.h> so the reference's C sources compile without an MPI
 * installation (none exists in this image).  TEST INFRASTRUCTURE: rank 0 of 1, Allreduce is
 * a copy, Send/Recv never happen at size 1.  Not part of the product. */
#ifndef PL_SHIM_MPI_H
#define PL_SHIM_MPI_H
#include <string.h>
#include <stddef.h>
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef int MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUM 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_C_FLOAT_COMPLEX 8
#define MPI_C_DOUBLE_COMPLEX 16
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) {
  (void)o; (void)c; memcpy(r, s, (size_t)n * (size_t)t); return 0; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; return 1; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) {
  (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; return 1; }
#endif
