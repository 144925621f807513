/* Single-rank stand-in for <mpi.h>, written for this repo's oracle harness only.
 * The build container has no MPI; the reference's native sources include <mpi.h>
 * (adFVM/cpp/include/parallel.hpp:5, adFVM/cpp/mesh.cpp:9-10, generated code via adpy/variable.py:465).
 * With one rank every point-to-point call is unreachable (no processor patches) and every
 * reduction is the identity, so these inline definitions are exact, not approximations. */
#ifndef ADFVM_B200_MPI_STUB_H
#define ADFVM_B200_MPI_STUB_H
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_INT 14
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
static inline int adfvm_stub_size(MPI_Datatype t) { return t == MPI_FLOAT ? 4 : (t == MPI_DOUBLE ? 8 : 4); }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return 0; }
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) {
    (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r;
    fprintf(stderr, "mpi stub: MPI_Isend reached on a single rank\n"); abort(); return 1; }
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) {
    (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r;
    fprintf(stderr, "mpi stub: MPI_Irecv reached on a single rank\n"); abort(); return 1; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n; (void)r; (void)s; return 0; }
static inline int MPI_Allreduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    (void)op; (void)c; memcpy(out, in, (size_t)n * adfvm_stub_size(t)); return 0; }
static inline int MPI_Reduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
    (void)op; (void)c; (void)root; memcpy(out, in, (size_t)n * adfvm_stub_size(t)); return 0; }
#endif
