/*
 * Single-rank stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference (a2d-shells / TACS) is written against MPI; this image has no
 * MPI installation.  To build the reference *unmodified* into oracle/_ref/ as a
 * one-rank checker we provide this header: rank 0 of a world of size 1.
 * Collectives degenerate to memcpy, point-to-point calls abort (they are never
 * reached on one rank), MPI-IO fails softly.  A datatype handle is its size in
 * bytes, which is all the collectives below need.
 */
#ifndef A2DS_ORACLE_STUB_MPI_H
#define A2DS_ORACLE_STUB_MPI_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
} MPI_Status;
typedef struct a2ds_stub_file *MPI_File;
typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);

#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_COMM_NULL 0
#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-2)
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_MAX_ERROR_STRING 256
#define MPI_MODE_RDONLY 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_CREATE 1
#define MPI_UNDEFINED (-32766)

/* datatype handle == extent in bytes */
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_DOUBLE_COMPLEX 16
#define MPI_LONG 8

#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3

static inline void a2ds_stub_die(const char *what) {
  fprintf(stderr, "[oracle mpi stub] %s called on a single-rank build\n", what);
  abort();
}
static inline void a2ds_stub_copy(const void *s, void *r, size_t nbytes) {
  if (s != MPI_IN_PLACE && s != r && nbytes) memcpy(r, s, nbytes);
}

static inline int MPI_Init(int *a, char ***b) { (void)a; (void)b; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *res) {
  *res = (a == b) ? MPI_IDENT : MPI_CONGRUENT; return 0;
}
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline double MPI_Wtime(void) {
  struct timeval tv; gettimeofday(&tv, 0);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}
static inline int MPI_Error_string(int e, char *s, int *len) {
  (void)e; s[0] = 0; *len = 0; return 0;
}
static inline int MPI_Op_create(MPI_User_function *f, int commute, MPI_Op *op) {
  (void)f; (void)commute; *op = MPI_SUM; return 0;
}
static inline int MPI_Op_free(MPI_Op *op) { (void)op; return 0; }

static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)root; (void)c; return 0;
}
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t,
                                MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; a2ds_stub_copy(s, r, (size_t)n * t); return 0;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t,
                             MPI_Op op, int root, MPI_Comm c) {
  (void)op; (void)root; (void)c; a2ds_stub_copy(s, r, (size_t)n * t); return 0;
}
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r,
                                int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c; a2ds_stub_copy(s, r, (size_t)sn * st); return 0;
}
static inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r,
                             int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  (void)rn; (void)rt; (void)root; (void)c;
  a2ds_stub_copy(s, r, (size_t)sn * st); return 0;
}
static inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r,
                              const int *rc, const int *displs, MPI_Datatype rt,
                              int root, MPI_Comm c) {
  (void)rc; (void)root; (void)c;
  if (s != MPI_IN_PLACE && sn)
    memcpy((char *)r + (size_t)(displs ? displs[0] : 0) * rt, s, (size_t)sn * st);
  return 0;
}
static inline int MPI_Scatter(const void *s, int sn, MPI_Datatype st, void *r,
                              int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  (void)sn; (void)st; (void)root; (void)c;
  if (r != MPI_IN_PLACE && rn) memcpy(r, s, (size_t)rn * rt);
  return 0;
}
static inline int MPI_Scatterv(const void *s, const int *sc, const int *displs,
                               MPI_Datatype st, void *r, int rn, MPI_Datatype rt,
                               int root, MPI_Comm c) {
  (void)sc; (void)root; (void)c;
  if (r != MPI_IN_PLACE && rn)
    memcpy(r, (const char *)s + (size_t)(displs ? displs[0] : 0) * st, (size_t)rn * rt);
  return 0;
}
static inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r,
                               int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c; a2ds_stub_copy(s, r, (size_t)sn * st); return 0;
}
static inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd,
                                MPI_Datatype st, void *r, const int *rc,
                                const int *rd, MPI_Datatype rt, MPI_Comm c) {
  (void)rc; (void)c;
  if (s != MPI_IN_PLACE && sc[0])
    memcpy((char *)r + (size_t)rd[0] * rt, (const char *)s + (size_t)sd[0] * st,
           (size_t)sc[0] * st);
  return 0;
}

static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; a2ds_stub_die("MPI_Send"); return 1;
}
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) {
  (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; a2ds_stub_die("MPI_Recv"); return 1;
}
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) {
  (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; a2ds_stub_die("MPI_Isend"); return 1;
}
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r) {
  (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)r; a2ds_stub_die("MPI_Irecv"); return 1;
}
static inline int MPI_Probe(int s, int tag, MPI_Comm c, MPI_Status *st) {
  (void)s; (void)tag; (void)c; (void)st; a2ds_stub_die("MPI_Probe"); return 1;
}
static inline int MPI_Get_count(const MPI_Status *s, MPI_Datatype t, int *n) {
  (void)s; (void)t; *n = 0; return 0;
}
static inline int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void)r; (void)s; return 0; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void)n; (void)r; (void)s; return 0; }
static inline int MPI_Waitany(int n, MPI_Request *r, int *idx, MPI_Status *s) {
  (void)n; (void)r; (void)s; *idx = MPI_UNDEFINED; return 0;
}

/* MPI-IO: fail softly (the .f5 / vector dump writers check the return code) */
static inline int MPI_File_open(MPI_Comm c, const char *fn, int mode, MPI_Info i, MPI_File *f) {
  (void)c; (void)fn; (void)mode; (void)i; *f = 0; return 1;
}
static inline int MPI_File_close(MPI_File *f) { (void)f; return 0; }
static inline int MPI_File_set_view(MPI_File f, MPI_Offset o, MPI_Datatype e, MPI_Datatype ft, const char *rep, MPI_Info i) {
  (void)f; (void)o; (void)e; (void)ft; (void)rep; (void)i; return 1;
}
static inline int MPI_File_set_size(MPI_File f, MPI_Offset s) { (void)f; (void)s; return 1; }
static inline int MPI_File_write(MPI_File f, const void *b, int n, MPI_Datatype t, MPI_Status *s) {
  (void)f; (void)b; (void)n; (void)t; (void)s; return 1;
}
static inline int MPI_File_read(MPI_File f, void *b, int n, MPI_Datatype t, MPI_Status *s) {
  (void)f; (void)b; (void)n; (void)t; (void)s; return 1;
}
static inline int MPI_File_write_at_all(MPI_File f, MPI_Offset o, const void *b, int n, MPI_Datatype t, MPI_Status *s) {
  (void)f; (void)o; (void)b; (void)n; (void)t; (void)s; return 1;
}
static inline int MPI_File_read_at_all(MPI_File f, MPI_Offset o, void *b, int n, MPI_Datatype t, MPI_Status *s) {
  (void)f; (void)o; (void)b; (void)n; (void)t; (void)s; return 1;
}

#ifdef __cplusplus
}
#endif
#endif
