/* Single-rank MPI stand-in used ONLY to build the unmodified reference as a
 * test oracle (oracle/_ref).  TEST INFRASTRUCTURE - never linked into the product.
 *
 * Valid because every hypercube loop in the reference is
 * `for(level=1; level<NTask; ...)` (gravtree.c:171, sidm.c:204) and the ORB loop is
 * `for(level=NTask; level>1; ...)` (domain.c:91): with NTask==1 no point-to-point
 * call is ever reached and every collective degenerates to a local copy.
 */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
#include <string.h>
#include <stdlib.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_BYTE   1
#define MPI_INT    4
#define MPI_FLOAT  5
#define MPI_DOUBLE 8
#define MPI_SUM 0
#define MPI_MIN 1
#define MPI_SUCCESS 0

static inline int stubmpi_size(MPI_Datatype t)
{ return t == MPI_BYTE ? 1 : (t == MPI_DOUBLE ? 8 : 4); }

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code ? code : 1); return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline double MPI_Wtime(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; if (s != r) memmove(r, s, (size_t)n * stubmpi_size(t)); return 0; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)op; (void)c; (void)root; if (s != r) memmove(r, s, (size_t)n * stubmpi_size(t)); return 0; }
static inline int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{ (void)nr; (void)tr; (void)c; if (s != r) memmove(r, s, (size_t)ns * stubmpi_size(ts)); return 0; }
static inline int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c)
{ (void)nr; (void)tr; (void)c; (void)root; if (s != r) memmove(r, s, (size_t)ns * stubmpi_size(ts)); return 0; }
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }

/* point-to-point: unreachable with one rank; abort loudly if ever called */
static inline int stubmpi_p2p(void) { abort(); return 1; }
#define MPI_Send(...)     stubmpi_p2p()
#define MPI_Ssend(...)    stubmpi_p2p()
#define MPI_Recv(...)     stubmpi_p2p()
#define MPI_Sendrecv(...) stubmpi_p2p()
#endif
