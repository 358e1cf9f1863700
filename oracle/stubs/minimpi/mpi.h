/* mini-MPI: the MPI-1 subset the reference's driver uses, for SEVERAL ranks on one box without an MPI installation
 * (SURVEY.md 8d): MPI_Init forks $MINIMPI_NP - 1 children, messages go through shared-memory mailboxes.
 * TEST INFRASTRUCTURE (oracle/): lets the tests start the reference's real main() with NTask >= 2 (main.c:39-53 refuses
 * one task) on the GPU drop-in.  A production build links the site's own MPI instead. */
#ifndef MINIMPI_H
#define MINIMPI_H
#ifdef __cplusplus
extern "C" {
#endif
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
#define MPI_MAX 2
#define MPI_SUCCESS 0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Comm_size(MPI_Comm c, int *size);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Barrier(MPI_Comm c);
double MPI_Wtime(void);
int MPI_Send(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Ssend(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Recv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Sendrecv(const void *sbuf, int ns, MPI_Datatype ts, int dst, int stag, void *rbuf, int nr, MPI_Datatype tr, int src, int rtag,
                 MPI_Comm c, MPI_Status *st);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c);
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c);
#ifdef __cplusplus
}
#endif
#endif
