/* mini-MPI (see mpi.h): fork + shared-memory mailboxes, one mailbox per ordered pair of ranks, messages cut into chunks.
 * Blocking semantics of the calls the reference makes; per-pair FIFO order; tags are carried but matched in order (the
 * reference posts its receives in the order of the sends).  TEST INFRASTRUCTURE. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <signal.h>
#include <sched.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include "mpi.h"

#define MM_MAXP  64
#define MM_CHUNK (1 << 18)

struct mbox { volatile int full; int tag; size_t total, len; char data[MM_CHUNK]; };
struct shared { volatile int arrived, sense; volatile int abort_code; struct mbox box[MM_MAXP][MM_MAXP]; };

static struct shared *S;
static int mm_rank = 0, mm_size = 1;
static pid_t mm_child[MM_MAXP];

static size_t tsize(MPI_Datatype t) { return t == MPI_BYTE ? 1 : (t == MPI_DOUBLE ? 8 : 4); }
static void spin(void) { if (S->abort_code) _exit(S->abort_code); sched_yield(); }

int MPI_Init(int *argc, char ***argv)
{
  const char *e = getenv("MINIMPI_NP");
  int np = e ? atoi(e) : 1, r;
  (void)argc; (void)argv;
  if (np < 1 || np > MM_MAXP) { fprintf(stderr, "minimpi: MINIMPI_NP must be 1..%d\n", MM_MAXP); exit(2); }
  S = (struct shared *)mmap(0, sizeof(struct shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (S == MAP_FAILED) { perror("minimpi: mmap"); exit(2); }
  memset((void *)S, 0, sizeof(int) * 4);
  for (int a = 0; a < MM_MAXP; a++) for (int b = 0; b < MM_MAXP; b++) S->box[a][b].full = 0;
  mm_size = np; mm_rank = 0;
  fflush(stdout); fflush(stderr);
  for (r = 1; r < np; r++) {
    pid_t p = fork();
    if (p < 0) { perror("minimpi: fork"); exit(2); }
    if (p == 0) { mm_rank = r; prctl(PR_SET_PDEATHSIG, SIGKILL); break; }
    mm_child[r] = p;
  }
  return 0;
}

int MPI_Finalize(void)
{
  MPI_Barrier(MPI_COMM_WORLD);
  fflush(stdout); fflush(stderr);
  if (mm_rank == 0) for (int r = 1; r < mm_size; r++) { int st; waitpid(mm_child[r], &st, 0); }
  else _exit(0);                       /* children leave here: the parent's exit code is the job's */
  return 0;
}
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = mm_rank; return 0; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = mm_size; return 0; }
int MPI_Abort(MPI_Comm c, int code)
{
  (void)c;
  fflush(stdout); fflush(stderr);
  if (S) S->abort_code = code ? code : 1;
  _exit(code ? code : 1);
  return 0;
}
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int MPI_Barrier(MPI_Comm c)
{
  (void)c;
  if (mm_size == 1) return 0;
  const int sense = S->sense;
  if (__sync_add_and_fetch(&S->arrived, 1) == mm_size) { S->arrived = 0; __sync_synchronize(); S->sense = !sense; }
  else while (S->sense == sense) spin();
  return 0;
}

/* one step of a send / a receive; return 1 when the whole message has gone */
struct xfer { const char *sbuf; char *rbuf; size_t total, done; int peer, tag, started; };
static int push(struct xfer *x)
{
  struct mbox *b = &S->box[mm_rank][x->peer];
  if (x->started && x->done >= x->total) return 1;
  if (b->full) return 0;
  size_t len = x->total - x->done; if (len > MM_CHUNK) len = MM_CHUNK;
  memcpy(b->data, x->sbuf + x->done, len);
  b->tag = x->tag; b->total = x->total; b->len = len;
  __sync_synchronize();
  b->full = 1;
  x->done += len; x->started = 1;
  return x->done >= x->total;
}
static int pull(struct xfer *x, size_t cap)
{
  struct mbox *b = &S->box[x->peer][mm_rank];
  if (x->started && x->done >= x->total) return 1;
  if (!b->full) return 0;
  __sync_synchronize();
  if (!x->started) { x->total = b->total; x->started = 1; if (x->total > cap) { fprintf(stderr, "minimpi: rank %d: message of %zu bytes from %d exceeds the receive buffer (%zu)\n", mm_rank, x->total, x->peer, cap); MPI_Abort(0, 3); } }
  memcpy(x->rbuf + x->done, b->data, b->len);
  x->done += b->len;
  __sync_synchronize();
  b->full = 0;
  return x->done >= x->total;
}

static int send_bytes(const void *buf, size_t bytes, int dst, int tag)
{
  struct xfer x = {(const char *)buf, 0, bytes, 0, dst, tag, 0};
  if (dst == mm_rank) { fprintf(stderr, "minimpi: send to self\n"); MPI_Abort(0, 3); }
  while (!push(&x)) spin();
  return 0;
}
static size_t recv_bytes(void *buf, size_t cap, int src, int tag)
{
  struct xfer x = {0, (char *)buf, 0, 0, src, tag, 0};
  while (!pull(&x, cap)) spin();
  return x.total;
}

int MPI_Send(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void)c; return send_bytes(buf, (size_t)n * tsize(t), dst, tag); }
int MPI_Ssend(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void)c; return send_bytes(buf, (size_t)n * tsize(t), dst, tag); }
int MPI_Recv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st)
{
  (void)c;
  recv_bytes(buf, (size_t)n * tsize(t), src, tag);
  if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
  return 0;
}
int MPI_Sendrecv(const void *sbuf, int ns, MPI_Datatype ts, int dst, int stag, void *rbuf, int nr, MPI_Datatype tr, int src, int rtag,
                 MPI_Comm c, MPI_Status *st)
{
  (void)c;
  if (dst == mm_rank && src == mm_rank) { memmove(rbuf, sbuf, (size_t)ns * tsize(ts)); return 0; }
  struct xfer xs = {(const char *)sbuf, 0, (size_t)ns * tsize(ts), 0, dst, stag, 0};
  struct xfer xr = {0, (char *)rbuf, 0, 0, src, rtag, 0};
  int ds = 0, dr = 0;
  while (!ds || !dr) {                 /* both directions make progress: no deadlock on messages longer than a mailbox */
    if (!ds) ds = push(&xs);
    if (!dr) dr = pull(&xr, (size_t)nr * tsize(tr));
    if (!ds || !dr) spin();
  }
  if (st) { st->MPI_SOURCE = src; st->MPI_TAG = rtag; st->MPI_ERROR = 0; }
  return 0;
}

#define TAG_COLL 0x7fff0001
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c)
{
  (void)c;
  const size_t bytes = (size_t)n * tsize(t);
  if (mm_size == 1) return 0;
  if (mm_rank == root) { for (int r = 0; r < mm_size; r++) if (r != root) send_bytes(buf, bytes, r, TAG_COLL); }
  else recv_bytes(buf, bytes, root, TAG_COLL);
  return 0;
}
static void combine(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op)
{
#define LOOP(T) { T *a = (T *)acc; const T *b = (const T *)in; for (int i = 0; i < n; i++) a[i] = op == MPI_SUM ? a[i] + b[i] : (op == MPI_MIN ? (b[i] < a[i] ? b[i] : a[i]) : (b[i] > a[i] ? b[i] : a[i])); }
  if (t == MPI_INT) LOOP(int) else if (t == MPI_FLOAT) LOOP(float) else if (t == MPI_DOUBLE) LOOP(double) else LOOP(unsigned char)
#undef LOOP
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
  (void)c;
  const size_t bytes = (size_t)n * tsize(t);
  if (mm_rank == root) {
    char *tmp = (char *)malloc(bytes ? bytes : 1);
    if (s != r) memmove(r, s, bytes);
    for (int q = 0; q < mm_size; q++) if (q != root) { recv_bytes(tmp, bytes, q, TAG_COLL); combine(r, tmp, n, t, op); }
    free(tmp);
  } else send_bytes(s, bytes, root, TAG_COLL);
  return 0;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  if (mm_size == 1) { if (s != r) memmove(r, s, (size_t)n * tsize(t)); return 0; }
  if (mm_rank != 0) {                  /* the send buffer may alias the receive buffer only on the root */
    MPI_Reduce(s, 0, n, t, op, 0, c);
  } else MPI_Reduce(s, r, n, t, op, 0, c);
  return MPI_Bcast(r, n, t, 0, c);
}
int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c)
{
  (void)c; (void)nr; (void)tr;
  const size_t bytes = (size_t)ns * tsize(ts);
  if (mm_rank == root) {
    for (int q = 0; q < mm_size; q++) {
      if (q == root) memmove((char *)r + (size_t)q * bytes, s, bytes);
      else recv_bytes((char *)r + (size_t)q * bytes, bytes, q, TAG_COLL);
    }
  } else send_bytes(s, bytes, root, TAG_COLL);
  return 0;
}
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{
  MPI_Gather(s, ns, ts, r, nr, tr, 0, c);
  return MPI_Bcast(r, ns * mm_size, ts, 0, c);
}
