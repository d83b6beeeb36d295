/* mpi_b200.c -- see mpi.h.  One shared-memory segment per job (created by ab2run, name in AB2_MPI_SHM):
 *   header | per-communicator barrier cells | per-rank contribution slots | per-(src,dst) mailboxes
 * A collective = [barrier] every member writes its slot [barrier] every member reads the members' slots and combines
 * them in rank order [barrier].  Without AB2_MPI_SHM the process is rank 0 of 1 and nothing is mapped. */
#include "mpi.h"
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <unistd.h>

#define MAXR 64
#define MAXC 64
#define SLOT (1 << 20)
#define MBOX 4096

typedef struct { volatile int count, gen; char pad[56]; } Bar;
typedef struct { volatile int full; int tag, bytes; char data[MBOX]; } Mail;
typedef struct {
  volatile int abort_code;      /* != 0: some rank called MPI_Abort */
  int size;
  Bar bar[MAXC];
  Mail mail[MAXR][MAXR];        /* [src][dst] */
  char slot[MAXR][SLOT];
} Shm;

typedef struct { int n, ranks[MAXR], me; } Comm;     /* world ranks of the members, my index among them (-1: not a member) */

static Shm *shm = NULL;
static int wrank = 0, wsize = 1;
static Comm comms[MAXC];
static int ncomm = 0;
static Comm groups[2*MAXC];
static int ngroup = 0;

static void die_if_aborted(void)
{
  if (shm && shm->abort_code) _exit(shm->abort_code);
}

static void barrier(int c)
{
  Comm *C = &comms[c];
  Bar *b;
  int gen;
  if (!shm || C->n <= 1) return;
  b = &shm->bar[c];
  gen = b->gen;
  if (__sync_add_and_fetch(&b->count, 1) == C->n) { b->count = 0; __sync_synchronize(); b->gen = gen + 1; }
  else { int spins = 0; while (b->gen == gen) { die_if_aborted(); if (++spins > 200) { sched_yield(); } } }
  __sync_synchronize();
}

int MPI_Init(int *argc, char ***argv)
{
  const char *name = getenv("AB2_MPI_SHM");
  (void)argc; (void)argv;
  comms[0].n = 1; comms[0].ranks[0] = 0; comms[0].me = 0; ncomm = 1;
  if (name && getenv("AB2_MPI_RANK") && getenv("AB2_MPI_SIZE")) {
    int fd, r;
    wrank = atoi(getenv("AB2_MPI_RANK")); wsize = atoi(getenv("AB2_MPI_SIZE"));
    if (wsize < 1 || wsize > MAXR || wrank < 0 || wrank >= wsize) { fprintf(stderr, "[mpi_b200]: bad AB2_MPI_RANK/SIZE\n"); return 1; }
    fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) { perror("[mpi_b200]: shm_open"); return 1; }
    shm = (Shm*)mmap(NULL, sizeof(Shm), PROT_READ|PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (shm == MAP_FAILED) { perror("[mpi_b200]: mmap"); shm = NULL; return 1; }
    comms[0].n = wsize; comms[0].me = wrank;
    for (r=0; r<wsize; r++) comms[0].ranks[r] = r;
  }
  return MPI_SUCCESS;
}

int MPI_Finalize(void) { barrier(0); return MPI_SUCCESS; }

int MPI_Abort(MPI_Comm comm, int code)
{
  (void)comm;
  if (shm) shm->abort_code = code ? code : 1;
  _exit(code ? code : 1);
}

double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + 1e-6*tv.tv_usec; }
int MPI_Get_processor_name(char *name, int *len) { gethostname(name, MPI_MAX_PROCESSOR_NAME); name[MPI_MAX_PROCESSOR_NAME-1] = 0; *len = (int)strlen(name); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { *rank = comms[c].me; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { *size = comms[c].n; return MPI_SUCCESS; }

int MPI_Comm_group(MPI_Comm c, MPI_Group *g)
{
  groups[ngroup] = comms[c];
  *g = ngroup++;
  return MPI_SUCCESS;
}

int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *out)
{ /* ranks index the members of g */
  int r;
  Comm *N = &groups[ngroup];
  if (ngroup >= 2*MAXC) { fprintf(stderr, "[mpi_b200]: too many groups\n"); return 1; }
  N->n = n; N->me = -1;
  for (r=0; r<n; r++) { N->ranks[r] = groups[g].ranks[ranks[r]]; if (N->ranks[r] == wrank) N->me = r; }
  *out = ngroup++;
  return MPI_SUCCESS;
}

int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *out)
{ /* collective over c: every rank of c calls it in the same order, so the new communicator gets the same index everywhere */
  (void)c;
  if (ncomm >= MAXC) { fprintf(stderr, "[mpi_b200]: too many communicators\n"); return 1; }
  comms[ncomm] = groups[g];
  *out = (groups[g].me >= 0) ? ncomm : MPI_COMM_NULL;
  ncomm++;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm c) { barrier(c); return MPI_SUCCESS; }

static int tsize(MPI_Datatype t) { return t & 0xff; }

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm c)
{
  Comm *C = &comms[c];
  long nb = (long)count*tsize(type), off;
  if (!shm || C->n <= 1) return MPI_SUCCESS;
  for (off = 0; off < nb; off += SLOT) {
    long n = (nb - off < SLOT) ? nb - off : SLOT;
    barrier(c);
    if (C->me == root) memcpy(shm->slot[C->ranks[root]], (char*)buf + off, n);
    barrier(c);
    if (C->me != root) memcpy((char*)buf + off, shm->slot[C->ranks[root]], n);
  }
  barrier(c);
  return MPI_SUCCESS;
}

static void combine(void *acc, const void *in, int count, MPI_Datatype type, MPI_Op op)
{
  int i;
#define LOOP(T) { T *a = (T*)acc; const T *b = (const T*)in; for (i=0; i<count; i++) { \
    if (op == MPI_SUM) a[i] += b[i]; else if (op == MPI_MAX) { if (b[i] > a[i]) a[i] = b[i]; } else { if (b[i] < a[i]) a[i] = b[i]; } } }
  switch (type) {
    case MPI_INT: LOOP(int) break;
    case MPI_LONG: LOOP(long) break;
    case MPI_FLOAT: LOOP(float) break;
    case MPI_DOUBLE: LOOP(double) break;
    default: fprintf(stderr, "[mpi_b200]: reduction on an unsupported datatype\n"); MPI_Abort(0, 1);
  }
#undef LOOP
}

static int reduce_all(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm c, int root)
{
  Comm *C = &comms[c];
  long nb = (long)count*tsize(type);
  int r;
  if (!shm || C->n <= 1) { if (rbuf != sbuf) memcpy(rbuf, sbuf, nb); return MPI_SUCCESS; }
  if (nb > SLOT) { fprintf(stderr, "[mpi_b200]: reduction of %ld bytes exceeds the slot size\n", nb); MPI_Abort(0, 1); }
  barrier(c);
  memcpy(shm->slot[wrank], sbuf, nb);
  barrier(c);
  if (root < 0 || C->me == root) {
    memcpy(rbuf, shm->slot[C->ranks[0]], nb);                   /* rank order: every rank forms the same sum */
    for (r=1; r<C->n; r++) combine(rbuf, shm->slot[C->ranks[r]], count, type, op);
  }
  barrier(c);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm c)
{ return reduce_all(sbuf, rbuf, count, type, op, c, -1); }
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm c)
{ return reduce_all(sbuf, rbuf, count, type, op, c, root); }

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm c)
{
  Mail *m;
  long nb = (long)count*tsize(type);
  if (!shm) return 1;
  if (nb > MBOX) { fprintf(stderr, "[mpi_b200]: MPI_Send of %ld bytes exceeds the mailbox\n", nb); MPI_Abort(0, 1); }
  m = &shm->mail[wrank][comms[c].ranks[dest]];
  while (m->full) { die_if_aborted(); sched_yield(); }
  memcpy(m->data, buf, nb); m->tag = tag; m->bytes = (int)nb;
  __sync_synchronize();
  m->full = 1;
  return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm c, MPI_Status *st)
{
  Mail *m;
  long nb = (long)count*tsize(type);
  if (!shm) return 1;
  m = &shm->mail[comms[c].ranks[src]][wrank];
  while (!m->full) { die_if_aborted(); sched_yield(); }
  __sync_synchronize();
  if (tag != MPI_ANY_TAG && m->tag != tag) { fprintf(stderr, "[mpi_b200]: MPI_Recv tag mismatch (%d vs %d)\n", m->tag, tag); MPI_Abort(0, 1); }
  memcpy(buf, m->data, (m->bytes < nb) ? m->bytes : nb);
  if (st) { st->MPI_SOURCE = src; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; }
  __sync_synchronize();
  m->full = 0;
  return MPI_SUCCESS;
}

/* what ab2run maps: exported so the launcher and the library agree on the size */
unsigned long ab2_mpi_shm_bytes(void) { return sizeof(Shm); }
