/* Single-rank MPI stand-in -- see mpi.h.  TEST INFRASTRUCTURE ONLY. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mpi.h"

struct mpi_stub_request {
  void *buf;            /* receive buffer (NULL for a send request) */
  long bytes;
  int comm, tag, done;
  struct mpi_stub_request *next;
};
typedef struct msg { void *data; long bytes; int comm, tag; struct msg *next; } msg;

static struct mpi_stub_request *posted_head = NULL, *posted_tail = NULL;   /* receives waiting for a send */
static msg *unexp_head = NULL, *unexp_tail = NULL;                         /* sends waiting for a receive */
static int next_comm = 2;
static void *attached = NULL; static int attached_size = 0;

static void die(const char *what)
{
  fprintf(stderr, "### mpi_stub: %s\n", what);
  exit(3);
}
static long tsize(MPI_Datatype t) { return t & 0xff; }
static int tag_match(int want, int have) { return want == MPI_ANY_TAG || want == have; }

static void check_self(int peer, const char *fn)
{
  if (peer != 0 && peer != MPI_ANY_SOURCE) {
    fprintf(stderr, "### mpi_stub: %s to/from rank %d in a one-rank run\n", fn, peer);
    exit(3);
  }
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) { (void)comm; exit(code ? code : 1); }
double MPI_Wtime(void)
{
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9*(double)ts.tv_nsec;
}
int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "stub"); *len = 4; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) { (void)comm; *group = 1; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup)
{
  int i, has0 = 0;
  (void)group;
  for (i=0; i<n; i++) { if (ranks[i] == 0) has0 = 1; else die("MPI_Group_incl with a rank other than 0"); }
  *newgroup = has0 ? 1 : 0;
  return MPI_SUCCESS;
}
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm)
{
  (void)comm;
  *newcomm = group ? next_comm++ : MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm comm) { (void)comm; return MPI_SUCCESS; }
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{ (void)buf; (void)count; (void)type; (void)comm; check_self(root, "MPI_Bcast"); return MPI_SUCCESS; }
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{ (void)op; (void)comm; if (sbuf != rbuf) memmove(rbuf, sbuf, (size_t)(count*tsize(type))); return MPI_SUCCESS; }
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm)
{ check_self(root, "MPI_Reduce"); return MPI_Allreduce(sbuf, rbuf, count, type, op, comm); }

static void deliver(const void *data, long bytes, int comm, int tag)
{
  struct mpi_stub_request *r = posted_head, *prev = NULL;
  for (; r; prev = r, r = r->next) {
    if (r->comm == comm && tag_match(r->tag, tag)) {
      if (bytes > r->bytes) die("message longer than the posted receive");
      memcpy(r->buf, data, (size_t)bytes);
      r->bytes = bytes; r->tag = tag; r->done = 1;
      if (prev) prev->next = r->next; else posted_head = r->next;
      if (posted_tail == r) posted_tail = prev;
      r->next = NULL;
      return;
    }
  }
  { msg *m = (msg*)malloc(sizeof(msg));
    m->data = malloc((size_t)(bytes > 0 ? bytes : 1)); memcpy(m->data, data, (size_t)bytes);
    m->bytes = bytes; m->comm = comm; m->tag = tag; m->next = NULL;
    if (unexp_tail) unexp_tail->next = m; else unexp_head = m;
    unexp_tail = m; }
}

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{ check_self(dest, "MPI_Send"); deliver(buf, count*tsize(type), comm, tag); return MPI_SUCCESS; }
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
  struct mpi_stub_request *r;
  check_self(dest, "MPI_Isend");
  deliver(buf, count*tsize(type), comm, tag);
  r = (struct mpi_stub_request*)calloc(1, sizeof(*r)); r->done = 1; r->comm = comm; r->tag = tag;
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Ibsend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{ return MPI_Isend(buf, count, type, dest, tag, comm, req); }

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Request *req)
{
  struct mpi_stub_request *r = (struct mpi_stub_request*)calloc(1, sizeof(*r));
  msg *m = unexp_head, *prev = NULL;
  check_self(src, "MPI_Irecv");
  r->buf = buf; r->bytes = count*tsize(type); r->comm = comm; r->tag = tag;
  *req = r;
  for (; m; prev = m, m = m->next) {
    if (m->comm == comm && tag_match(tag, m->tag)) {
      if (m->bytes > r->bytes) die("queued message longer than the receive");
      memcpy(buf, m->data, (size_t)m->bytes);
      r->bytes = m->bytes; r->tag = m->tag; r->done = 1;
      if (prev) prev->next = m->next; else unexp_head = m->next;
      if (unexp_tail == m) unexp_tail = prev;
      free(m->data); free(m);
      return MPI_SUCCESS;
    }
  }
  if (posted_tail) posted_tail->next = r; else posted_head = r;
  posted_tail = r;
  return MPI_SUCCESS;
}

static void fill_status(MPI_Status *st, const struct mpi_stub_request *r)
{ if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = r->tag; st->MPI_ERROR = MPI_SUCCESS; st->stub_bytes = (int)r->bytes; } }

int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st)
{
  if (*req == MPI_REQUEST_NULL) { *flag = 1; return MPI_SUCCESS; }
  *flag = (*req)->done;
  if (*flag) { fill_status(st, *req); free(*req); *req = MPI_REQUEST_NULL; }
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *req, MPI_Status *st)
{
  int flag;
  if (*req != MPI_REQUEST_NULL && !(*req)->done) die("MPI_Wait on a receive nobody has sent to (deadlock in a one-rank run)");
  return MPI_Test(req, &flag, st);
}
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *st)
{
  MPI_Request r;
  MPI_Irecv(buf, count, type, src, tag, comm, &r);
  return MPI_Wait(&r, st);
}
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts)
{
  int i;
  for (i=0; i<n; i++) MPI_Wait(&reqs[i], sts ? &sts[i] : NULL);
  return MPI_SUCCESS;
}
int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st)
{
  int i, active = 0, flag;
  for (i=0; i<n; i++) {
    if (reqs[i] == MPI_REQUEST_NULL) continue;
    active = 1;
    if (reqs[i]->done) { *index = i; return MPI_Test(&reqs[i], &flag, st); }
  }
  if (active) die("MPI_Waitany: no request can complete (deadlock in a one-rank run)");
  *index = MPI_UNDEFINED;
  return MPI_SUCCESS;
}
int MPI_Pack_size(int count, MPI_Datatype type, MPI_Comm comm, int *size) { (void)comm; *size = (int)(count*tsize(type)); return MPI_SUCCESS; }
int MPI_Buffer_attach(void *buf, int size) { attached = buf; attached_size = size; return MPI_SUCCESS; }
int MPI_Buffer_detach(void *buf, int *size) { *(void**)buf = attached; *size = attached_size; attached = NULL; attached_size = 0; return MPI_SUCCESS; }
