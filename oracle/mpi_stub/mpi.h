/* Single-rank MPI stand-in.  TEST INFRASTRUCTURE ONLY (oracle scaffolding, SURVEY.md 8c option ii).
 *
 * The reference's SMR + ionising-radiation hand-down (ionradiation/ionrad_smr.c) only works below level 1 in its
 * MPI branch -- the branch its published runs used; the serial branch indexes out of bounds.  There is
 * no MPI in this image, so oracle/build_ref.py compiles the UNMODIFIED reference with -DMPI_PARALLEL against this
 * header and mpi_stub.c: one process, rank 0 of 1, messages to self delivered by tag matching.  Only the calls the
 * reference makes are provided.  Nothing in the product includes or links this.
 */
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, stub_bytes; } MPI_Status;
typedef struct mpi_stub_request *MPI_Request;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-2)
#define MPI_MAX_PROCESSOR_NAME 256

/* datatypes: the value is the size in bytes plus a tag in the high bits */
#define MPI_CHAR   (0x100|1)
#define MPI_BYTE   (0x200|1)
#define MPI_INT    (0x300|4)
#define MPI_FLOAT  (0x400|4)
#define MPI_LONG   (0x500|8)
#define MPI_DOUBLE (0x600|8)

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
int MPI_Get_processor_name(char *name, int *len);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Ibsend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st);
int MPI_Pack_size(int count, MPI_Datatype type, MPI_Comm comm, int *size);
int MPI_Buffer_attach(void *buf, int size);
int MPI_Buffer_detach(void *buf, int *size);

#endif
