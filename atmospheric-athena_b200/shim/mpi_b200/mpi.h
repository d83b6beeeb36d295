/* mpi.h -- the rank layer the drop-in's MPI build compiles against when no MPI is installed: the subset of MPI-1 the
 * reference's DRIVER calls (main.c, par.c, init_mesh.c, output*.c, dump_history.c, utils.c, ath_signal.c, baton.c),
 * for the processes of ONE node, one per GPU, started by `ab2run -np N` (ab2run.c).
 *
 * It carries control-plane traffic only -- the athinput broadcast (par.c:par_dist_mpi), rank/size, the Domain
 * communicators of init_mesh.c, error and zone-cycle reductions, the history sums.  The data plane of the hot path (ghost
 * zones, timestep and sub-cycle reductions) never goes through it: that is NCCL inside libathena_b200 (csrc/comm.cu),
 * whose 128-byte unique id is the one thing athena_shim.c sends through MPI_Bcast.  With a real MPI installed, compile
 * the driver with mpicc instead and drop this directory; the shim uses nothing beyond MPI_Bcast / MPI_Comm_rank / size.
 *
 * Collectives run through a shared-memory segment (mpi_b200.c): contributions in per-rank slots between sense-reversing
 * barriers.  Only the calls, datatypes and operations the driver uses exist.
 */
#ifndef AB2_MPI_H
#define AB2_MPI_H

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL (-1)
#define MPI_COMM_WORLD 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_TAG (-1)
#define MPI_MAX_PROCESSOR_NAME 256

/* datatype = (kind << 8) | size in bytes */
#define MPI_CHAR   ((1<<8)|1)
#define MPI_BYTE   ((2<<8)|1)
#define MPI_INT    ((3<<8)|4)
#define MPI_FLOAT  ((4<<8)|4)
#define MPI_LONG   ((5<<8)|8)
#define MPI_DOUBLE ((6<<8)|8)

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
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *st);

#endif
