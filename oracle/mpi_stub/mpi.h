/* mpi.h -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
 *
 * An in-process stand-in for the handful of MPI calls the reference makes, so that its communication
 * path -- src/comms.c:5-196 (fast_transfer_boundary_fluxes), src/init.c:162-225 (init_mpi_grid) and the
 * scalar reductions of src/solver.c:1186-1196, 1391-1420 -- can be compiled UNMODIFIED with -DMPI in an
 * image that has no MPI and executed here: one pthread per rank, messages through an
 * in-memory mailbox (oracle/mpi_stub.c).  Only what the reference uses is declared.
 *
 * Semantics kept from the MPI standard (what the reference relies on):
 *   - MPI_Cart_create on a non-periodic grid numbers ranks row-major (last dimension fastest);
 *     MPI_Cart_shift(dim, disp) returns (rank_source, rank_dest) = the neighbours at -disp / +disp, or
 *     MPI_PROC_NULL across a non-periodic border;
 *   - MPI_PROC_NULL is -1, MPICH's value: comms.c:118,146 compares neighbours with the literal -1
 *     (SURVEY F9);
 *   - point-to-point messages between a pair of ranks with the same tag are delivered in the order they
 *     were sent; a send buffer may be reused once MPI_Wait returns (the stub copies at MPI_Isend: the
 *     "eager" protocol every MPI implementation may choose);
 *   - MPI_Reduce / MPI_Allreduce with MPI_SUM add the contributions in rank order (the standard leaves
 *     the order to the implementation).
 */
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* (kind << 24) | bytes per element */
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef struct mpi_stub_request *MPI_Request;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_PROC_NULL (-1)
#define MPI_FLOAT  ((1 << 24) | 4)
#define MPI_LONG   ((2 << 24) | 8)
#define MPI_INT    ((3 << 24) | 4)
#define MPI_DOUBLE ((4 << 24) | 8)
#define MPI_SUM 1

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *cart);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
double MPI_Wtime(void);

/* the harness side (oracle/ref_harness.c): how many ranks there are and which one the calling thread is */
void mpi_stub_world(int nranks);
void mpi_stub_set_rank(int rank);
int mpi_stub_run(int nranks, void (*fn)(int rank, void *arg), void *arg);   /* fn(rank) on one thread per rank */

#endif
