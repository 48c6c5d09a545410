/* candmc/mpi.h — the slice of the MPI API that the reference's CANMM drivers and communicator macros use
 * (alg/shared/comm.h:110-201, test/MM/topo_pdgemm_unit.cxx, test/MM/test_spc.cxx, bench/MM/topo_pdgemm_bench.cxx,
 * bench/MM/bench_spc.cxx), implemented by libcandmc_b200.so on top of its NCCL grid communicators.
 *
 * An MPI_Comm here IS a candmc_comm_t*, so `CommData_t.cm` can be handed straight to the C ABI.  One process per GPU:
 * MPI_Init binds the process to GPU $LOCAL_RANK and joins the world communicator of $WORLD_SIZE ranks; the NCCL id
 * travels through the directory $CANDMC_RENDEZVOUS (tools/candmc_run sets all of these).  Host-buffer collectives
 * (MPI_Bcast, MPI_Reduce, MPI_Allreduce, MPI_Barrier) are staged through device memory; they exist for the drivers'
 * bookkeeping (pass flags, input replication), the hot path uses the C ABI with device pointers.
 * Only what those files need is provided — this is not a general MPI.
 */
#ifndef CANDMC_MPI_COMPAT_H
#define CANDMC_MPI_COMPAT_H

#include "../candmc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef candmc_comm_t* MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Request;
typedef struct MPI_Status {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD (candmc_mpi_comm_world())
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)-1)

/* datatype = size in bytes | kind << 8 */
#define MPI_CHAR 0x101
#define MPI_BYTE 0x201
#define MPI_INT 0x304
#define MPI_DOUBLE 0x408
#define MPI_INT64_T 0x508
#define MPI_LONG 0x608
#define MPI_FLOAT 0x704

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_BAND 4
#define MPI_BOR 5

MPI_Comm candmc_mpi_comm_world(void);
int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif
