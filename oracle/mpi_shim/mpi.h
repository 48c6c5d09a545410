/* mini-MPI — TEST INFRASTRUCTURE ONLY (part of oracle/).
 *
 * The image has no MPI (no mpicxx / mpirun / mpi.h / libmpi), so the unmodified reference CANMM sources
 * (/root/reference/alg/MM, alg/shared, test/MM, bench/MM) are compiled against this header and linked with
 * mpi_shim.c: ranks are forked processes that talk through POSIX shared-memory rings.  It implements exactly the
 * MPI surface the MM path touches and nothing else.  The product library never includes or links
 * this file.
 */
#ifndef MINI_MPI_H
#define MINI_MPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Win;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)

/* datatypes: value = size in bytes | (kind << 8) */
#define MPI_CHAR 0x101
#define MPI_BYTE 0x201
#define MPI_INT 0x304
#define MPI_DOUBLE 0x408
#define MPI_INT64_T 0x508
#define MPI_LONG 0x608
#define MPI_FLOAT 0x704
#define MPI_UNSIGNED 0x804
#define MPI_LONG_LONG 0x908
#define MPI_LONG_LONG_INT 0x908
#define MPI_DOUBLE_COMPLEX 0xa10

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_BAND 4
#define MPI_BOR 5
#define MPI_LAND 6
#define MPI_LOR 7
#define MPI_PROD 8

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
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* status);
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* statuses);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status);
int MPI_Sendrecv_replace(void* buf, int count, MPI_Datatype type, int dest, int sendtag, int source, int recvtag,
                         MPI_Comm comm, MPI_Status* status);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                const int* displs, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Win_create(void* base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win* win);
int MPI_Win_fence(int assert_, MPI_Win win);
int MPI_Win_free(MPI_Win* win);
int MPI_Put(const void* origin, int origin_count, MPI_Datatype origin_type, int target_rank, MPI_Aint target_disp,
            int target_count, MPI_Datatype target_type, MPI_Win win);

/* the reference's profiler (alg/shared/timer.cxx:35-42,206-227) references these unconditionally */
int PMPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int PMPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int PMPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int PMPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int PMPI_Barrier(MPI_Comm comm);
int PMPI_Comm_rank(MPI_Comm comm, int* rank);
int PMPI_Comm_size(MPI_Comm comm, int* size);
double PMPI_Wtime(void);

#ifdef __cplusplus
}
#endif
#endif
