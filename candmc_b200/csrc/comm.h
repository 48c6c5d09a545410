// candmc_b200 — processor-grid communicators on NCCL (internal header).
// A candmc_comm stands where the reference has CommData_t.cm (an MPI_Comm, alg/shared/comm.h:32-37): one
// communicator per grid axis, created by splitting the world communicator with MPI_Comm_split semantics.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include "common.cuh"

struct candmc_comm {
  ncclComm_t nccl = nullptr;
  int rank = 0;
  int size = 1;
};

namespace candmc {

#define CANDMC_NCCL(call)                                                                          \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) {                                                                      \
      ::candmc::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
      return ::candmc::ERR_NCCL;                                                                   \
    }                                                                                              \
  } while (0)

// thin wrappers, device pointers only, size-1 communicators short-circuit (no NCCL call)
int comm_bcast(candmc_comm* c, const double* send, double* recv, int64_t count, int root, cudaStream_t st);
int comm_allreduce(candmc_comm* c, const double* send, double* recv, int64_t count, cudaStream_t st);
// grouped point-to-point exchange: send `scount` doubles to `dst`, receive `rcount` from `src` (either may be
// skipped with a negative peer); self-exchange degenerates to a device copy.
int comm_sendrecv(candmc_comm* c, const double* send, int64_t scount, int dst, double* recv, int64_t rcount, int src,
                  cudaStream_t st);

}  // namespace candmc
