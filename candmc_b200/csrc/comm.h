// candmc_b200 — processor-grid communicators on NCCL (internal header).
// A candmc_comm stands where the reference has CommData_t.cm (an MPI_Comm, alg/shared/comm.h:32-37): one
// communicator per grid axis, created by splitting the world communicator with MPI_Comm_split semantics.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include "common.cuh"

struct candmc_comm {
  ncclComm_t nccl = nullptr;      // full-width communicator: exposed collectives (depth all-reduce, host bookkeeping)
  ncclComm_t nccl_bg = nullptr;   // same ranks, few CTAs: panel traffic that runs UNDER a GEMM (created on first use)
  int rank = 0;
  int size = 1;
  void* fused_ctx = nullptr;      // candmc::FusedCtx* of the fused GEMM + depth all-reduce (ipc.h), owned by this handle
  bool fused_failed = false;      // CUDA IPC unavailable: stay on ncclAllReduce
  void* transport = nullptr;      // candmc::PanelTransport* (transport.h): copy-engine panel transport over peer windows
  void* p2p = nullptr;            // candmc::P2PTransport* (transport.h): the same for point-to-point exchanges
  bool transport_failed = false;  // peer windows / stream memory operations unavailable: stay on ncclBroadcast
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

// Communicator for traffic overlapped with compute: NCCL moves data with SM-resident kernels, and a persistent GEMM
// that owns every SM loses ~10 % when a full-width NCCL kernel co-runs (measured, profiles/).  A second communicator
// capped at `runtime().bg_max_ctas` CTAs keeps the links busy enough (panels are ~8x smaller than the GEMM they hide
// under) while holding only a handful of SMs.  Collective over the communicator on first use.
int comm_background(candmc_comm* c, ncclComm_t* out);

// thin wrappers, device pointers only, size-1 communicators short-circuit (no NCCL call)
int comm_bcast(candmc_comm* c, const double* send, double* recv, int64_t count, int root, cudaStream_t st,
               bool background = false);
// Do all ranks of `c` hold the same 0/1 `flag`?  One 8-byte all-reduce on `st` and a synchronisation of `st`; collective.
int comm_flags_agree(candmc_comm* c, int flag, bool* agree, cudaStream_t st);
int comm_allreduce(candmc_comm* c, const double* send, double* recv, int64_t count, cudaStream_t st,
                   bool background = false);
// grouped point-to-point exchange: send `scount` doubles to `dst`, receive `rcount` from `src` (either may be
// skipped with a negative peer); self-exchange degenerates to a device copy.
int comm_sendrecv(candmc_comm* c, const double* send, int64_t scount, int dst, double* recv, int64_t rcount, int src,
                  cudaStream_t st, bool background = false);

}  // namespace candmc
