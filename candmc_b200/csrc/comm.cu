// candmc_b200 — processor-grid communicators: the MPI calls of the CANMM hot path re-expressed on NCCL over
// NVLink 5 / NVSwitch.  Reference call sites: MPI_Comm_split in SETUP_SUB_COMM / RSETUP_KDIR_COMM /
// RSETUP_LAYER_COMM (alg/shared/comm.h:145-195), MPI_Bcast via POST_BCAST (comm.h:110-112; summa.cxx:63-84,
// d25_summa.cxx:126-162, dual_cannon.cxx:153-162), MPI_Allreduce (d25_summa.cxx:149,221; qr_2d.cxx:265),
// MPI_Isend/Irecv (dual_cannon.cxx:116-135,198-209), MPI_Put/Win_fence (spcannon.cxx:64-71,139-152,217-224).
#include "comm.h"

#include <string.h>

#include <algorithm>

#include "../../include/candmc_b200.h"
#include "ipc.h"
#include "transport.h"
#include "runtime.h"
#include "staging.h"

namespace candmc {

int comm_background(candmc_comm* c, ncclComm_t* out) {
  if (c->nccl_bg == nullptr) {
    if (runtime().bg_max_ctas <= 0 || c->size == 1) {
      c->nccl_bg = c->nccl;
    } else {
      // ncclCommSplit must not race with outstanding operations on the parent: the background communicator is normally
      // created together with its parent (candmc_comm_split / candmc_comm_init_rank); on this late path every rank
      // drains its device first (all ranks reach this point in the same call, so the parent's operations complete)
      CANDMC_CUDA(cudaDeviceSynchronize());
      ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
      cfg.minCTAs = 1;
      cfg.maxCTAs = runtime().bg_max_ctas;
      CANDMC_NCCL(ncclCommSplit(c->nccl, 0, c->rank, &c->nccl_bg, &cfg));
    }
  }
  *out = c->nccl_bg;
  return OK;
}

int comm_bcast(candmc_comm* c, const double* send, double* recv, int64_t count, int root, cudaStream_t st,
               bool background) {
  CANDMC_CHECK(c != nullptr, "bcast: null communicator");
  CANDMC_CHECK(root >= 0 && root < c->size, "bcast: root %d outside communicator of size %d", root, c->size);
  if (count <= 0) return OK;
  if (c->size == 1) {
    if (send != recv)
      CANDMC_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, st));
    return OK;
  }
  ncclComm_t comm = c->nccl;
  if (background) CANDMC_TRY(comm_background(c, &comm));
  CANDMC_NCCL(ncclBroadcast(send, recv, (size_t)count, ncclDouble, root, comm, st));
  return OK;
}

int comm_allreduce(candmc_comm* c, const double* send, double* recv, int64_t count, cudaStream_t st, bool background) {
  CANDMC_CHECK(c != nullptr, "allreduce: null communicator");
  if (count <= 0) return OK;
  if (c->size == 1) {
    if (send != recv)
      CANDMC_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, st));
    return OK;
  }
  ncclComm_t comm = c->nccl;
  if (background) CANDMC_TRY(comm_background(c, &comm));
  CANDMC_NCCL(ncclAllReduce(send, recv, (size_t)count, ncclDouble, ncclSum, comm, st));
  return OK;
}

int comm_flags_agree(candmc_comm* c, int flag, bool* agree, cudaStream_t st) {
  CANDMC_CHECK(c != nullptr && agree != nullptr, "flags_agree: null argument");
  *agree = true;
  if (c->size == 1) return OK;
  static double* dev = nullptr;   // one process-lifetime scratch word (calls are collective and serialised by the sync below)
  if (!dev) CANDMC_CUDA(cudaMalloc(&dev, sizeof(double)));
  double v = flag ? 1.0 : 0.0;
  CANDMC_CUDA(cudaMemcpyAsync(dev, &v, sizeof(double), cudaMemcpyHostToDevice, st));
  CANDMC_NCCL(ncclAllReduce(dev, dev, 1, ncclDouble, ncclSum, c->nccl, st));
  CANDMC_CUDA(cudaMemcpyAsync(&v, dev, sizeof(double), cudaMemcpyDeviceToHost, st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  *agree = (v == 0.0 || v == static_cast<double>(c->size));
  return OK;
}

int comm_sendrecv(candmc_comm* c, const double* send, int64_t scount, int dst, double* recv, int64_t rcount, int src,
                  cudaStream_t st, bool background) {
  CANDMC_CHECK(c != nullptr, "sendrecv: null communicator");
  const bool do_send = dst >= 0 && scount > 0, do_recv = src >= 0 && rcount > 0;
  if (do_send && do_recv && dst == c->rank && src == c->rank) {
    CANDMC_CHECK(scount == rcount, "sendrecv: self exchange with mismatched counts");
    if (send != recv)
      CANDMC_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * scount, cudaMemcpyDeviceToDevice, st));
    return OK;
  }
  CANDMC_CHECK(!(do_send && dst == c->rank) && !(do_recv && src == c->rank), "sendrecv: unmatched self transfer");
  if (!do_send && !do_recv) return OK;
  if (p2p_transport_usable(c, std::max(do_send ? scount : 0, do_recv ? rcount : 0))) {   // copy engines + flags, no NCCL kernel
    if (do_send) CANDMC_TRY(p2p_transport_send(c, send, scount, dst, st));
    if (do_recv) CANDMC_TRY(p2p_transport_recv(c, recv, rcount, src, st));
    return OK;
  }
  ncclComm_t comm = c->nccl;
  if (background) CANDMC_TRY(comm_background(c, &comm));
  CANDMC_NCCL(ncclGroupStart());
  if (do_send) CANDMC_NCCL(ncclSend(send, (size_t)scount, ncclDouble, dst, comm, st));
  if (do_recv) CANDMC_NCCL(ncclRecv(recv, (size_t)rcount, ncclDouble, src, comm, st));
  CANDMC_NCCL(ncclGroupEnd());
  return OK;
}

}  // namespace candmc

using namespace candmc;

extern "C" {

int candmc_get_unique_id(void* out_128_bytes) {
  CANDMC_CHECK(out_128_bytes != nullptr, "candmc_get_unique_id: null output");
  static_assert(sizeof(ncclUniqueId) <= CANDMC_UNIQUE_ID_BYTES, "ncclUniqueId larger than the ABI slot");
  ncclUniqueId id;
  CANDMC_NCCL(ncclGetUniqueId(&id));
  memset(out_128_bytes, 0, CANDMC_UNIQUE_ID_BYTES);
  memcpy(out_128_bytes, &id, sizeof(id));
  return OK;
}

int candmc_comm_init_rank(const void* unique_id_128_bytes, int nranks, int rank, candmc_comm_t** out) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(unique_id_128_bytes && out, "candmc_comm_init_rank: null argument");
  CANDMC_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "candmc_comm_init_rank: rank %d / %d", rank, nranks);
  ncclUniqueId id;
  memcpy(&id, unique_id_128_bytes, sizeof(id));
  candmc_comm* c = new candmc_comm();
  c->rank = rank;
  c->size = nranks;
  ncclResult_t r = ncclCommInitRank(&c->nccl, nranks, id, rank);
  if (r != ncclSuccess) {
    set_last_error("ncclCommInitRank(%d/%d) failed: %s", rank, nranks, ncclGetErrorString(r));
    delete c;
    return ERR_NCCL;
  }
  if (nranks > 1) {
    ncclComm_t bg;
    CANDMC_TRY(comm_background(c, &bg));  // eager: nothing is in flight on the new communicator yet
  }
  *out = c;
  return OK;
}

int candmc_comm_split(candmc_comm_t* parent, int color, int key, candmc_comm_t** out) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(parent && out, "candmc_comm_split: null argument");
  candmc_comm* c = new candmc_comm();
  if (parent->size == 1) {  // nothing to negotiate; keep a size-1 handle without an NCCL communicator
    c->nccl = nullptr;
    c->rank = 0;
    c->size = 1;
    *out = c;
    return OK;
  }
  ncclResult_t r = ncclCommSplit(parent->nccl, color, key, &c->nccl, nullptr);
  if (r != ncclSuccess) {
    set_last_error("ncclCommSplit(color=%d,key=%d) failed: %s", color, key, ncclGetErrorString(r));
    delete c;
    return ERR_NCCL;
  }
  CANDMC_NCCL(ncclCommUserRank(c->nccl, &c->rank));
  CANDMC_NCCL(ncclCommCount(c->nccl, &c->size));
  if (c->size > 1) {
    ncclComm_t bg;
    CANDMC_TRY(comm_background(c, &bg));  // eager: nothing is in flight on the new communicator yet
  }
  *out = c;
  return OK;
}

int candmc_comm_free(candmc_comm_t* comm) {
  if (!comm) return OK;
  // collectives enqueued on this communicator may still be waiting on their streams (every call with device operands is
  // asynchronous): let them run before the communicator goes away, as MPI_Comm_free lets pending operations complete
  cudaDeviceSynchronize();
  if (comm->fused_ctx) {
    candmc::FusedCtx* f = static_cast<candmc::FusedCtx*>(comm->fused_ctx);
    candmc::window_destroy(f->win);
    delete f;
  }
  if (comm->transport) {
    candmc::panel_transport_destroy(static_cast<candmc::PanelTransport*>(comm->transport));
  }
  if (comm->p2p) {
    candmc::p2p_transport_destroy(static_cast<candmc::P2PTransport*>(comm->p2p));
  }
  if (comm->nccl_bg && comm->nccl_bg != comm->nccl) ncclCommDestroy(comm->nccl_bg);
  if (comm->nccl) ncclCommDestroy(comm->nccl);
  delete comm;
  return OK;
}

int candmc_comm_rank(const candmc_comm_t* comm, int* rank) {
  CANDMC_CHECK(comm && rank, "candmc_comm_rank: null argument");
  *rank = comm->rank;
  return OK;
}

int candmc_comm_size(const candmc_comm_t* comm, int* size) {
  CANDMC_CHECK(comm && size, "candmc_comm_size: null argument");
  *size = comm->size;
  return OK;
}

int candmc_comm_barrier(candmc_comm_t* comm) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(comm != nullptr, "candmc_comm_barrier: null communicator");
  CANDMC_CUDA(cudaDeviceSynchronize());
  if (comm->size > 1) {
    double* d = nullptr;
    CANDMC_CUDA(cudaMalloc(&d, sizeof(double)));
    CANDMC_CUDA(cudaMemsetAsync(d, 0, sizeof(double), runtime().comm_stream));
    int rc = comm_allreduce(comm, d, d, 1, runtime().comm_stream);
    cudaError_t e = cudaStreamSynchronize(runtime().comm_stream);
    cudaFree(d);
    if (rc != OK) return rc;
    CANDMC_CUDA(e);
  }
  return OK;
}

int candmc_comm_bcast(candmc_comm_t* comm, double* buf, int64_t count, int root, void* stream) {
  CANDMC_TRY(runtime_require());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return OK;
  if (is_device_ptr(buf)) return comm_bcast(comm, buf, buf, count, root, st);
  StagedMatrix s;
  CANDMC_CHECK(comm != nullptr, "candmc_comm_bcast: null communicator");
  CANDMC_TRY(s.open(buf, count, 1, count, comm->rank == root, st));
  CANDMC_TRY(comm_bcast(comm, s.ptr(), s.ptr(), count, root, st));
  if (comm->rank != root) CANDMC_TRY(s.close_out(st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int candmc_comm_allreduce_sum(candmc_comm_t* comm, const double* sendbuf, double* recvbuf, int64_t count,
                              void* stream) {
  CANDMC_TRY(runtime_require());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return OK;
  if (is_device_ptr(sendbuf) && is_device_ptr(recvbuf)) return comm_allreduce(comm, sendbuf, recvbuf, count, st);
  StagedMatrix s, r;
  CANDMC_TRY(s.open(sendbuf, count, 1, count, true, st));
  CANDMC_TRY(r.open(recvbuf, count, 1, count, false, st));
  CANDMC_TRY(comm_allreduce(comm, s.ptr(), r.ptr(), count, st));
  CANDMC_TRY(r.close_out(st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

}  // extern "C"
