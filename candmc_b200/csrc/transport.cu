// candmc_b200 — SUMMA panel transport over peer memory with copy engines only (see transport.h).
#include "transport.h"

#include <algorithm>

#include <cuda.h>

#include "../../include/candmc_b200.h"
#include "runtime.h"

namespace candmc {

namespace {

typedef CUresult (*PFN_waitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
PFN_waitValue32 g_wait32 = nullptr;
bool g_wait32_tried = false;
unsigned g_wait_flags = CU_STREAM_WAIT_VALUE_GEQ;   // | FLUSH where the device can flush outstanding remote writes behind a wait
uint32_t* g_vals = nullptr;   // device array vals[i] = i: the 4-byte sources of the flag DMAs

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__global__ void iota_u32_kernel(uint32_t* v, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

int ensure_globals() {
  if (!g_wait32_tried) {
    g_wait32_tried = true;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && fn != nullptr &&
        q == cudaDriverEntryPointSuccess)
      g_wait32 = reinterpret_cast<PFN_waitValue32>(fn);
    else
      cudaGetLastError();
    // the flags are written by a peer's copy engine: ask for the remote-write flush behind the wait where it exists, so that
    // the chunk the flag announces is visible to the kernel that follows
    int can_flush = 0;
    if (cudaDeviceGetAttribute(&can_flush, cudaDevAttrCanFlushRemoteWrites, runtime().device) == cudaSuccess && can_flush)
      g_wait_flags |= CU_STREAM_WAIT_VALUE_FLUSH;
    else
      cudaGetLastError();
  }
  if (g_wait32 == nullptr) return ERR_CUDA;
  if (g_vals == nullptr) {
    CANDMC_CUDA(cudaMalloc(&g_vals, sizeof(uint32_t) * kPanelMaxCalls));
    iota_u32_kernel<<<runtime().num_sms * 4, 256, 0, runtime().aux_stream>>>(g_vals, kPanelMaxCalls);
    CANDMC_CUDA(cudaGetLastError());
    CANDMC_CUDA(cudaStreamSynchronize(runtime().aux_stream));
    runtime().launches++;
  }
  return OK;
}

int stream_wait_geq(cudaStream_t st, const uint32_t* addr, uint32_t value) {
  CUresult r = g_wait32(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(addr), value, g_wait_flags);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuStreamWaitValue32 failed (CUresult %d)", (int)r);
    return ERR_CUDA;
  }
  return OK;
}

}  // namespace

int panel_transport_get(candmc_comm* c, int64_t half_elems, PanelTransport** out) {
  *out = nullptr;
  if (c == nullptr || c->size < 2 || c->size > kMaxPeers || c->transport_failed) return OK;
  PanelTransport* t = static_cast<PanelTransport*>(c->transport);
  // (a transport whose call counter is about to run out of the flag-value table is rebuilt like one that has to grow: every
  // rank counts the same calls, so every rank gets here in the same call)
  if (t != nullptr && t->half_elems >= half_elems && t->call + 8 < kPanelMaxCalls) {
    *out = t;
    return OK;
  }
  if (t != nullptr && t->half_elems > half_elems) half_elems = t->half_elems;
  if (ensure_globals() != OK) {   // same outcome on every rank of a node (one driver)
    c->transport_failed = true;
    return OK;
  }
  if (t != nullptr) {   // grow: every rank gets here in the same call
    CANDMC_CUDA(cudaDeviceSynchronize());
    CANDMC_TRY(candmc_comm_barrier(c));   // nobody still copies into the old windows
    panel_transport_destroy(t);
    c->transport = nullptr;
  }
  t = new PanelTransport();
  t->half_elems = half_elems;
  t->off_ready = 0;
  t->off_done = align_up(sizeof(uint32_t) * 2 * kPanelMaxOps, 256);
  t->off_data = t->off_done + 256;
  const size_t total = t->off_data + sizeof(double) * 2 * static_cast<size_t>(half_elems);
  PeerWindow* w = nullptr;
  if (window_create(c, total, &w) != OK) {   // zero-filled: every flag starts at 0 = "never"
    delete t;
    c->transport_failed = true;   // sticky: NCCL from now on
    return OK;
  }
  t->win = w;
  c->transport = t;
  *out = t;
  return OK;
}

void panel_transport_begin(PanelTransport* t) {
  t->call += 1;
  for (int p = 0; p < kMaxPeers; ++p) t->waited_done[p] = false;
}

int panel_transport_send(PanelTransport* t, candmc_comm* c, int op, int64_t slot_off, const double* src, int64_t count,
                         cudaStream_t copy) {
  CANDMC_CHECK(op >= 0 && op < kPanelMaxOps && slot_off >= 0 && slot_off + count <= t->half_elems, "panel transport: slot out of range");
  CANDMC_CHECK(t->call < kPanelMaxCalls, "panel transport: call counter exhausted");
  const int half = static_cast<int>(t->call & 1u);
  char* mine = t->win->base[c->rank];
  for (int p = 0; p < c->size; ++p) {
    if (p == c->rank) continue;
    if (!t->waited_done[p]) {
      // peer p has finished the call that used this half before (call - 2); flags start at 0 and calls at 2, so the very
      // first two calls pass immediately
      if (t->call >= 4)
        CANDMC_TRY(stream_wait_geq(copy, reinterpret_cast<const uint32_t*>(mine + t->off_done) + p, t->call - 2));
      t->waited_done[p] = true;
    }
    char* pb = t->win->base[p];
    double* dst = reinterpret_cast<double*>(pb + t->off_data) + static_cast<int64_t>(half) * t->half_elems + slot_off;
    CANDMC_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyDeviceToDevice, copy));
    uint32_t* flag = reinterpret_cast<uint32_t*>(pb + t->off_ready) + half * kPanelMaxOps + op;
    CANDMC_CUDA(cudaMemcpyAsync(flag, g_vals + t->call, sizeof(uint32_t), cudaMemcpyDeviceToDevice, copy));
    runtime().transport_sends++;
  }
  return OK;
}

int panel_transport_wait(PanelTransport* t, candmc_comm* c, int op, int64_t slot_off, cudaStream_t compute, const double** data) {
  CANDMC_CHECK(op >= 0 && op < kPanelMaxOps, "panel transport: slot out of range");
  const int half = static_cast<int>(t->call & 1u);
  char* mine = t->win->base[c->rank];
  // a ready flag of this half holds the number of the last call that filled the slot: call, or call - 2, - 4, ... before that
  CANDMC_TRY(stream_wait_geq(compute, reinterpret_cast<const uint32_t*>(mine + t->off_ready) + half * kPanelMaxOps + op, t->call));
  *data = reinterpret_cast<const double*>(mine + t->off_data) + static_cast<int64_t>(half) * t->half_elems + slot_off;
  return OK;
}

int panel_transport_end(PanelTransport* t, candmc_comm* c, cudaStream_t compute, cudaStream_t copy) {
  cudaEvent_t e;
  CANDMC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CANDMC_CUDA(cudaEventRecord(e, compute));
  CANDMC_CUDA(cudaStreamWaitEvent(copy, e, 0));
  CANDMC_CUDA(cudaEventDestroy(e));   // released when the wait has consumed it
  for (int p = 0; p < c->size; ++p) {
    if (p == c->rank) continue;
    uint32_t* flag = reinterpret_cast<uint32_t*>(t->win->base[p] + t->off_done) + c->rank;
    CANDMC_CUDA(cudaMemcpyAsync(flag, g_vals + t->call, sizeof(uint32_t), cudaMemcpyDeviceToDevice, copy));
  }
  return OK;
}

void panel_transport_destroy(PanelTransport* t) {
  if (!t) return;
  window_destroy(t->win);
  delete t;
}

// ---- point-to-point --------------------------------------------------------------------------------------------------------
int p2p_transport_prepare(candmc_comm* c, int64_t slot_elems) {
  if (c == nullptr || !runtime().panel_transport || c->size < 2 || c->size > kMaxPeers || c->transport_failed) return OK;
  P2PTransport* t = static_cast<P2PTransport*>(c->p2p);
  uint32_t max_seq = 0;   // the message counters of the busiest pair; symmetric exchanges keep them equal on every rank
  if (t != nullptr)
    for (int p = 0; p < kMaxPeers; ++p) max_seq = std::max(max_seq, std::max(t->send_seq[p], t->recv_seq[p]));
  if (t != nullptr && t->slot_elems >= slot_elems && max_seq + 4096 < kPanelMaxCalls) return OK;
  if (t != nullptr && t->slot_elems > slot_elems) slot_elems = t->slot_elems;
  if (ensure_globals() != OK) {
    c->transport_failed = true;
    return OK;
  }
  if (t != nullptr) {   // grow: every rank gets here in the same call; the message counters start over with the new window
    CANDMC_CUDA(cudaDeviceSynchronize());
    CANDMC_TRY(candmc_comm_barrier(c));
    p2p_transport_destroy(t);
    c->p2p = nullptr;
  }
  t = new P2PTransport();
  t->slot_elems = slot_elems;
  t->off_ready = 0;
  t->off_ack = align_up(sizeof(uint32_t) * kMaxPeers * kP2PSlots, 256);
  t->off_data = t->off_ack + 256;
  const size_t total = t->off_data + sizeof(double) * static_cast<size_t>(c->size) * kP2PSlots * static_cast<size_t>(slot_elems);
  PeerWindow* w = nullptr;
  if (window_create(c, total, &w) != OK) {
    delete t;
    c->transport_failed = true;
    return OK;
  }
  t->win = w;
  c->p2p = t;
  return OK;
}

bool p2p_transport_usable(const candmc_comm* c, int64_t count) {
  const P2PTransport* t = c ? static_cast<const P2PTransport*>(c->p2p) : nullptr;
  return t != nullptr && count <= t->slot_elems;
}

int p2p_transport_send(candmc_comm* c, const double* send, int64_t count, int dst, cudaStream_t st) {
  P2PTransport* t = static_cast<P2PTransport*>(c->p2p);
  CANDMC_CHECK(t != nullptr && dst >= 0 && dst < c->size && dst != c->rank && count <= t->slot_elems, "p2p transport: bad send");
  const uint32_t k = ++t->send_seq[dst];
  CANDMC_CHECK(k < kPanelMaxCalls, "p2p transport: message counter exhausted");
  const int slot = static_cast<int>(k % kP2PSlots);
  char* mine = t->win->base[c->rank];
  if (k > static_cast<uint32_t>(kP2PSlots))   // the message that used this slot before has been copied out by dst
    CANDMC_TRY(stream_wait_geq(st, reinterpret_cast<const uint32_t*>(mine + t->off_ack) + dst, k - kP2PSlots));
  char* pb = t->win->base[dst];
  double* slotp = reinterpret_cast<double*>(pb + t->off_data) + (static_cast<int64_t>(c->rank) * kP2PSlots + slot) * t->slot_elems;
  CANDMC_CUDA(cudaMemcpyAsync(slotp, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, st));
  uint32_t* flag = reinterpret_cast<uint32_t*>(pb + t->off_ready) + c->rank * kP2PSlots + slot;
  CANDMC_CUDA(cudaMemcpyAsync(flag, g_vals + k, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  runtime().transport_sends++;
  return OK;
}

int p2p_transport_recv(candmc_comm* c, double* recv, int64_t count, int src, cudaStream_t st) {
  P2PTransport* t = static_cast<P2PTransport*>(c->p2p);
  CANDMC_CHECK(t != nullptr && src >= 0 && src < c->size && src != c->rank && count <= t->slot_elems, "p2p transport: bad receive");
  const uint32_t k = ++t->recv_seq[src];
  const int slot = static_cast<int>(k % kP2PSlots);
  char* mine = t->win->base[c->rank];
  CANDMC_TRY(stream_wait_geq(st, reinterpret_cast<const uint32_t*>(mine + t->off_ready) + src * kP2PSlots + slot, k));
  const double* slotp = reinterpret_cast<const double*>(mine + t->off_data) + (static_cast<int64_t>(src) * kP2PSlots + slot) * t->slot_elems;
  CANDMC_CUDA(cudaMemcpyAsync(recv, slotp, sizeof(double) * count, cudaMemcpyDeviceToDevice, st));
  uint32_t* ack = reinterpret_cast<uint32_t*>(t->win->base[src] + t->off_ack) + c->rank;
  CANDMC_CUDA(cudaMemcpyAsync(ack, g_vals + k, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  return OK;
}

void p2p_transport_destroy(P2PTransport* t) {
  if (!t) return;
  window_destroy(t->win);
  delete t;
}

}  // namespace candmc
