// candmc_b200 — SUMMA panel transport over peer memory with copy engines only (no SM, no NCCL kernel).
//
// NCCL moves the panel chunks with SM-resident kernels, so while panels are in flight the persistent GEMM gives up
// `bg_max_ctas` SMs (1.35 % of the DMMA rate on 148 SMs) and the two kinds of kernels still interfere at launch boundaries.
// With this transport the root of a panel chunk writes it straight into the consumers' memory (CUDA-IPC-mapped windows,
// ipc.h) with cudaMemcpyAsync on the copy stream — a DMA over NVLink — followed on the same stream by a 4-byte DMA that
// raises the chunk's ready flag in the consumer's window; the consumer's compute stream waits for the flag with
// cuStreamWaitValue32 (a stream memory operation: no kernel either) and multiplies straight out of its window.
//
// Reuse across calls: a window holds two halves, call k of a communicator uses half k & 1.  When a rank has issued the
// last multiply of call k it DMA-writes k into done[me] of every peer; a root starts writing half (k & 1) of peer p only
// after done[p] >= k - 2 (again a stream wait on its own window).  Inside one call every (panel, chunk) has its own slot.
//
// Opt-in (candmc_set_panel_transport) until it has been measured on B200s; validated on the CPU simulator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "comm.h"
#include "ipc.h"

namespace candmc {

constexpr int kPanelMaxOps = 256;          // (panels x chunks) of one sweep on one communicator
constexpr uint32_t kPanelMaxCalls = 1u << 22;

struct PanelTransport {
  PeerWindow* win = nullptr;
  int64_t half_elems = 0;      // capacity of one half in doubles
  uint32_t call = 1;           // sequence number of the sweep in progress (0 = "never")
  size_t off_ready = 0, off_done = 0, off_data = 0;
  bool waited_done[kMaxPeers] = {false};   // root side: done[p] >= call - 2 already enqueued for this call
};

// Returns (creating or growing collectively over `c`) the transport of communicator `c` with room for `half_elems` doubles
// per half, or nullptr in *out when peer windows / stream memory operations are not available (callers then use NCCL).
int panel_transport_get(candmc_comm* c, int64_t half_elems, PanelTransport** out);
// New sweep on this communicator (same call order on every rank).
void panel_transport_begin(PanelTransport* t);
// Root: chunk `op` (slot offset `slot_off` doubles inside the half, `count` doubles from `src`) to every other rank of `c`.
int panel_transport_send(PanelTransport* t, candmc_comm* c, int op, int64_t slot_off, const double* src, int64_t count,
                         cudaStream_t copy);
// Consumer: make `compute` wait until chunk `op` of this call has landed; *data = where it is.
int panel_transport_wait(PanelTransport* t, candmc_comm* c, int op, int64_t slot_off, cudaStream_t compute, const double** data);
// After the last multiply of the call has been enqueued on `compute`: tell every peer that this rank is done with the half.
int panel_transport_end(PanelTransport* t, candmc_comm* c, cudaStream_t compute, cudaStream_t copy);
void panel_transport_destroy(PanelTransport* t);

// ---- point-to-point exchanges (Cannon staggers and shifts) the same way -------------------------------------------------
// Every rank's window has, for every possible source rank, a ring of kP2PSlots message slots with a ready flag each, and one
// acknowledgement flag per destination.  Message number k from s to d (k = 1, 2, ... per ordered pair, counted on both sides):
//   sender:   wait ack[d] >= k - kP2PSlots (own window)  ->  DMA the data into d's slot k % kP2PSlots of lane s
//             ->  DMA k into d's ready flag of that slot;
//   receiver: wait ready[s][k % kP2PSlots] >= k  ->  DMA the slot into the destination buffer  ->  DMA k into s's ack[d].
// Nothing but copy engines and stream memory operations, so a shift runs under the GEMM of the current step without taking
// SMs from it (the NCCL version could only run at GEMM boundaries, DESIGN.md §7 "known issue").
constexpr int kP2PSlots = 2;
struct P2PTransport {
  PeerWindow* win = nullptr;
  int64_t slot_elems = 0;
  uint32_t send_seq[kMaxPeers] = {0};
  uint32_t recv_seq[kMaxPeers] = {0};
  size_t off_ready = 0, off_ack = 0, off_data = 0;
};
// Collective over `c`: (re)creates the transport with room for messages of `slot_elems` doubles; c->p2p stays null when the
// transport is off or unavailable.
int p2p_transport_prepare(candmc_comm* c, int64_t slot_elems);
bool p2p_transport_usable(const candmc_comm* c, int64_t count);
int p2p_transport_send(candmc_comm* c, const double* send, int64_t count, int dst, cudaStream_t st);
int p2p_transport_recv(candmc_comm* c, double* recv, int64_t count, int src, cudaStream_t st);
void p2p_transport_destroy(P2PTransport* t);

}  // namespace candmc
