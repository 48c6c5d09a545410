// candmc_b200 — symmetric peer windows over CUDA IPC: every rank of a communicator allocates the same number of
// bytes and maps everyone else's allocation, so kernels (P2P loads/stores over NVLink/NVSwitch) and copy engines
// can address peer memory directly.  One process per GPU, so the mappings come from cudaIpcGetMemHandle /
// cudaIpcOpenMemHandle; the 64-byte handles are exchanged with an NCCL all-gather at window creation only.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "comm.h"

namespace candmc {

constexpr int kMaxPeers = 8;

struct PeerWindow {
  int size = 0, rank = 0;
  size_t bytes = 0;
  char* base[kMaxPeers] = {nullptr};  // base[r] = rank r's allocation in MY address space (base[rank] is local)
};

// Collective over `c` (size <= kMaxPeers).  On failure every rank gets an error (the caller falls back to NCCL).
int window_create(candmc_comm* c, size_t bytes, PeerWindow** out);
void window_destroy(PeerWindow* w);

// ---- fused GEMM + depth all-reduce (see gemm_f64.cu, FusedParams) ------------------------------------------------------
struct FusedParams {
  int c = 1, me = 0;         // depth group size and my rank in it
  int tiles_n_per_owner = 0; // tile columns owned by each rank (tilesN / c)
  uint32_t epoch = 0;        // value a flag must reach in THIS call
  int64_t ld = 0;            // leading dimension of the stage / final regions (= b)
  int parity = 0;            // which half of the double-buffered final region this call uses
  // peer-mapped pointers, indexed by depth rank
  double* stage[kMaxPeers];     // stage[o]: region in owner o's window that receives MY partial tiles (o != me)
  uint32_t* sflag[kMaxPeers];   // sflag[o]: per-tile flags in owner o's window for source me (o != me)
  double* cfinal[kMaxPeers];    // cfinal[p]: slab in peer p's window that receives MY owned (final) tiles
  uint32_t* done[kMaxPeers];    // done[p]: counter in peer p's window, +1 per final tile delivered
  // local views
  double* stage_local;          // (c-1) slots of ld x (ld/c)
  uint32_t* sflag_local;        // (c-1) x tiles_per_owner
  const double* Cin;            // beta source (partial sums of earlier k-chunks) — may be null when beta == 0
  int64_t ldin;
};

struct FusedCtx {
  PeerWindow* win = nullptr;
  int64_t b = 0;
  int c = 0;
  uint32_t epoch = 0;
  uint32_t done_expected = 0;
  size_t off_flags = 0, off_done = 0, off_stage = 0, off_final = 0;
  size_t slab_elems = 0;  // ld * (ld / c)
};

// Returns (creating or re-creating collectively when b changes) the fused-reduce context of depth communicator `kdir`,
// or nullptr in *out if the fused path cannot be used (the caller then uses ncclAllReduce).
int fused_ctx_get(candmc_comm* kdir, int64_t b, FusedCtx** out);
// Fill the kernel parameters for the next fused call (advances the epoch).
void fused_params_next(FusedCtx* ctx, int me, FusedParams* p);
// The epoch and the delivery count are advanced before the steps that can still fail (operand checks, tensor maps, the launch).
// A call that returns an error before its fused launch was enqueued takes both back, so that the next call on the communicator
// does not wait for deliveries that were never made.  (Argument errors are the same on every depth rank; a peer that dies
// mid-call still leaves the others spinning — there is no timeout in the kernels.)
struct FusedEpochGuard {
  FusedCtx* ctx = nullptr;
  uint32_t epoch0 = 0, done0 = 0;
  unsigned long long launches0 = 0;
  void arm(FusedCtx* c, unsigned long long fused_launches_now) {   // call BEFORE fused_params_next
    ctx = c; epoch0 = c->epoch; done0 = c->done_expected; launches0 = fused_launches_now;
  }
  void settle(unsigned long long fused_launches_now) {
    if (ctx && fused_launches_now == launches0) { ctx->epoch = epoch0; ctx->done_expected = done0; }
    ctx = nullptr;
  }
};
// After the fused GEMM: wait (on `st`) until every peer delivered its owned tiles, then copy them into C (ld ldc).
int fused_finish(FusedCtx* ctx, int me, const FusedParams& p, double* C, int64_t ldc, cudaStream_t st);

}  // namespace candmc
