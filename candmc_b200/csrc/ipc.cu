// candmc_b200 — CUDA-IPC peer windows and the host side of the fused GEMM + depth all-reduce.
#include "ipc.h"

#include <string.h>

#include <vector>

#include "runtime.h"

namespace candmc {

int window_create(candmc_comm* c, size_t bytes, PeerWindow** out) {
  *out = nullptr;
  CANDMC_CHECK(c != nullptr && c->size >= 1 && c->size <= kMaxPeers, "peer window: communicator size %d unsupported",
               c ? c->size : 0);
  PeerWindow* w = new PeerWindow();
  w->size = c->size;
  w->rank = c->rank;
  w->bytes = bytes;
  void* local = nullptr;
  cudaError_t e = cudaMalloc(&local, bytes);
  int ok = (e == cudaSuccess);
  if (ok) e = cudaMemset(local, 0, bytes);
  ok = ok && (e == cudaSuccess);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = (cudaIpcGetMemHandle(&mine, local) == cudaSuccess);
  if (!ok) cudaGetLastError();   // a failed export must not surface as the "last error" of the next kernel launch
  // all-gather {ok flag, handle} through NCCL (device staging), then open the peers' handles
  struct Rec {
    int ok;
    cudaIpcMemHandle_t h;
  };
  Rec rec;
  rec.ok = ok;
  rec.h = mine;
  std::vector<Rec> all(c->size);
  // Both exchanges below are collective: a rank whose local CUDA call fails must still take part in them (with ok = 0), or
  // its peers would wait in the all-gather for good.  Only an NCCL error — the communicator itself is broken — returns early.
  cudaStream_t st = runtime().comm_stream;
  auto gather = [&](const void* mine_rec, void* all_recs, size_t rec_bytes, bool* local_ok) -> int {
    char *ds = nullptr, *dr = nullptr;
    bool good = cudaMalloc(&ds, rec_bytes) == cudaSuccess && cudaMalloc(&dr, rec_bytes * c->size) == cudaSuccess;
    good = good && cudaMemcpyAsync(ds, mine_rec, rec_bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
    int rc = OK;
    if (good) {   // (without device scratch there is nothing to take part with; that rank's peers see an NCCL timeout)
      ncclResult_t nr = ncclAllGather(ds, dr, rec_bytes, ncclChar, c->nccl, st);
      if (nr != ncclSuccess) { set_last_error("peer window: ncclAllGather failed"); rc = ERR_NCCL; }
      good = rc == OK && cudaMemcpyAsync(all_recs, dr, rec_bytes * c->size, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
    }
    if (!good) cudaGetLastError();
    cudaFree(ds);
    cudaFree(dr);
    *local_ok = good;
    return rc;
  };
  bool all_ok = true;
  if (c->size == 1) {
    all[0] = rec;
  } else {
    bool got = false;
    int rc = gather(&rec, all.data(), sizeof(Rec), &got);
    if (rc != OK) { if (local) cudaFree(local); delete w; return rc; }
    if (!got) { all_ok = false; for (int r = 0; r < c->size; ++r) all[r].ok = 0; }
  }
  for (int r = 0; r < c->size; ++r) all_ok = all_ok && all[r].ok;
  if (all_ok) {
    for (int r = 0; r < c->size && all_ok; ++r) {
      if (r == c->rank) {
        w->base[r] = static_cast<char*>(local);
      } else {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          all_ok = false;
        }
        w->base[r] = static_cast<char*>(p);
      }
    }
  }
  // agree on the outcome (one more tiny all-gather) so that either every rank uses the window or none does
  if (c->size > 1) {
    int mine_ok = all_ok ? 1 : 0;
    std::vector<int> oks(c->size, 0);
    bool got = false;
    int rc = gather(&mine_ok, oks.data(), sizeof(int), &got);
    if (rc != OK) { w->base[c->rank] = static_cast<char*>(local); window_destroy(w); return rc; }
    all_ok = all_ok && got;
    for (int r = 0; r < c->size; ++r) all_ok = all_ok && oks[r];
  }
  if (!all_ok) {
    w->base[c->rank] = static_cast<char*>(local);
    window_destroy(w);
    set_last_error("peer window: CUDA IPC mapping unavailable on at least one rank");
    return ERR_CUDA;
  }
  *out = w;
  return OK;
}

void window_destroy(PeerWindow* w) {
  if (!w) return;
  for (int r = 0; r < w->size; ++r) {
    if (!w->base[r]) continue;
    if (r == w->rank) cudaFree(w->base[r]);
    else cudaIpcCloseMemHandle(w->base[r]);
  }
  delete w;
}

namespace {
size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// wait until *counter (written by peers with system-scope atomics) reaches `target` (wrap-safe), one warp
__global__ void wait_counter_kernel(const uint32_t* counter, uint32_t target) {
  if (threadIdx.x == 0) {
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (static_cast<int32_t>(v - target) < 0);
  }
}
}  // namespace

int fused_ctx_get(candmc_comm* kdir, int64_t b, FusedCtx** out) {
  *out = nullptr;
  if (!runtime().fused_reduce || kdir == nullptr || kdir->size < 2 || kdir->size > kMaxPeers) return OK;
  const int c = kdir->size;
  if (b % (128 * c) != 0) return OK;  // whole tile columns per owner, no ragged tiles
  if (kdir->fused_failed) return OK;
  FusedCtx* ctx = static_cast<FusedCtx*>(kdir->fused_ctx);
  if (ctx && ctx->b == b) {
    *out = ctx;
    return OK;
  }
  if (ctx) {  // block size changed: all ranks get here in the same call -> re-create collectively
    CANDMC_CUDA(cudaDeviceSynchronize());
    window_destroy(ctx->win);
    delete ctx;
    kdir->fused_ctx = nullptr;
  }
  ctx = new FusedCtx();
  ctx->b = b;
  ctx->c = c;
  const int64_t tiles_owner = (b / 128) * (b / 128 / c);
  ctx->slab_elems = static_cast<size_t>(b) * static_cast<size_t>(b / c);
  ctx->off_flags = 0;
  ctx->off_done = align_up(sizeof(uint32_t) * (c - 1) * tiles_owner, 256);
  ctx->off_stage = ctx->off_done + 256;
  ctx->off_final = ctx->off_stage + align_up(sizeof(double) * (c - 1) * ctx->slab_elems, 256);
  const size_t total = ctx->off_final + sizeof(double) * 2 * (c - 1) * ctx->slab_elems;  // final: double-buffered
  PeerWindow* w = nullptr;
  int rc = window_create(kdir, total, &w);
  if (rc != OK) {
    delete ctx;
    kdir->fused_failed = true;  // sticky: the NCCL all-reduce is used from now on
    return OK;
  }
  ctx->win = w;
  kdir->fused_ctx = ctx;
  *out = ctx;
  return OK;
}

void fused_params_next(FusedCtx* ctx, int me, FusedParams* p) {
  const int c = ctx->c;
  const int64_t tiles_owner = (ctx->b / 128) * (ctx->b / 128 / c);
  ctx->epoch += 1;
  ctx->done_expected += static_cast<uint32_t>((c - 1) * tiles_owner);
  p->c = c;
  p->me = me;
  p->tiles_n_per_owner = static_cast<int>(ctx->b / 128 / c);
  p->epoch = ctx->epoch;
  p->ld = ctx->b;
  p->parity = static_cast<int>(ctx->epoch & 1);
  for (int r = 0; r < kMaxPeers; ++r) {
    p->stage[r] = nullptr;
    p->sflag[r] = nullptr;
    p->cfinal[r] = nullptr;
    p->done[r] = nullptr;
  }
  for (int o = 0; o < c; ++o) {
    if (o == me) continue;
    char* ob = ctx->win->base[o];
    const int slot_me_at_o = (me - o - 1 + c) % c;   // slot of source `me` in owner o's stage / flag arrays
    p->stage[o] = reinterpret_cast<double*>(ob + ctx->off_stage) + static_cast<size_t>(slot_me_at_o) * ctx->slab_elems;
    p->sflag[o] = reinterpret_cast<uint32_t*>(ob + ctx->off_flags) + static_cast<size_t>(slot_me_at_o) * tiles_owner;
    // my owned slab lands in peer o's final region, slot of owner `me` at o, parity half
    p->cfinal[o] = reinterpret_cast<double*>(ob + ctx->off_final) +
                   (static_cast<size_t>(p->parity) * (c - 1) + slot_me_at_o) * ctx->slab_elems;
    p->done[o] = reinterpret_cast<uint32_t*>(ob + ctx->off_done);
  }
  char* lb = ctx->win->base[me];
  p->stage_local = reinterpret_cast<double*>(lb + ctx->off_stage);
  p->sflag_local = reinterpret_cast<uint32_t*>(lb + ctx->off_flags);
  p->Cin = nullptr;
  p->ldin = 0;
}

int fused_finish(FusedCtx* ctx, int me, const FusedParams& p, double* C, int64_t ldc, cudaStream_t st) {
  const int c = ctx->c;
  char* lb = ctx->win->base[me];
  wait_counter_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(lb + ctx->off_done), ctx->done_expected);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  const int64_t b = ctx->b, cols = b / c;
  for (int o = 0; o < c; ++o) {
    if (o == me) continue;
    const int slot_o_at_me = (o - me - 1 + c) % c;
    const double* src = reinterpret_cast<const double*>(lb + ctx->off_final) +
                        (static_cast<size_t>(p.parity) * (c - 1) + slot_o_at_me) * ctx->slab_elems;
    CANDMC_TRY(lda_copy_f64(b, cols, b, ldc, src, C + static_cast<int64_t>(o) * cols * ldc, st));
  }
  return OK;
}

}  // namespace candmc
