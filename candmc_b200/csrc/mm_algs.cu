// candmc_b200 — host schedules of the distributed multiplies (the CANMM algorithms) on CUDA streams + NCCL.
//
//   candmc_summa            <- summa            (alg/MM/topo_pdgemm/summa.cxx:26-101)
//   candmc_d25_summa        <- d25_summa[_ovp]  (alg/MM/topo_pdgemm/d25_summa.cxx:33-281)
//   candmc_bcast_cannon_4d  <- bcast_cannon_4d  (alg/MM/topo_pdgemm/dual_cannon.cxx:40-215, intended semantics)
//   candmc_spcannon         <- kput_cannon / kuni_cannon (alg/MM/splitdim_cannon/spcannon.cxx:33-347)
//   candmc_upd_A            <- the GEMM pair + allreduce + trsm of upd_A (alg/QR/qr_2d/qr_2d.cxx:259-275)
//
// What is kept from the reference: which block every rank owns, which rank is the root of every panel, the order of
// the panels, what is summed into C and where the result lives.  What is new: every panel is cut into k-chunks whose
// NCCL broadcast (comm stream) runs under the DMMA GEMM of the previous chunk (compute stream = the caller's stream);
// the root multiplies straight out of the caller's matrices (no self-copy); the B panel is re-laid out chunk-major
// by the pack kernel while it is being copied out of its lda anyway, so every message is contiguous; Cannon shifts
// land in the alternate buffer while the current step multiplies (pointer swap instead of the reference's memcpy
// back, spcannon.cxx:76-77); the transposes of split-dim Cannon are folded into the NT GEMM where possible.
#include <algorithm>
#include <functional>
#include <vector>

#include "../../include/candmc_b200.h"
#include "comm.h"
#include "common.cuh"
#include "ipc.h"
#include "runtime.h"
#include "staging.h"
#include "transport.h"

namespace candmc {

namespace {

// ---- small utilities ------------------------------------------------------------------------------------------
class EventPool {
 public:
  cudaEvent_t get() {
    if (next_ == ev_.size()) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      ev_.push_back(e);
    }
    return ev_[next_++];
  }
  void reset() { next_ = 0; }

 private:
  std::vector<cudaEvent_t> ev_;
  size_t next_ = 0;
};
EventPool g_events;

// `to` waits for everything enqueued so far on `from`
int stream_wait(cudaStream_t to, cudaStream_t from) {
  if (to == from) return OK;
  cudaEvent_t e = g_events.get();
  CANDMC_CHECK(e != nullptr, "event pool: cudaEventCreate failed");
  CANDMC_CUDA(cudaEventRecord(e, from));
  CANDMC_CUDA(cudaStreamWaitEvent(to, e, 0));
  return OK;
}

bool is_t(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }
bool is_n(char t) { return t == 'N' || t == 'n'; }

// While background NCCL traffic is in flight the GEMMs leave that many SMs free (see summa_sweep).
struct ReserveGuard {
  int saved;
  explicit ReserveGuard(int r) : saved(runtime().gemm_reserve_sms) { runtime().gemm_reserve_sms = r; }
  ~ReserveGuard() { runtime().gemm_reserve_sms = saved; }
};

// number of k-chunks a b-wide panel is cut into: as many as 8, each at least `min_kchunk` wide and even
int pick_chunks(int64_t b) {
  const int64_t min_kc = runtime().min_kchunk;
  for (int nc = 8; nc > 1; nc >>= 1)
    if (b % nc == 0 && (b / nc) >= min_kc && (b / nc) % 2 == 0) return nc;
  return 1;
}

// ---- one SUMMA sweep: C (+)= sum_{i in [i0,i1)} A_i * B_i over a q x q grid ---------------------------------------
struct SummaArgs {
  char tA, tB;
  int64_t b;
  int i0, i1;
  const double* myA;  // my block of A (device), used when I am the root of an A panel
  int64_t ldA;
  const double* myB;
  int64_t ldB;
  double* C;
  int64_t ldC;
  bool first_beta_zero;  // beta = 0 on the very first multiply (reference: (i>0)*1.0, summa.cxx:97)
  candmc_comm* row;      // along my grid row: rank = my column (cdt_row)
  candmc_comm* col;      // along my grid column: rank = my row (cdt_col)
  double* ws;            // >= 4*b*b doubles: packA | locB | bufA | bufB
  cudaStream_t compute;
  // operands that are still being uploaded from host memory (see upload_chunks): chunk t of myA / myB is valid once
  // a_ready[t] / b_ready[t] has fired; b_chunk_major: myB is already laid out chunk-major (chunk t = kc x b, ld = kc)
  const std::vector<cudaEvent_t>* a_ready = nullptr;
  const std::vector<cudaEvent_t>* b_ready = nullptr;
  bool b_chunk_major = false;
  // when set, the LAST multiply of the sweep also performs the depth all-reduce in its epilogue (ipc.h): it reads the
  // partial sums of the earlier multiplies from C (beta) and writes the reduced block to fused_out
  FusedParams* fused = nullptr;
  double* fused_out = nullptr;
  int64_t fused_ldout = 0;
  // Early finalisation of C for callers that still have to ship it somewhere slow (host memory): the second half of the
  // LAST panel's k-chunks is multiplied column slab by column slab (fin_slabs slabs), and slab_done(c0, w) is called once
  // columns [c0, c0 + w) of C hold their final value of this sweep, while the remaining slabs are still being multiplied.
  int fin_slabs = 0;
  std::function<int(int64_t, int64_t)> slab_done;
};

// number of column slabs a b-wide C block is finalised in (0: not worth it)
int pick_fin_slabs(int64_t b) {
  for (int s = 8; s > 1; s >>= 1)
    if (b % (s * 128) == 0) return s;
  return 0;
}

// The fused epilogue only exists in the TMA kernel: give it 16-byte aligned operands with even leading dimensions.
int tma_ready_operand(const double** p, int64_t* ld, int64_t rows, int64_t cols, double* scratch, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(*p) % 16 == 0 && *ld % 2 == 0) return OK;
  CANDMC_TRY(lda_copy_f64(rows, cols, *ld, rows, *p, scratch, st));
  *p = scratch;
  *ld = rows;
  return OK;
}

// CTAs of the kernel that gathers a k-chunk of B out of pinned host memory: 16 x 256 threads x 4 x 16 B = 256 KiB in flight
// keeps a PCIe link busy (~75 KiB at 50 GB/s x 1.5 us) and fits the two SMs a sweep with traffic in flight leaves free
constexpr int kHostGatherCtas = 16;

// the device-side address of pinned (page-locked, mapped) host memory, or nullptr for pageable memory
const double* device_alias_of_pinned(const double* host) {
  if (host == nullptr || !runtime().host_gather) return nullptr;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) return nullptr;
  return static_cast<const double*>(attr.devicePointer);
}

// number of k-chunks a sweep over b-wide panels uses (callers that upload operands chunk-wise must agree with it)
int sweep_chunks(int64_t b, char tA, char tB, const candmc_comm* row, const candmc_comm* col) {
  const bool need_comm = row->size > 1 || col->size > 1;
  return (is_n(tA) && is_n(tB) && need_comm) ? pick_chunks(b) : 1;
}

// Host -> device upload of an NN operand pair on `h2d`.  B goes first and whole: a k-chunk of B is a ROW slab, and a 2-D DMA
// of 16 KiB-wide rows is descriptor-bound (measured: it made the N = 2 end-to-end step 0.5 s slower than staging the block
// whole), so B is moved with full-height columns (one event for every chunk) and re-laid out per chunk on the device by the
// pack kernel.  A follows in k-chunks (contiguous column slabs), one event each, so the multiply of chunk t starts while
// chunk t+1 is still on the wire.
//
// PINNED host B (what `b_zero_copy` says; cudaHostAlloc / cudaHostRegister memory is addressable from the device): no copy of
// the whole block in front of the first multiply at all — every k-chunk is gathered STRAIGHT OUT OF HOST MEMORY by the pack
// kernel (coalesced 16-byte reads over PCIe, a capped grid that fits the SMs the multiplies leave free) and lands chunk-major,
// ready to send and to multiply; B's chunk t and A's chunk t go up back to back, so the first multiply starts after 1/nchunks
// of the two blocks instead of after all of B (2 GiB = 40-90 ms of PCIe at the headline sizes).
int upload_chunks(const double* hA, int64_t lda, const double* hB, int64_t ldb, int64_t rows, int64_t cols, int64_t k,
                  int nchunks, double* dA, double* dB, cudaStream_t h2d, std::vector<cudaEvent_t>* a_ready,
                  std::vector<cudaEvent_t>* b_ready, const double* b_zero_copy) {
  const int64_t kc = k / nchunks;
  a_ready->assign(nchunks, nullptr);
  b_ready->assign(nchunks, nullptr);
  if (hB && b_zero_copy) {
    for (int t = 0; t < nchunks; ++t) {
      CANDMC_TRY(lda_copy_f64_capped(kc, cols, ldb, kc, b_zero_copy + t * kc, dB + t * kc * cols, h2d, kHostGatherCtas));
      (*b_ready)[t] = g_events.get();
      CANDMC_CHECK((*b_ready)[t] != nullptr, "event pool exhausted");
      CANDMC_CUDA(cudaEventRecord((*b_ready)[t], h2d));
      if (hA) {
        CANDMC_CUDA(cudaMemcpy2DAsync(dA + t * kc * rows, rows * 8, hA + t * kc * lda, lda * 8, rows * 8, kc,
                                      cudaMemcpyHostToDevice, h2d));
        (*a_ready)[t] = g_events.get();
        CANDMC_CHECK((*a_ready)[t] != nullptr, "event pool exhausted");
        CANDMC_CUDA(cudaEventRecord((*a_ready)[t], h2d));
      }
    }
    return OK;
  }
  if (hB) {
    // opt-in (candmc_set_b_first_chunk_early; only reached for pageable memory or with the host gather off): the rows of the first k-chunk go ahead in their own 2-D copy
    // (narrow rows, but only 1/nchunks of B), so the first multiply does not wait for the whole block
    const int64_t k0 = (runtime().b_first_chunk_early && nchunks > 1) ? kc : 0;
    if (k0 > 0) {
      CANDMC_CUDA(cudaMemcpy2DAsync(dB, k * 8, hB, ldb * 8, k0 * 8, cols, cudaMemcpyHostToDevice, h2d));
      cudaEvent_t e0 = g_events.get();
      CANDMC_CHECK(e0 != nullptr, "event pool exhausted");
      CANDMC_CUDA(cudaEventRecord(e0, h2d));
      (*b_ready)[0] = e0;
    }
    CANDMC_CUDA(cudaMemcpy2DAsync(dB + k0, k * 8, hB + k0, ldb * 8, (k - k0) * 8, cols, cudaMemcpyHostToDevice, h2d));  // ld = k
    cudaEvent_t e = g_events.get();
    CANDMC_CHECK(e != nullptr, "event pool exhausted");
    CANDMC_CUDA(cudaEventRecord(e, h2d));
    for (int t = (k0 > 0 ? 1 : 0); t < nchunks; ++t) (*b_ready)[t] = e;
  }
  for (int t = 0; t < nchunks; ++t) {
    if (hA) {
      CANDMC_CUDA(cudaMemcpy2DAsync(dA + t * kc * rows, rows * 8, hA + t * kc * lda, lda * 8, rows * 8, kc,
                                    cudaMemcpyHostToDevice, h2d));
      (*a_ready)[t] = g_events.get();
      CANDMC_CHECK((*a_ready)[t] != nullptr, "event pool exhausted");
      CANDMC_CUDA(cudaEventRecord((*a_ready)[t], h2d));
    }
  }
  return OK;
}

// ---- one launch group: k-chunks [lo, lo + mg) of a panel (or of a k-slice), multiplied onto C ----------------------------
// The chunks go in ONE launch when their operands line up: A's chunks are consecutive column slabs of one matrix, B's are
// either consecutive row slabs of one matrix (the panel's root multiplies out of the caller's block) or the chunk-major slots
// the panel travels in (one tensor map over all of them, gemm_f64_bchunked).  Otherwise one launch per chunk.  With `slabs`
// the group is multiplied column slab by column slab and slab_done(c0, w) is called behind each slab's last launch.
struct GroupLaunch {
  char tA, tB;
  int64_t b, kc;
  double* C;
  int64_t ldC;
  cudaStream_t compute;
  FusedParams* fused = nullptr;   // the group is the single chunk whose epilogue performs the depth sum
  double* fused_out = nullptr;
  int64_t fused_ldout = 0;
  double* scratchA = nullptr;     // kc * b doubles each: TMA-ready copies of the fused chunk's operands if needed
  double* scratchB = nullptr;
};

int multiply_group(const GroupLaunch& g, int mg, const std::vector<const double*>& pa, const std::vector<int64_t>& lda,
                   const std::vector<const double*>& pb, const std::vector<int64_t>& ldb, double beta0,
                   const std::vector<int64_t>* slabs, const std::function<int(int64_t, int64_t)>& slab_done) {
  const int64_t b = g.b, kc = g.kc;
  const bool nn = is_n(g.tA) && is_n(g.tB);
  bool a_one = true, b_plain = true, b_chunked = true;
  for (int u = 0; u < mg; ++u) {
    a_one = a_one && lda[u] == lda[0] && pa[u] == pa[0] + u * kc * lda[0];
    b_plain = b_plain && ldb[u] == ldb[0] && pb[u] == pb[0] + u * kc;
    b_chunked = b_chunked && ldb[u] == kc && pb[u] == pb[0] + u * kc * b;
  }
  if (mg == 1) b_chunked = false;   // a single chunk is a plain matrix whatever its leading dimension
  const bool one_launch = mg == 1 || (nn && a_one && ((b_chunked && gemm_f64_bchunked_ok(pa[0], lda[0], pb[0], b, mg * kc, kc)) ||
                                                      (b_plain && !b_chunked)));
  // column ranges of this group's launches: the whole width, or the slabs of the early finalisation
  std::vector<std::pair<int64_t, int64_t>> cols;
  if (slabs != nullptr) {
    int64_t c0 = 0;
    for (int64_t w : *slabs) { cols.push_back({c0, w}); c0 += w; }
  } else {
    cols.push_back({0, b});
  }
  for (const auto& cw : cols) {
    const int64_t c0 = cw.first, w = cw.second;
    double* Cs = g.C + c0 * g.ldC;
    if (g.fused != nullptr) {
      // operands of the fused chunk must be TMA-able (16-byte aligned, even leading dimension): repack if they are not
      CANDMC_CHECK(mg == 1 && slabs == nullptr, "the fused depth sum takes one chunk over the full width");
      const double* fa = pa[0]; const double* fb = pb[0];
      int64_t flda = lda[0], fldb = ldb[0];
      CANDMC_TRY(tma_ready_operand(&fa, &flda, is_t(g.tA) ? kc : b, is_t(g.tA) ? b : kc, g.scratchA, g.compute));
      CANDMC_TRY(tma_ready_operand(&fb, &fldb, is_t(g.tB) ? b : kc, is_t(g.tB) ? kc : b, g.scratchB, g.compute));
      g.fused->Cin = g.C;
      g.fused->ldin = g.ldC;
      CANDMC_TRY(gemm_f64_fused(g.tA, g.tB, b, b, kc, 1.0, fa, flda, fb, fldb, beta0, g.fused_out, g.fused_ldout, g.compute, g.fused));
    } else if (one_launch && mg > 1 && b_chunked) {
      if (w == b) CANDMC_TRY(gemm_f64_bchunked('N', b, b, mg * kc, 1.0, pa[0], lda[0], pb[0], kc, beta0, Cs, g.ldC, g.compute));
      else CANDMC_TRY(gemm_f64_bchunked_cols('N', b, w, mg * kc, 1.0, pa[0], lda[0], pb[0], kc, b, c0, beta0, Cs, g.ldC, g.compute));
      runtime().merged_chunked++;
    } else if (one_launch) {
      // one chunk, or several that form one plain matrix
      const int64_t kk = (nn ? mg : 1) * kc;
      const double* pbs = is_t(g.tB) ? pb[0] + c0 : pb[0] + c0 * ldb[0];
      CANDMC_TRY(gemm_f64(g.tA, g.tB, b, w, kk, 1.0, pa[0], lda[0], pbs, ldb[0], beta0, Cs, g.ldC, g.compute));
      if (mg > 1) runtime().merged_plain++;
    } else {
      // layouts that do not line up (odd leading dimensions, operands staged whole): one launch per chunk
      for (int u = 0; u < mg; ++u)
        CANDMC_TRY(gemm_f64('N', 'N', b, w, kc, 1.0, pa[u], lda[u], pb[u] + c0 * ldb[u], ldb[u], (u == 0) ? beta0 : 1.0, Cs, g.ldC,
                            g.compute));
    }
    if (slabs != nullptr) CANDMC_TRY(slab_done(c0, w));
  }
  return OK;
}

// the launch groups of `nchunks` k-chunks whose operands are still coming up from host memory (about six times slower than
// they are multiplied): chunk 0, chunks 1-2, the rest; `lim` < nchunks leaves the chunks from lim on to themselves
void host_upload_groups(int nchunks, int lim, std::vector<int>* grp_hi) {
  grp_hi->resize(nchunks);
  for (int t = 0; t < nchunks; ++t) (*grp_hi)[t] = t + 1;
  if (lim > 2) {
    (*grp_hi)[1] = std::min(3, lim);
    if (lim > 3) (*grp_hi)[3] = lim;
  }
}

// Which k-chunks of a panel go into one launch (summa_sweep has the reasoning): grp_hi[t] = end of the group that starts at
// chunk t.  Pure arithmetic, exposed to the tests as candmc_debug_launch_groups.
void plan_launch_groups(int nchunks, int mode, bool first_panel, bool last_panel, bool host_ops, bool all_dma, bool fused, bool nn,
                        std::vector<int>* grp_hi) {
  grp_hi->resize(nchunks);
  for (int t = 0; t < nchunks; ++t) (*grp_hi)[t] = t + 1;
  if (mode <= 0 || nchunks <= 2 || !nn) return;
  const int lim = (last_panel && fused) ? nchunks - 1 : nchunks;   // the fused depth sum's chunk keeps its own launch
  if (mode == 2) {
    if (first_panel && host_ops) {
      host_upload_groups(nchunks, lim, grp_hi);
    } else if (!first_panel && all_dma) {
      (*grp_hi)[0] = lim;
    } else if (lim > 1) {
      (*grp_hi)[1] = lim;
    }
  } else if (mode == 3 && !host_ops) {
    for (int lo = 2, sz = 2; lo < lim; lo += sz, sz *= 2) (*grp_hi)[lo] = std::min(lim, lo + sz);
  } else if (mode == 1 && !host_ops && last_panel && !first_panel && !fused) {
    (*grp_hi)[0] = nchunks - 1;
  }
}

// column slabs a b-wide C block is finalised in: graduated b/2, b/4, b/8, b/8 (the wide first slab has the rest of the multiply
// to leave under, only the narrow last one is exposed), equal slabs when b does not divide that way, none when it is too ragged
std::vector<int64_t> fin_slab_widths(int64_t b, int fin_slabs) {
  std::vector<int64_t> w;
  if (fin_slabs <= 1) return w;
  if (b % 1024 == 0) w = {b / 2, b / 4, b / 8, b / 8};
  else if (b % fin_slabs == 0) w.assign(fin_slabs, b / fin_slabs);
  return w;
}

int summa_sweep(SummaArgs& a) {
  const int64_t b = a.b, bb = b * b;
  const int my_col = a.row->rank, my_row = a.col->rank;
  cudaStream_t comm = runtime().comm_stream;
  double* packA = a.ws;
  double* locB = a.ws + bb;
  double* bufA = a.ws + 2 * bb;
  double* bufB = a.ws + 3 * bb;
  const bool need_comm = a.row->size > 1 || a.col->size > 1;
  // transposed panels are moved whole (the flags only reach the local GEMM); a 1x1 grid has nothing to pipeline
  const int nchunks = sweep_chunks(b, a.tA, a.tB, a.row, a.col);
  const int64_t kc = b / nchunks;
  auto wait_ready = [](cudaStream_t s, const std::vector<cudaEvent_t>* ev, int t) -> int {
    if (ev && t < (int)ev->size() && (*ev)[t]) CANDMC_CUDA(cudaStreamWaitEvent(s, (*ev)[t], 0));
    return OK;
  };

  // panels by copy engines into peer windows (transport.h; the default) instead of ncclBroadcast, per grid axis
  PanelTransport *tr_row = nullptr, *tr_col = nullptr;
  if (runtime().panel_transport && need_comm && (a.i1 - a.i0) * nchunks <= kPanelMaxOps) {
    if (a.row->size > 1) CANDMC_TRY(panel_transport_get(a.row, (a.i1 - a.i0) * bb, &tr_row));
    if (a.col->size > 1) CANDMC_TRY(panel_transport_get(a.col, (a.i1 - a.i0) * bb, &tr_col));
    if (tr_row) panel_transport_begin(tr_row);
    if (tr_col) panel_transport_begin(tr_col);
  }
  const bool all_dma = (a.row->size == 1 || tr_row) && (a.col->size == 1 || tr_col);   // no NCCL kernel in this sweep
  if (need_comm) CANDMC_TRY(stream_wait(comm, a.compute));  // inputs (and earlier users of ws) are ready
  // NCCL moves data with SM-resident kernels, and the persistent GEMM owns every SM it is given (all registers, 193 KiB
  // smem), so a broadcast enqueued while a GEMM runs would only start when that GEMM ends.  While panels are in flight
  // the GEMMs therefore leave as many SMs free as the background communicators may use.
  // (NCCL path with launch groups — the fallback when peer windows are unavailable: a merged launch needs the rest of its panel
  // within one chunk's multiply, which the 2-CTA background communicators cannot deliver (127 TFLOP/s on 4 B200s); the
  // full-width communicators between the launches can, with no SM held back: 138 TFLOP/s, profiles/r02_4gpu/ "bgctas_0")
  const bool nccl_full_width = need_comm && !all_dma && runtime().merge_panels == 2;
  ReserveGuard reserve_guard((need_comm && !all_dma && !nccl_full_width) ? std::max(runtime().bg_max_ctas, runtime().gemm_reserve_sms)
                                                                         : runtime().gemm_reserve_sms);
  std::vector<cudaEvent_t> done_prev(nchunks, nullptr);
  bool first = a.first_beta_zero;
  // Operands that are already on the device are packed for sending BEFORE the first multiply is enqueued (A out of its
  // leading dimension, B chunk-major; a rank is the root of at most one A and one B panel per sweep): the pack kernels need
  // SMs, and once a persistent GEMM owns all of them a pack enqueued behind it would only run when that GEMM ends — the
  // first 4-GPU timeline of the copy-engine transport showed exactly that, 0.4 ms per chunk and 7 ms in front of a merged
  // launch (profiles/r02_4gpu/).  Operands still coming up from host memory land send-ready (ld = b, B chunk-major).
  const bool prepacked = need_comm && a.a_ready == nullptr && a.b_ready == nullptr;
  if (prepacked) {
    if (a.row->size > 1 && my_col >= a.i0 && my_col < a.i1 && a.ldA != b)
      CANDMC_TRY(lda_copy_f64(b, b, a.ldA, b, a.myA, packA, comm));
    if (a.col->size > 1 && my_row >= a.i0 && my_row < a.i1 && !a.b_chunk_major && (nchunks > 1 || a.ldB != b))
      for (int t = 0; t < nchunks; ++t) CANDMC_TRY(lda_copy_f64(kc, b, a.ldB, kc, a.myB + t * kc, locB + t * kc * b, comm));
  }
  for (int i = a.i0; i < a.i1; ++i) {
    const bool rootA = (my_col == i), rootB = (my_row == i);
    std::vector<cudaEvent_t> ready(nchunks, nullptr);
    // ---- communication for panel i, chunk by chunk, on the comm stream ----
    if (need_comm) {
      for (int t = 0; t < nchunks; ++t) {
        // the very first chunk has nothing to hide under: full-width communicator (as is everything when launches are merged)
        const bool bg = !(i == a.i0 && t == 0) && !nccl_full_width;
        // buf slot t is free again (the copy-engine transport gives every panel and chunk of a sweep its own window slot)
        if (done_prev[t] && !all_dma) CANDMC_CUDA(cudaStreamWaitEvent(comm, done_prev[t], 0));
        const int op = (i - a.i0) * nchunks + t;   // transport slot of this (panel, chunk): its own, never reused in a sweep
        if (a.row->size > 1) {
          double* slot = bufA + t * kc * b;
          if (rootA) {
            CANDMC_TRY(wait_ready(comm, a.a_ready, t));
            const double* src = a.myA + t * kc * a.ldA;  // column slab: contiguous iff ldA == b
            if (a.ldA != b) {
              if (!prepacked) CANDMC_TRY(lda_copy_f64(b, kc, a.ldA, b, src, packA + t * kc * b, comm));
              src = packA + t * kc * b;
            }
            if (tr_row) CANDMC_TRY(panel_transport_send(tr_row, a.row, op, op * kc * b, src, kc * b, comm));
            else CANDMC_TRY(comm_bcast(a.row, src, const_cast<double*>(src), kc * b, i, comm, bg));
          } else if (!tr_row) {
            CANDMC_TRY(comm_bcast(a.row, slot, slot, kc * b, i, comm, bg));
          }
        }
        if (a.col->size > 1) {
          double* slot = bufB + t * kc * b;
          if (rootB) {
            CANDMC_TRY(wait_ready(comm, a.b_ready, t));
            const double* src = a.b_chunk_major ? a.myB + t * kc * b : a.myB + t * kc;  // row slab of B
            if (!a.b_chunk_major && (nchunks > 1 || a.ldB != b)) {
              if (!prepacked) CANDMC_TRY(lda_copy_f64(kc, b, a.ldB, kc, src, locB + t * kc * b, comm));  // chunk-major, ld = kc
              src = locB + t * kc * b;
            }
            if (tr_col) CANDMC_TRY(panel_transport_send(tr_col, a.col, op, op * kc * b, src, kc * b, comm));
            else CANDMC_TRY(comm_bcast(a.col, src, const_cast<double*>(src), kc * b, i, comm, bg));
          } else if (!tr_col) {
            CANDMC_TRY(comm_bcast(a.col, slot, slot, kc * b, i, comm, bg));
          }
        }
        ready[t] = g_events.get();
        CANDMC_CHECK(ready[t] != nullptr, "event pool exhausted");
        CANDMC_CUDA(cudaEventRecord(ready[t], comm));
      }
    }
    // ---- multiplies for panel i on the compute stream ----
    auto operands = [&](int t, const double** pa, int64_t* lda, const double** pb, int64_t* ldb) -> int {
      if (ready[t]) CANDMC_CUDA(cudaStreamWaitEvent(a.compute, ready[t], 0));
      if (rootA || a.row->size == 1) {
        CANDMC_TRY(wait_ready(a.compute, a.a_ready, t));
        *pa = a.myA + t * kc * a.ldA;
        *lda = a.ldA;
      } else if (tr_row) {
        const int op = (i - a.i0) * nchunks + t;
        CANDMC_TRY(panel_transport_wait(tr_row, a.row, op, op * kc * b, a.compute, pa));
        *lda = b;
      } else {
        *pa = bufA + t * kc * b;
        *lda = b;
      }
      if (rootB || a.col->size == 1) {
        CANDMC_TRY(wait_ready(a.compute, a.b_ready, t));
        *pb = a.b_chunk_major ? a.myB + t * kc * b : a.myB + t * kc;
        *ldb = a.b_chunk_major ? kc : a.ldB;
      } else if (tr_col) {
        const int op = (i - a.i0) * nchunks + t;
        CANDMC_TRY(panel_transport_wait(tr_col, a.col, op, op * kc * b, a.compute, pb));
        *ldb = kc;  // chunk-major, as the root packed it
      } else {
        *pb = bufB + t * kc * b;
        *ldb = kc;  // chunk-major
      }
      return OK;
    };
    // ---- which k-chunks go into one launch ----
    // Every launch pays its own epilogue (a read-modify-write of the C tile) and its own tail wave, so the chunks of a panel
    // are multiplied in as few launches as the arrival of the data allows:
    //   * the first panel of a sweep has nothing to hide its transfer under: chunk 0 alone (its transfer is the exposed start
    //     of the sweep), then everything else in one launch — over NVLink the rest of the panel arrives while chunk 0 is being
    //     multiplied.  Operands that are still coming up from host memory arrive ~6x slower (PCIe): chunk 0, chunks 1-2, the rest.
    //   * later panels were moved under the previous panel's multiplies: with the copy-engine transport (own window slot per
    //     panel and chunk) the whole panel is there — one launch; with ncclBroadcast the chunks' buffer slots only come free as
    //     the previous panel's launches finish — chunk 0, then the rest (its broadcasts run under chunk 0's multiply).
    //   * the fused depth sum keeps the last chunk apart: its launch is the one with the reducing epilogue.
    //   * merge_panels = 0 restores one launch per chunk (tests, A/B measurements); modes 1 and 3 are the earlier experiments.
    const bool last_panel = (i + 1 == a.i1), first_panel = (i == a.i0);
    const bool host_ops = (a.a_ready != nullptr || a.b_ready != nullptr);
    const bool nn = is_n(a.tA) && is_n(a.tB);
    std::vector<int> grp_hi;   // chunks [t, grp_hi[t]) go in one launch when t starts a group
    plan_launch_groups(nchunks, runtime().merge_panels, first_panel, last_panel, host_ops, all_dma, a.fused != nullptr, nn, &grp_hi);
    // Early finalisation (host C): the LAST launch group of the sweep is cut into column slabs — each slab still covers all of
    // the group's k, so it is a handful of large launches, not one per chunk and slab — and slab_done() ships a slab (depth
    // sum, download) while the next ones multiply.  Graduated widths b/2, b/4, b/8, b/8: the wide first slab has the rest of
    // the multiply to leave under, only the narrow last one is exposed.
    std::vector<int64_t> slab_w;
    if (last_panel && a.slab_done && a.fused == nullptr && nn) slab_w = fin_slab_widths(b, a.fin_slabs);
    int last_group_lo = 0;
    for (int t = 0; t < nchunks; t = grp_hi[t]) last_group_lo = t;

    for (int t = 0; t < nchunks; t = grp_hi[t]) {
      const int mg_lo = t, mg = grp_hi[t] - t;
      const bool slabbed = !slab_w.empty() && mg_lo == last_group_lo;
      std::vector<const double*> pa(mg), pb(mg);
      std::vector<int64_t> lda(mg), ldb(mg);
      // the waits operands() enqueues are the ones the launches below need, merged or not
      for (int u = 0; u < mg; ++u) CANDMC_TRY(operands(mg_lo + u, &pa[u], &lda[u], &pb[u], &ldb[u]));
      const bool fused_here = last_panel && mg_lo + mg == nchunks && a.fused != nullptr;
      GroupLaunch gl;
      gl.tA = a.tA; gl.tB = a.tB; gl.b = b; gl.kc = kc; gl.C = a.C; gl.ldC = a.ldC; gl.compute = a.compute;
      gl.fused = fused_here ? a.fused : nullptr; gl.fused_out = a.fused_out; gl.fused_ldout = a.fused_ldout;
      gl.scratchA = bufA + mg_lo * kc * b; gl.scratchB = bufB + mg_lo * kc * b;
      CANDMC_TRY(multiply_group(gl, mg, pa, lda, pb, ldb, first ? 0.0 : 1.0, slabbed ? &slab_w : nullptr, a.slab_done));
      first = false;
      if (need_comm && !last_panel) {   // the buffer slots of the group's chunks come free together
        cudaEvent_t e = g_events.get();
        CANDMC_CHECK(e != nullptr, "event pool exhausted");
        CANDMC_CUDA(cudaEventRecord(e, a.compute));
        for (int u = 0; u < mg; ++u) done_prev[mg_lo + u] = e;
      }
    }
  }
  // every multiply of the sweep is enqueued: the peers may reuse the window halves this call read from
  if (tr_row) CANDMC_TRY(panel_transport_end(tr_row, a.row, a.compute, comm));
  if (tr_col) CANDMC_TRY(panel_transport_end(tr_col, a.col, a.compute, comm));
  return OK;
}

// ---- the cut of the host-streamed multiply below into column panels of C and of its first panel into k-chunks ----
// Uniform (NP equal panels, NP equal chunks: what B200s have measured with NP = 8) or graduated (the automatic choice for large
// products).  What stays exposed is the upload in front of the first multiply, any upload the first panel's multiplies fail
// to cover, and the download behind the last multiply.  The first panel has to hide the upload of ALL of A plus its own B
// row slabs (2-D copies with one narrow row per column: about 2 us per row whatever its width), and chunk t+1 can only be
// covered by the multiply of chunk t — so chunks that grow fast make the panel upload-bound (doubling chunks are WORSE than
// equal ones; tools/host_pipeline_model.py).  The graduated cut therefore makes the first panel twice as wide as the others
// (n/4: twice the multiply time per uploaded byte of A), starts its k-chunks at k/16 and lets them grow by a tenth each, and
// shrinks the last panels (n/16, n/32, n/32) so that only 1/32 of C is downloaded after the last multiply.  Every piece but
// the last is a multiple of the CTA tile / the k-tile.
void host_pipeline_cut(int64_t n, int64_t k, int panels, std::vector<int64_t>* widths, std::vector<int64_t>* kchunks) {
  auto round_up = [](int64_t x, int64_t q) { return (x + q - 1) / q * q; };
  widths->clear();
  kchunks->clear();
  const bool graduated = panels < 0 || (panels == 0 && n >= 8192 && k >= 8192);
  if (!graduated) {
    const int NP = panels > 0 ? panels : 8;
    const int64_t nb = round_up((n + NP - 1) / NP, 128), kc = round_up((k + NP - 1) / NP, 16);
    for (int64_t c0 = 0; c0 < n; c0 += nb) widths->push_back(std::min(nb, n - c0));
    for (int64_t k0 = 0; k0 < k; k0 += kc) kchunks->push_back(std::min(kc, k - k0));
    return;
  }
  const int64_t base = round_up((n + 7) / 8, 128);
  const int64_t t2 = std::max<int64_t>(128, round_up(base / 2, 128)), t4 = std::max<int64_t>(128, round_up(base / 4, 128));
  const int64_t tail = t2 + 2 * t4;
  int64_t rem = n;
  while (rem - base >= tail) {
    widths->push_back(base);
    rem -= base;
  }
  if (widths->size() >= 3) {   // the first panel: two body panels in one
    widths->erase(widths->begin());
    (*widths)[0] += base;
  }
  const int64_t mid = rem > tail ? (rem - tail) / 128 * 128 : 0;
  if (mid > 0) {
    widths->push_back(mid);
    rem -= mid;
  }
  for (int64_t w : {t2, t4}) {
    if (rem > w) {
      widths->push_back(w);
      rem -= w;
    }
  }
  if (rem > 0) widths->push_back(rem);   // <= t4 + 127 <= base
  const int64_t first = round_up((k + 15) / 16, 16);
  double size = static_cast<double>(first);
  for (int64_t k0 = 0; k0 < k;) {
    int64_t kc = std::min(round_up(static_cast<int64_t>(size), 16), k - k0);
    if (k - k0 - kc < first / 2) kc = k - k0;   // no crumb at the end
    kchunks->push_back(kc);
    k0 += kc;
    size *= 1.1;
  }
}

// ---- host-resident operands on a 1x1 grid: stream the multiply through PCIe -----------------------------------------
// C(m x n) = A(m x k) * B(k x n) with A, B, C in HOST memory (what the reference's callers own).  Instead of
// "copy everything in, multiply, copy everything out" the product is cut into column panels of C: panel j+1's slice of B
// is uploaded and panel j-1's slice of C is downloaded while panel j multiplies; the first panel is additionally cut
// along k so that its GEMMs start as soon as the first column slab of A has landed.  Only the first A slab and the last
// C panel are exposed (host_pipeline_cut above decides how large those are).  Three streams: H2D (aux), compute (caller's), D2H (comm — idle on a 1x1 grid).
int host_pipelined_gemm_nn(int64_t m, int64_t n, int64_t k, const double* hA, int64_t lda, const double* hB, int64_t ldb,
                           double* hC, int64_t ldc, cudaStream_t st) {
  std::vector<int64_t> widths, kchunks;
  host_pipeline_cut(n, k, runtime().host_pipeline_panels, &widths, &kchunks);
  const int npanels = (int)widths.size(), nchunks = (int)kchunks.size();
  const int64_t nb = *std::max_element(widths.begin(), widths.end());   // the double buffers hold the widest panel
  cudaStream_t h2d = runtime().aux_stream, d2h = runtime().comm_stream;
  void* wsv = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * (m * k + 2 * k * nb + 2 * m * nb + 8), &wsv));
  double* dA = static_cast<double*>(wsv);
  double* dB[2] = {dA + m * k + (m * k & 1), nullptr};
  dB[1] = dB[0] + k * nb;
  double* dC[2] = {dB[1] + k * nb, nullptr};
  dC[1] = dC[0] + m * nb;
  CANDMC_TRY(stream_wait(h2d, st));
  CANDMC_TRY(stream_wait(d2h, st));
  std::vector<cudaEvent_t> g_done(npanels, nullptr), c_free(npanels, nullptr);
  auto ev = [&](cudaEvent_t* e, cudaStream_t s) -> int {
    *e = g_events.get();
    CANDMC_CHECK(*e != nullptr, "event pool exhausted");
    CANDMC_CUDA(cudaEventRecord(*e, s));
    return OK;
  };
  int64_t c0 = 0;
  for (int j = 0; j < npanels; ++j) {
    const int slot = j & 1;
    const int64_t nbj = widths[j];
    if (j == 0) {
      int64_t k0 = 0;
      for (int t = 0; t < nchunks; ++t) {
        const int64_t kct = kchunks[t];
        CANDMC_CUDA(cudaMemcpy2DAsync(dA + k0 * m, m * 8, hA + k0 * lda, lda * 8, m * 8, kct, cudaMemcpyHostToDevice, h2d));
        CANDMC_CUDA(cudaMemcpy2DAsync(dB[0] + k0, k * 8, hB + k0, ldb * 8, kct * 8, nbj, cudaMemcpyHostToDevice, h2d));
        cudaEvent_t ready;
        CANDMC_TRY(ev(&ready, h2d));
        CANDMC_CUDA(cudaStreamWaitEvent(st, ready, 0));
        CANDMC_TRY(gemm_f64('N', 'N', m, nbj, kct, 1.0, dA + k0 * m, m, dB[0] + k0, k, t ? 1.0 : 0.0, dC[0], m, st));
        k0 += kct;
      }
    } else {
      if (j >= 2) CANDMC_CUDA(cudaStreamWaitEvent(h2d, g_done[j - 2], 0));  // dB[slot] no longer read
      CANDMC_CUDA(cudaMemcpy2DAsync(dB[slot], k * 8, hB + c0 * ldb, ldb * 8, k * 8, nbj, cudaMemcpyHostToDevice, h2d));
      cudaEvent_t ready;
      CANDMC_TRY(ev(&ready, h2d));
      CANDMC_CUDA(cudaStreamWaitEvent(st, ready, 0));
      if (j >= 2) CANDMC_CUDA(cudaStreamWaitEvent(st, c_free[j - 2], 0));   // dC[slot] has been downloaded
      CANDMC_TRY(gemm_f64('N', 'N', m, nbj, k, 1.0, dA, m, dB[slot], k, 0.0, dC[slot], m, st));
    }
    CANDMC_TRY(ev(&g_done[j], st));
    CANDMC_CUDA(cudaStreamWaitEvent(d2h, g_done[j], 0));
    CANDMC_CUDA(cudaMemcpy2DAsync(hC + c0 * ldc, ldc * 8, dC[slot], m * 8, m * 8, nbj, cudaMemcpyDeviceToHost, d2h));
    CANDMC_TRY(ev(&c_free[j], d2h));
    c0 += nbj;
  }
  CANDMC_TRY(stream_wait(st, d2h));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int check_grid_args(const candmc_ctb_args_t* args, candmc_comm* row, candmc_comm* col, int64_t* b_out) {
  CANDMC_CHECK(args != nullptr && row != nullptr && col != nullptr, "null argument");
  CANDMC_CHECK(row->size == col->size, "processor grid must be square (np_row=%d, np_col=%d)", col->size,
               row->size);  // ASSERT(np_row == np_col), summa.cxx:45
  CANDMC_CHECK(args->n > 0 && args->n % row->size == 0, "n=%lld must be a positive multiple of the grid dimension %d",
               (long long)args->n, row->size);  // ASSERT(n % np_row == 0), summa.cxx:46
  CANDMC_CHECK((is_t(args->trans_A) || is_n(args->trans_A)) && (is_t(args->trans_B) || is_n(args->trans_B)),
               "bad transpose flags");
  *b_out = args->n / row->size;
  return OK;
}

// ---- lower-triangular solve W <- T^-1 W (T b x b lower, non-unit; W b x kb) — cdtrsm('L','L','N','N') of qr_2d.cxx:271
constexpr int TRSM_NC = 8;    // right-hand sides per CTA
constexpr int TRSM_NB = 32;   // diagonal block
__global__ void __launch_bounds__(256)
trsm_llnn_kernel(int b, int kb, const double* __restrict__ T, int64_t ldt, double* __restrict__ W, int64_t ldw) {
  extern __shared__ double sw[];  // b x TRSM_NC, column-major
  const int c0 = blockIdx.x * TRSM_NC;
  const int nc = min(TRSM_NC, kb - c0);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < b * nc; e += nt) sw[e] = W[(e % b) + static_cast<int64_t>(c0 + e / b) * ldw];
  __syncthreads();
  for (int i0 = 0; i0 < b; i0 += TRSM_NB) {
    const int ib = min(TRSM_NB, b - i0);
    for (int r = i0; r < i0 + ib; ++r) {
      if (tid < nc) sw[r + tid * b] /= T[r + static_cast<int64_t>(r) * ldt];
      __syncthreads();
      const int rem = i0 + ib - r - 1;
      for (int e = tid; e < rem * nc; e += nt) {
        const int rr = r + 1 + e % rem, c = e / rem;
        sw[rr + c * b] -= T[rr + static_cast<int64_t>(r) * ldt] * sw[r + c * b];
      }
      __syncthreads();
    }
    const int i1 = i0 + ib, rest = b - i1;
    for (int e = tid; e < rest * nc; e += nt) {
      const int row = i1 + e % rest, c = e / rest;
      double acc = 0.0;
      for (int p = 0; p < ib; ++p) acc += T[row + static_cast<int64_t>(i0 + p) * ldt] * sw[i0 + p + c * b];
      sw[row + c * b] -= acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < b * nc; e += nt) W[(e % b) + static_cast<int64_t>(c0 + e / b) * ldw] = sw[e];
}

// Variant with one WARP per right-hand side (the default since round 2; candmc_set_trsm_variant(0) selects the kernel above).  The kernel above meets at
// a block barrier twice per row of T (1024 barriers for b = 512); here a column never leaves its warp: lane r of warp w owns
// row i0 + r of column c0 + w of the current 32-row block, the part of the solution that is already final is applied
// left-looking from shared memory (T streamed through a 32 x 128 tile that the eight warps of the CTA share), and the
// 32 x 32 diagonal block is solved with shuffles.  Block barriers: two per T tile, 104 for b = 512.
constexpr int TRSMW_TILE = 128;
__global__ void __launch_bounds__(256)
trsm_llnn_warp_kernel(int b, int kb, const double* __restrict__ T, int64_t ldt, double* __restrict__ W, int64_t ldw) {
  extern __shared__ double sx[];                         // 8 columns x b (the solution as it becomes final)
  __shared__ double tt[32][TRSMW_TILE + 1];              // T[i0 + r][j0 + c]; 129 = 1 mod 16: conflict-free across lanes
  __shared__ double td[32][33];                          // the diagonal block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * TRSM_NC + warp;
  const bool live = col < kb;
  double* x = sx + static_cast<int64_t>(warp) * b;
  double* wcol = W + static_cast<int64_t>(live ? col : 0) * ldw;
  for (int i0 = 0; i0 < b; i0 += 32) {
    const int ib = min(32, b - i0);
    double acc = 0.0;
    for (int j0 = 0; j0 < i0; j0 += TRSMW_TILE) {
      const int jb = min(TRSMW_TILE, i0 - j0);
      __syncthreads();                                   // the previous tile has been consumed by every warp
      for (int e = threadIdx.x; e < ib * jb; e += 256) {
        const int r = e % ib, c = e / ib;                // consecutive threads: consecutive rows of one column of T
        tt[r][c] = T[(i0 + r) + static_cast<int64_t>(j0 + c) * ldt];
      }
      __syncthreads();
      if (live && lane < ib)
        for (int c = 0; c < jb; ++c) acc += tt[lane][c] * x[j0 + c];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ib * ib; e += 256) {
      const int r = e % ib, c = e / ib;
      td[r][c] = T[(i0 + r) + static_cast<int64_t>(i0 + c) * ldt];
    }
    __syncthreads();
    double v = (live && lane < ib) ? wcol[i0 + lane] - acc : 0.0;
    for (int s = 0; s < ib; ++s) {                        // forward substitution inside the block, one row per lane
      const double xs = __shfl_sync(0xffffffffu, v, s) / td[s][s];
      if (lane == s) v = xs;
      else if (lane > s && lane < ib) v -= td[lane][s] * xs;
    }
    if (lane < ib) x[i0 + lane] = v;
    __syncwarp();
  }
  if (live)
    for (int r = lane; r < b; r += 32) wcol[r] = x[r];
}

int g_trsm_variant = 1;   // one warp per right-hand side: 125.4 against 121.7 TFLOP/s on config 5 (4 x B200, profiles/r02_4gpu_b/), results checked

int trsm_llnn(int64_t b, int64_t kb, const double* T, int64_t ldt, double* W, int64_t ldw, cudaStream_t st) {
  if (b <= 0 || kb <= 0) return OK;
  if (g_trsm_variant == 1 && sizeof(double) * b * TRSM_NC <= 160 * 1024) {
    const size_t smem = sizeof(double) * b * TRSM_NC;
    static size_t configured_w = 0;
    if (smem > configured_w) {   // the kernel's 41 KiB of static shared memory count towards the 48 KiB default: always opt in
      CANDMC_CUDA(cudaFuncSetAttribute(trsm_llnn_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured_w = smem;
    }
    trsm_llnn_warp_kernel<<<(int)((kb + TRSM_NC - 1) / TRSM_NC), 256, smem, st>>>((int)b, (int)kb, T, ldt, W, ldw);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
    return OK;
  }
  const size_t smem = sizeof(double) * b * TRSM_NC;
  CANDMC_CHECK(smem <= 200 * 1024, "upd_A: panel width b=%lld too large for the triangular solve kernel", (long long)b);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CANDMC_CUDA(cudaFuncSetAttribute(trsm_llnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = (int)((kb + TRSM_NC - 1) / TRSM_NC);
  trsm_llnn_kernel<<<grid, 256, smem, st>>>((int)b, (int)kb, T, ldt, W, ldw);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

}  // namespace
}  // namespace candmc

using namespace candmc;

extern "C" {

int candmc_debug_launch_groups(int nchunks, int mode, int first_panel, int last_panel, int host_ops, int all_dma, int fused, int nn,
                               int* grp_hi) {
  CANDMC_CHECK(nchunks >= 1 && nchunks <= 64 && grp_hi != nullptr, "candmc_debug_launch_groups: bad arguments");
  std::vector<int> g;
  plan_launch_groups(nchunks, mode, first_panel != 0, last_panel != 0, host_ops != 0, all_dma != 0, fused != 0, nn != 0, &g);
  std::copy(g.begin(), g.end(), grp_hi);
  return OK;
}

int candmc_debug_fin_slab_widths(int64_t b, int fin_slabs, int64_t* widths, int cap, int* count) {
  CANDMC_CHECK(widths != nullptr && count != nullptr, "candmc_debug_fin_slab_widths: null output");
  const std::vector<int64_t> w = fin_slab_widths(b, fin_slabs);
  CANDMC_CHECK((int)w.size() <= cap, "candmc_debug_fin_slab_widths: %zu slabs do not fit", w.size());
  std::copy(w.begin(), w.end(), widths);
  *count = (int)w.size();
  return OK;
}

int candmc_set_host_gather(int on) {
  runtime().host_gather = (on != 0);
  return OK;
}

int candmc_set_panel_transport(int on) {
  runtime().panel_transport = (on != 0);
  return OK;
}

unsigned long long candmc_panel_transport_sends(void) { return runtime().transport_sends; }
unsigned long long candmc_merged_panel_launches(int chunk_major_b) {
  return chunk_major_b ? runtime().merged_chunked : runtime().merged_plain;
}

int candmc_set_trsm_variant(int variant) {
  CANDMC_CHECK(variant == 0 || variant == 1, "candmc_set_trsm_variant: 0 (block barriers per row) or 1 (one warp per right-hand side)");
  g_trsm_variant = variant;
  return OK;
}

int candmc_host_pipeline_cut(int64_t n, int64_t k, int panels, int64_t* widths, int64_t* kchunks, int cap, int* npanels,
                             int* nchunks) {
  CANDMC_CHECK(n > 0 && k > 0 && widths && kchunks && npanels && nchunks, "candmc_host_pipeline_cut: bad arguments");
  std::vector<int64_t> w, c;
  host_pipeline_cut(n, k, panels, &w, &c);
  CANDMC_CHECK((int)w.size() <= cap && (int)c.size() <= cap, "candmc_host_pipeline_cut: %zu panels / %zu chunks do not fit", w.size(),
               c.size());
  std::copy(w.begin(), w.end(), widths);
  std::copy(c.begin(), c.end(), kchunks);
  *npanels = (int)w.size();
  *nchunks = (int)c.size();
  return OK;
}

int candmc_set_host_pipeline_panels(int panels) {
  CANDMC_CHECK(panels >= -1 && panels <= 64, "candmc_set_host_pipeline_panels: -1 (graduated), 0 (automatic), 1 .. 64 (uniform)");
  runtime().host_pipeline_panels = panels;
  return OK;
}

int candmc_set_host_pipeline_min(int64_t min_n) {
  CANDMC_CHECK(min_n >= 1, "candmc_set_host_pipeline_min: must be >= 1");
  runtime().host_pipeline_min = min_n;
  return OK;
}

int candmc_set_merge_last_panel(int on) { return candmc_set_merge_panels(on ? 1 : 0); }

int candmc_set_merge_panels(int mode) {
  CANDMC_CHECK(mode >= 0 && mode <= 3, "candmc_set_merge_panels: 0 (off), 1 (last panel), 2 (every panel) or 3 (doubling groups)");
  runtime().merge_panels = mode;
  return OK;
}

int candmc_set_min_kchunk(int64_t min_kchunk) {
  CANDMC_CHECK(min_kchunk >= 2, "candmc_set_min_kchunk: must be >= 2");
  runtime().min_kchunk = min_kchunk;
  return OK;
}

// ================================================================================================================
int candmc_summa(const candmc_ctb_args_t* args, const double* mat_A, const double* mat_B, double* mat_C,
                 double* buffer, candmc_comm_t* cdt_row, candmc_comm_t* cdt_col, void* stream) {
  NvtxRange nvtx_range("d2_topo_bcast_gemm");   // summa.cxx:58
  CANDMC_TRY(runtime_require());
  g_events.reset();
  int64_t b;
  CANDMC_TRY(check_grid_args(args, cdt_row, cdt_col, &b));
  // "make sure we have enough buffer space", summa.cxx:22-24,44 — kept for drop-in behaviour even though the
  // device workspace is internal
  CANDMC_CHECK(buffer == nullptr || args->buffer_size >= 4 * b * b * (int64_t)sizeof(double),
               "summa: buffer_size %lld < 4*b*b*8", (long long)args->buffer_size);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StagedMatrix sA, sB, sC;
  CANDMC_TRY(sA.open(mat_A, b, b, args->lda_A, true, st));
  CANDMC_TRY(sB.open(mat_B, b, b, args->lda_B, true, st));
  CANDMC_TRY(sC.open(mat_C, b, b, args->lda_C, false, st));
  void* ws = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * (cdt_row->size > 1 ? 4 * b * b : 2), &ws));
  SummaArgs a;
  a.tA = args->trans_A; a.tB = args->trans_B; a.b = b; a.i0 = 0; a.i1 = cdt_row->size;
  a.myA = sA.ptr(); a.ldA = sA.ld(); a.myB = sB.ptr(); a.ldB = sB.ld();
  a.C = sC.ptr(); a.ldC = sC.ld(); a.first_beta_zero = true;
  a.row = cdt_row; a.col = cdt_col; a.ws = static_cast<double*>(ws); a.compute = st;
  CANDMC_TRY(summa_sweep(a));
  CANDMC_TRY(sC.close_out(st));
  if (sA.staged() || sB.staged() || sC.staged()) CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

// ================================================================================================================
int candmc_d25_summa(const candmc_ctb_args_t* args, const double* mat_A, const double* mat_B, double* mat_C,
                     double* buffer, candmc_comm_t* cdt_row, candmc_comm_t* cdt_col, candmc_comm_t* cdt_kdir,
                     int ovp, void* stream) {
  NvtxRange nvtx_range("d25_summa_gemm");   // d25_summa.cxx:122
  CANDMC_TRY(runtime_require());
  g_events.reset();
  int64_t b;
  CANDMC_TRY(check_grid_args(args, cdt_row, cdt_col, &b));
  CANDMC_CHECK(cdt_kdir != nullptr, "d25_summa: null depth communicator");
  const int q = cdt_row->size, c = cdt_kdir->size, layer = cdt_kdir->rank;
  const bool ksplit = (q == 1 && c > 1);  // extension: 1 x 1 x c grid splits k; the reference asserts q % c == 0
  CANDMC_CHECK(ksplit || q % c == 0, "d25_summa: grid dimension %d not divisible by replication factor %d", q,
               c);  // ASSERT(np_row % c_rep == 0), d25_summa.cxx:63
  CANDMC_CHECK(!ksplit || b % c == 0, "d25_summa (1x1xc k-split): n must be divisible by c");
  const int64_t need = (ovp ? 5 : 3) * b * b * (int64_t)sizeof(double);  // buffer_space_req, d25_summa.cxx:25-31
  CANDMC_CHECK(buffer == nullptr || args->buffer_size >= need, "d25_summa: buffer_size %lld < %lld",
               (long long)args->buffer_size, (long long)need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // profiling: a zero-flop entry at the start of the call, so a timeline shows what precedes the first multiply (uploads, the
  // first panel chunk) and, through the next call's entry, what follows the last one (depth sum, download of C)
  if (runtime().profile) {
    CANDMC_TRY(profile_begin_launch(st, 0.0));
    CANDMC_TRY(profile_end_launch(st));
  }
  if (q == 1 && c == 1 && b >= runtime().host_pipeline_min && is_n(args->trans_A) && is_n(args->trans_B) && !is_device_ptr(mat_A) &&
      !is_device_ptr(mat_B) && !is_device_ptr(mat_C))
    return host_pipelined_gemm_nn(b, b, b, mat_A, args->lda_A, mat_B, args->lda_B, mat_C, args->lda_C, st);
  // Host operands (what the reference's callers own): with untransposed inputs they are uploaded in k-chunks on a copy
  // stream while the multiply is already running on the chunks that have landed (a 1x1xc grid uploads only its k-slice);
  // otherwise they are staged whole.
  // a layer only ever sends/multiplies its own A block if it owns one of the layer's panel columns (B: panel rows)
  const int pi0 = ksplit ? 0 : layer * (q / c), pi1 = ksplit ? 1 : (layer + 1) * (q / c);
  const bool skip_unused = runtime().skip_unused_uploads;
  const bool useA = !skip_unused || ksplit || (cdt_row->rank >= pi0 && cdt_row->rank < pi1);
  const bool useB = !skip_unused || ksplit || (cdt_col->rank >= pi0 && cdt_col->rank < pi1);
  if (!useA) mat_A = nullptr;
  if (!useB) mat_B = nullptr;
  const bool hostA = useA && !is_device_ptr(mat_A), hostB = useB && !is_device_ptr(mat_B);
  const bool nn = is_n(args->trans_A) && is_n(args->trans_B);
  const int64_t kloc = ksplit ? b / c : b;                       // k extent this rank touches of its own blocks
  int up_chunks = 1;
  if (nn && (hostA || hostB)) {
    // the chunking must be the one the consumer uses: the sweep's own for q > 1, the k-slice loop below for 1 x 1 x c;
    // a 1 x 1 x 1 grid below the streaming threshold multiplies in one piece, so it stages whole
    if (q > 1) up_chunks = sweep_chunks(b, args->trans_A, args->trans_B, cdt_row, cdt_col);
    else if (ksplit) for (int nc = 8; nc > 1; nc >>= 1) if (kloc % (2 * nc) == 0 && kloc / nc >= 256) { up_chunks = nc; break; }
  }
  const bool chunked = nn && (hostA || hostB) && (up_chunks > 1 || ksplit);
  std::vector<cudaEvent_t> a_ready, b_ready;
  StagedMatrix sA, sB, sC;
  const double* dA_ptr = nullptr; const double* dB_ptr = nullptr;
  int64_t dA_ld = 0, dB_ld = 0;
  bool b_chunk_major = false, host_gather_active = false;
  if (chunked) {
    void* pool = nullptr;
    const int64_t needA = hostA ? b * kloc : 0, needB = hostB ? kloc * b : 0;
    CANDMC_TRY(stage_pool_get(sizeof(double) * (needA + needB + 4), &pool));
    double* upA = static_cast<double*>(pool);
    double* upB = upA + needA + (needA & 1);
    cudaStream_t h2d = runtime().aux_stream;
    CANDMC_TRY(stream_wait(h2d, st));
    const double* hA = hostA ? mat_A + (ksplit ? layer * kloc * args->lda_A : 0) : nullptr;
    const double* hB = hostB ? mat_B + (ksplit ? layer * kloc : 0) : nullptr;
    // pinned B: gathered chunk by chunk straight out of host memory, lands chunk-major (even chunk depth: 16-byte accesses)
    const double* hB_dev = (hostB && up_chunks > 1 && (kloc / up_chunks) % 2 == 0) ? device_alias_of_pinned(hB) : nullptr;
    CANDMC_TRY(upload_chunks(hA, args->lda_A, hB, args->lda_B, b, b, kloc, up_chunks, upA, upB, h2d, &a_ready, &b_ready, hB_dev));
    if (hostA) { dA_ptr = upA; dA_ld = b; }
    else if (useA) { dA_ptr = mat_A + (ksplit ? layer * kloc * args->lda_A : 0); dA_ld = args->lda_A; }
    if (hostB && hB_dev) { dB_ptr = upB; dB_ld = kloc / up_chunks; b_chunk_major = true; host_gather_active = true; }
    else if (hostB) { dB_ptr = upB; dB_ld = kloc; }
    else if (useB) { dB_ptr = mat_B + (ksplit ? layer * kloc : 0); dB_ld = args->lda_B; }
  } else {
    CANDMC_TRY(sA.open(mat_A, b, b, args->lda_A, true, st));
    CANDMC_TRY(sB.open(mat_B, b, b, args->lda_B, true, st));
    // my k-slice of the stored operands: columns of A / rows of B, the other way round where an operand is transposed
    const int64_t offA = !ksplit ? 0 : (is_t(args->trans_A) ? layer * kloc : layer * kloc * sA.ld());
    const int64_t offB = !ksplit ? 0 : (is_t(args->trans_B) ? layer * kloc * sB.ld() : layer * kloc);
    if (useA) { dA_ptr = sA.ptr() + offA; dA_ld = sA.ld(); }
    if (useB) { dB_ptr = sB.ptr() + offB; dB_ld = sB.ld(); }
  }
  if (c > 1 && runtime().check_peer_args) {   // the depth-sum protocol below follows from the kind of mat_C (candmc_b200.h)
    bool agree = true;
    CANDMC_TRY(comm_flags_agree(cdt_kdir, is_device_ptr(mat_C) ? 0 : 1, &agree, runtime().comm_stream));
    CANDMC_CHECK(agree, "d25_summa: the ranks of a depth group must all pass host or all pass device memory for mat_C");
  }
  CANDMC_TRY(sC.open(mat_C, b, b, args->lda_C, false, st));
  void* wsv = nullptr;
  // packA | locB | bufA | bufB only when panels travel (q > 1); bufC only when there is a depth sum (c > 1)
  const int64_t ws_panels = (q > 1) ? 4 * b * b : 0;
  const int64_t ws_scratch = ksplit ? 2 * b * (kloc / (chunked ? up_chunks : 1)) + 8 : 0;  // TMA-alignment scratch
  CANDMC_TRY(workspace_get(sizeof(double) * (ws_panels + (c > 1 ? b * b : 0) + 2 + ws_scratch), &wsv));
  double* ws = static_cast<double*>(wsv);
  double* bufC = ws + ws_panels;
  // with replication the partial product goes to a contiguous scratch block and the depth sum writes mat_C
  double* Cpart = (c > 1) ? bufC : sC.ptr();
  const int64_t ldCpart = (c > 1) ? b : sC.ld();

  // depth all-reduce: fused into the epilogue of the last GEMM over peer memory when the block shape allows it
  // (whole 128-wide tile columns per depth rank, CUDA IPC available), otherwise ncclAllReduce
  // A HOST C block is the slowest thing this call moves (PCIe, after the last multiply): when the multiply is pipelined in
  // k-chunks anyway, the second half of the last panel's chunks is multiplied column slab by column slab, and every slab is
  // summed over the depth and downloaded while the next slabs still multiply — only the last
  // slab's transfer stays exposed.  The fused depth sum works on whole square blocks, so it stays off on this path.
  const int fin_slabs = (sC.staged() && nn && runtime().early_c_download) ? pick_fin_slabs(b) : 0;
  const int nch_consumer = ksplit ? (chunked ? up_chunks : 1) : sweep_chunks(b, args->trans_A, args->trans_B, cdt_row, cdt_col);
  const bool slab_mode = fin_slabs > 1 && nch_consumer >= 2 && !(c > 1 && !ksplit && runtime().fused_reduce_grids) &&
                         !fin_slab_widths(b, fin_slabs).empty();
  int slabs_out = 0;
  cudaStream_t d2h = runtime().aux_stream;
  auto slab_done = [&](int64_t c0, int64_t w) -> int {
    cudaEvent_t e = g_events.get();
    CANDMC_CHECK(e != nullptr, "event pool exhausted");
    CANDMC_CUDA(cudaEventRecord(e, st));
    if (c > 1) {   // depth sum of the slab in place in bufC (ld = b: the slab is contiguous), then straight to the host
      cudaStream_t cs = runtime().comm_stream;
      CANDMC_CUDA(cudaStreamWaitEvent(cs, e, 0));
      // full-width communicator: the one the single depth sum has always used (an all-reduce on a CTA-capped communicator has
      // never run on this NCCL; grouped send/recv on one hung).  It shares SMs with the slab multiplies.
      CANDMC_TRY(comm_allreduce(cdt_kdir, bufC + c0 * b, bufC + c0 * b, b * w, cs, false));
      cudaEvent_t r = g_events.get();
      CANDMC_CHECK(r != nullptr, "event pool exhausted");
      CANDMC_CUDA(cudaEventRecord(r, cs));
      CANDMC_CUDA(cudaStreamWaitEvent(d2h, r, 0));
      // The next slab's multiply waits for this slab's depth sum.  Left to itself the all-reduce kernel only becomes eligible
      // when the slab's GEMM has completed, by which time the next GEMM of the compute stream — eligible at the same instant,
      // without a cross-stream event in between — owns every SM again: on 8 B200s all four slab sums and with them the whole
      // download of C ended up BEHIND the last multiply (180 ms tail, profiles/r02_8gpu/r02_timeline8_e2e.txt).  The sum of a
      // slab over NVLink costs the multiplies a few milliseconds in total; the download it releases is what bounds the step.
      CANDMC_CUDA(cudaStreamWaitEvent(st, r, 0));
      CANDMC_TRY(sC.close_out_cols(bufC + c0 * b, b, c0, w, d2h));
    } else {
      CANDMC_CUDA(cudaStreamWaitEvent(d2h, e, 0));
      CANDMC_TRY(sC.close_out_cols(sC.ptr() + c0 * sC.ld(), sC.ld(), c0, w, d2h));
    }
    ++slabs_out;
    return OK;
  };

  // the kernels that gather B's k-chunks out of pinned host memory run beside the multiplies: leave them their SMs
  ReserveGuard gather_guard(host_gather_active ? std::max(2, runtime().gemm_reserve_sms) : runtime().gemm_reserve_sms);
  FusedCtx* fctx = nullptr;
  FusedParams fparams;
  // (on q > 1 grids the fused path is opt-in, candmc_set_fused_reduce(2): parity-green on 8 B200s but slower there than the
  // all-reduce it replaces — its P2P stores cost the last chunk's launch 6 ms, the all-reduce 4.6 ms)
  if (c > 1 && !slab_mode && (ksplit || runtime().fused_reduce_grids)) CANDMC_TRY(fused_ctx_get(cdt_kdir, b, &fctx));
  struct EpochScope {   // an error return below, before the fused launch is enqueued, takes the epoch back (ipc.h)
    FusedEpochGuard g;
    ~EpochScope() { g.settle(runtime().fused_launches); }
  } epoch_scope;
  if (fctx) {
    epoch_scope.g.arm(fctx, runtime().fused_launches);
    fused_params_next(fctx, layer, &fparams);
  }

  if (ksplit) {
    // my k-slice in launch groups of k-chunks as they land (one chunk when the operands are already on the device)
    const int nch = chunked ? up_chunks : 1;
    const int64_t kc = kloc / nch;
    auto operands = [&](int t, const double** pa, int64_t* lda, const double** pb, int64_t* ldb) -> int {
      if (t < (int)a_ready.size() && a_ready[t]) CANDMC_CUDA(cudaStreamWaitEvent(st, a_ready[t], 0));
      if (t < (int)b_ready.size() && b_ready[t]) CANDMC_CUDA(cudaStreamWaitEvent(st, b_ready[t], 0));
      *pa = dA_ptr + t * kc * dA_ld;
      *pb = b_chunk_major ? dB_ptr + t * kc * b : dB_ptr + t * kc;
      *lda = dA_ld;
      *ldb = b_chunk_major ? kc : dB_ld;
      return OK;
    };
    std::vector<int> grp_hi(nch);
    for (int t = 0; t < nch; ++t) grp_hi[t] = t + 1;
    if (runtime().merge_panels > 0 && nch > 2) host_upload_groups(nch, fctx ? nch - 1 : nch, &grp_hi);
    std::vector<int64_t> slab_w;
    if (slab_mode) slab_w = fin_slab_widths(b, fin_slabs);
    int last_group_lo = 0;
    for (int t = 0; t < nch; t = grp_hi[t]) last_group_lo = t;
    double* scratch = ws + ws_panels + (c > 1 ? b * b : 0) + 2;
    for (int t = 0; t < nch; t = grp_hi[t]) {
      const int mg = grp_hi[t] - t;
      std::vector<const double*> pa(mg), pb(mg);
      std::vector<int64_t> lda(mg), ldb(mg);
      for (int u = 0; u < mg; ++u) CANDMC_TRY(operands(t + u, &pa[u], &lda[u], &pb[u], &ldb[u]));
      const bool fused_here = fctx != nullptr && t + mg == nch;
      GroupLaunch gl;
      gl.tA = args->trans_A; gl.tB = args->trans_B; gl.b = b; gl.kc = kc; gl.C = Cpart; gl.ldC = ldCpart; gl.compute = st;
      gl.fused = fused_here ? &fparams : nullptr; gl.fused_out = sC.ptr(); gl.fused_ldout = sC.ld();
      gl.scratchA = scratch; gl.scratchB = scratch + b * kc + (b * kc & 1);
      CANDMC_TRY(multiply_group(gl, mg, pa, lda, pb, ldb, t ? 1.0 : 0.0, (!slab_w.empty() && t == last_group_lo) ? &slab_w : nullptr,
                                slab_done));
    }
  } else {
    SummaArgs a;
    a.tA = args->trans_A; a.tB = args->trans_B; a.b = b;
    a.i0 = layer * (q / c); a.i1 = (layer + 1) * (q / c);  // d25_summa.cxx:124,151
    a.myA = dA_ptr; a.ldA = dA_ld; a.myB = dB_ptr; a.ldB = dB_ld;
    a.C = Cpart; a.ldC = ldCpart; a.first_beta_zero = true;  // intended semantics, SURVEY App. A-1
    a.row = cdt_row; a.col = cdt_col; a.ws = ws; a.compute = st;
    if (chunked) {
      a.a_ready = &a_ready;
      a.b_ready = &b_ready;
      a.b_chunk_major = b_chunk_major;
    }
    if (fctx) {
      a.fused = &fparams;
      a.fused_out = sC.ptr();
      a.fused_ldout = sC.ld();
    }
    if (slab_mode) {
      a.fin_slabs = fin_slabs;
      a.slab_done = slab_done;
    }
    CANDMC_TRY(summa_sweep(a));
  }
  if (slabs_out > 0) {
    // every slab has been summed over the depth and is on its way to the host: nothing left but to wait for the transfers
    CANDMC_CHECK(slabs_out == (int)fin_slab_widths(b, fin_slabs).size(), "d25_summa: %d C slabs finalised, expected %d", slabs_out,
                 (int)fin_slab_widths(b, fin_slabs).size());
    CANDMC_TRY(stream_wait(st, d2h));
    CANDMC_TRY(stream_wait(st, runtime().comm_stream));
    CANDMC_CUDA(cudaStreamSynchronize(st));
    return OK;
  }
  if (fctx) {
    CANDMC_TRY(fused_finish(fctx, layer, fparams, sC.ptr(), sC.ld(), st));
  } else if (c > 1) {
    // MPI_Allreduce(buf_C, mat_C, b*b, SUM, cdt_kdir) — d25_summa.cxx:149,221 (result on every layer)
    if (sC.ld() == b) {
      CANDMC_TRY(comm_allreduce(cdt_kdir, bufC, sC.ptr(), b * b, st));
    } else {
      CANDMC_TRY(comm_allreduce(cdt_kdir, bufC, bufC, b * b, st));
      CANDMC_TRY(lda_copy_f64(b, b, b, sC.ld(), bufC, sC.ptr(), st));
    }
  }
  CANDMC_TRY(sC.close_out(st));
  if (chunked) CANDMC_TRY(stream_wait(st, runtime().aux_stream));
  if (chunked || sA.staged() || sB.staged() || sC.staged()) CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

// ================================================================================================================
int candmc_bcast_cannon_4d(const candmc_ctb_args_t* args, const double* mat_A, const double* mat_B, double* mat_C,
                           double* buffer, candmc_comm_t* cdt_x1, candmc_comm_t* cdt_y1, candmc_comm_t* cdt_x2,
                           candmc_comm_t* cdt_y2, void* stream) {
  NvtxRange nvtx_range("bcast_cannon_4d");   // dual_cannon.cxx:40 (the reference opens no timer here)
  CANDMC_TRY(runtime_require());
  g_events.reset();
  CANDMC_CHECK(args && cdt_x1 && cdt_y1 && cdt_x2 && cdt_y2, "bcast_cannon_4d: null argument");
  const int x1_np = cdt_x1->size, x2_np = cdt_x2->size;
  CANDMC_CHECK(x1_np == cdt_y1->size && x2_np == cdt_y2->size, "bcast_cannon_4d: grids must be square");  // :71-72
  CANDMC_CHECK(args->n > 0 && args->n % ((int64_t)x1_np * x2_np) == 0, "bcast_cannon_4d: n %% (x1_np*x2_np) != 0");
  // trans_A / trans_B: as in the reference (dual_cannon.cxx:163-166,188-194) the flags reach the local multiply only — the
  // square b x b blocks travel as they are stored and every block product is op(A block) * op(B block)
  CANDMC_CHECK((is_n(args->trans_A) || is_t(args->trans_A)) && (is_n(args->trans_B) || is_t(args->trans_B)),
               "bcast_cannon_4d: trans_A / trans_B must be 'N' or 'T'");
  const int64_t b = args->n / ((int64_t)x1_np * x2_np), bb = b * b;
  const int64_t need = (args->ovp ? 5 : 3) * bb * (int64_t)sizeof(double);  // dual_cannon.cxx:31-37
  CANDMC_CHECK(buffer == nullptr || args->buffer_size >= need, "bcast_cannon_4d: buffer_size too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStream_t shift = runtime().aux_stream;
  StagedMatrix sA, sB, sC;
  CANDMC_TRY(sA.open(mat_A, b, b, args->lda_A, true, st));
  CANDMC_TRY(sB.open(mat_B, b, b, args->lda_B, true, st));
  CANDMC_TRY(sC.open(mat_C, b, b, args->lda_C, false, st));
  void* wsv = nullptr;
  // 4 b^2 for the inner SUMMA sweep + two ping-pong pairs for the Cannon level
  CANDMC_TRY(workspace_get(sizeof(double) * 8 * bb, &wsv));
  double* ws = static_cast<double*>(wsv);
  double* pingA[2] = {ws + 4 * bb, ws + 5 * bb};
  double* pingB[2] = {ws + 6 * bb, ws + 7 * bb};

  const double* curA = sA.ptr();
  const double* curB = sB.ptr();
  int64_t ldA = sA.ld(), ldB = sB.ld();
  int nextA = 0, nextB = 0;  // ping-pong slot the next incoming block lands in (never the current block)
  const int x2 = cdt_x2->rank, y2 = cdt_y2->rank;
  auto wrap = [](int a, int m) { return ((a % m) + m) % m; };

  if (x2_np > 1) {
    // staggers and shifts by copy engines into peer windows (the default) instead of grouped ncclSend/ncclRecv: collective set-up
    CANDMC_TRY(p2p_transport_prepare(cdt_x2, bb));
    CANDMC_TRY(p2p_transport_prepare(cdt_y2, bb));
    CANDMC_TRY(stream_wait(shift, st));
    // messages must be contiguous: get the blocks out of their lda first (dual_cannon.cxx:89-102)
    if (ldA != b) {
      CANDMC_TRY(lda_copy_f64(b, b, ldA, b, curA, pingA[0], shift));
      curA = pingA[0]; ldA = b; nextA = 1;
    }
    if (ldB != b) {
      CANDMC_TRY(lda_copy_f64(b, b, ldB, b, curB, pingB[0], shift));
      curB = pingB[0]; ldB = b; nextB = 1;
    }
    // stagger (dual_cannon.cxx:106-137, with the tags/waits it meant): A to x2-y2 along cdt_x2, B to y2-x2 along cdt_y2
    const int tgtA = wrap(x2 - y2, x2_np), srcA = wrap(x2 + y2, x2_np);
    const int tgtB = wrap(y2 - x2, x2_np), srcB = wrap(y2 + x2, x2_np);
    if (tgtA != x2) {
      CANDMC_TRY(comm_sendrecv(cdt_x2, curA, bb, tgtA, pingA[nextA], bb, srcA, shift));
      curA = pingA[nextA]; nextA ^= 1;
    }
    if (tgtB != y2) {
      CANDMC_TRY(comm_sendrecv(cdt_y2, curB, bb, tgtB, pingB[nextB], bb, srcB, shift));
      curB = pingB[nextB]; nextB ^= 1;
    }
    CANDMC_TRY(stream_wait(st, shift));
  }

  // The shifts travel by copy engines into the neighbours' peer windows (p2p_transport_prepare above; default since round 2:
  // 90.6 -> 94.4 % at n = 24576 on 4 B200s, profiles/r02_4gpu_b/) and so run under the multiplies.  The NCCL fallback uses the
  // full-width communicators — grouped ncclSend/ncclRecv on the CTA-capped ones hung (profiles/r01_cannon_bg_hang.txt) — and
  // only gets SMs at GEMM boundaries.
  for (int i2 = 0; i2 < x2_np; ++i2) {
    const double* nxtA = curA;
    const double* nxtB = curB;
    cudaEvent_t shift_done = nullptr;
    if (i2 < x2_np - 1) {
      // shift by -1 (dual_cannon.cxx:196-213) into the alternate buffers WHILE this step multiplies.  The wait makes
      // sure the previous step's multiplies (the last readers of the alternate buffers) are finished.
      CANDMC_TRY(stream_wait(shift, st));
      CANDMC_TRY(comm_sendrecv(cdt_x2, curA, bb, wrap(x2 - 1, x2_np), pingA[nextA], bb, wrap(x2 + 1, x2_np), shift));
      CANDMC_TRY(comm_sendrecv(cdt_y2, curB, bb, wrap(y2 - 1, x2_np), pingB[nextB], bb, wrap(y2 + 1, x2_np), shift));
      nxtA = pingA[nextA]; nextA ^= 1;
      nxtB = pingB[nextB]; nextB ^= 1;
      shift_done = g_events.get();
      CANDMC_CHECK(shift_done != nullptr, "event pool exhausted");
      CANDMC_CUDA(cudaEventRecord(shift_done, shift));
    }
    SummaArgs a;
    a.tA = args->trans_A; a.tB = args->trans_B; a.b = b; a.i0 = 0; a.i1 = x1_np;
    a.myA = curA; a.ldA = ldA; a.myB = curB; a.ldB = ldB;
    a.C = sC.ptr(); a.ldC = sC.ld(); a.first_beta_zero = (i2 == 0);  // beta = (i1>0 || i2>0), dual_cannon.cxx:188
    a.row = cdt_x1; a.col = cdt_y1; a.ws = ws; a.compute = st;
    CANDMC_TRY(summa_sweep(a));
    if (shift_done) {
      CANDMC_CUDA(cudaStreamWaitEvent(st, shift_done, 0));
      curA = nxtA; curB = nxtB;
    }
  }
  CANDMC_TRY(sC.close_out(st));
  if (sA.staged() || sB.staged() || sC.staged()) CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

// ================================================================================================================
// split-dimensional Cannon.  Slices are contiguous ranges of the canonical A (m x k) and B^T (n x k) arrays, exactly
// as in the reference; every MPI_Put between two fences becomes a matched ncclSend/ncclRecv inside one group.
namespace {

struct Spc {
  candmc_comm* world;
  int rank, kary, ndim, half;
  int64_t n, m, k;
  double alpha;
  double* A[2];
  double* B[2];
  int cur = 0;
  double* C;
  int64_t ldc;
  cudaStream_t compute, comm;
  cudaEvent_t last_reader[2] = {nullptr, nullptr};  // latest GEMM that read buffer pair p
  cudaEvent_t last_xchg = nullptr;                  // latest exchange (produced the current pair)
};

int digit(int r, int kary, int pos) {
  for (int i = 0; i < pos; ++i) r /= kary;
  return r % kary;
}
int ipow(int a, int e) {
  int r = 1;
  while (e-- > 0) r *= a;
  return r;
}
int wrapi(int a, int m) { return ((a % m) + m) % m; }

struct Xfer {
  const double* send;
  double* recv;
  int64_t count;
  int dst, src;
};

int spc_exchange(Spc& s, const std::vector<Xfer>& xs, bool background) {
  const int p = s.cur;
  // destination pair 1-p must no longer be read by a GEMM; the source pair p must have been produced
  if (s.last_reader[1 - p]) CANDMC_CUDA(cudaStreamWaitEvent(s.comm, s.last_reader[1 - p], 0));
  bool any_remote = false;
  for (const Xfer& x : xs) {
    if (x.dst == s.rank) {
      CANDMC_CHECK(x.src == s.rank, "spcannon: asymmetric self transfer");
      CANDMC_CUDA(cudaMemcpyAsync(x.recv, x.send, sizeof(double) * x.count, cudaMemcpyDeviceToDevice, s.comm));
    } else {
      any_remote = true;
    }
  }
  int64_t largest = 0;
  for (const Xfer& x : xs) largest = std::max(largest, x.count);
  if (any_remote && p2p_transport_usable(s.world, largest)) {
    // every "put between two fences" of the reference as DMAs into the peers' windows: all sends first, then the receives
    for (const Xfer& x : xs)
      if (x.dst != s.rank) CANDMC_TRY(p2p_transport_send(s.world, x.send, x.count, x.dst, s.comm));
    for (const Xfer& x : xs)
      if (x.dst != s.rank) CANDMC_TRY(p2p_transport_recv(s.world, x.recv, x.count, x.src, s.comm));
  } else if (any_remote) {
    ncclComm_t comm = s.world->nccl;
    if (background) CANDMC_TRY(comm_background(s.world, &comm));  // shifts run under the GEMM; the stagger does not
    CANDMC_NCCL(ncclGroupStart());
    for (const Xfer& x : xs) {
      if (x.dst == s.rank) continue;
      CANDMC_NCCL(ncclSend(x.send, (size_t)x.count, ncclDouble, x.dst, comm, s.comm));
      CANDMC_NCCL(ncclRecv(x.recv, (size_t)x.count, ncclDouble, x.src, comm, s.comm));
    }
    CANDMC_NCCL(ncclGroupEnd());
  }
  s.last_xchg = g_events.get();
  CANDMC_CHECK(s.last_xchg != nullptr, "event pool exhausted");
  CANDMC_CUDA(cudaEventRecord(s.last_xchg, s.comm));
  s.cur = 1 - p;  // A = buf_A, B = buf_B (spcannon.cxx:76-77) as a pointer swap
  return OK;
}

int spc_stagger(Spc& s, int level) {  // uni_stagger, spcannon.cxx:33-84
  const int64_t bA = 2 * s.m * s.k / s.ndim, bB = 2 * s.k * s.n / s.ndim;
  std::vector<Xfer> xs;
  const int p = s.cur;
  for (int j = 0; j < s.half; ++j) {
    const int i = (j + level) % s.half;
    const int tA = digit(s.rank, s.kary, 2 * j), tB = digit(s.rank, s.kary, 2 * j + 1);
    const int sA = ipow(s.kary, 2 * j), sB = ipow(s.kary, 2 * j + 1);
    // I put to the rank whose digit is (tA - tB); the rank whose digit is (tA + tB) puts to me
    xs.push_back({s.A[p] + i * bA, s.A[1 - p] + i * bA, bA, s.rank + (wrapi(tA - tB, s.kary) - tA) * sA,
                  s.rank + (wrapi(tA + tB, s.kary) - tA) * sA});
    xs.push_back({s.B[p] + i * bB, s.B[1 - p] + i * bB, bB, s.rank + (wrapi(tB - tA, s.kary) - tB) * sB,
                  s.rank + (wrapi(tB + tA, s.kary) - tB) * sB});
  }
  CANDMC_TRY(spc_exchange(s, xs, false));
  if (level < s.half - 1) return spc_stagger(s, level + 1);
  return OK;
}

int spc_shift(Spc& s, int bidir, int level, double beta) {  // bdr_shift :87-162 / uni_shift :165-234
  double dbeta = beta;
  for (int ka = 0; ka < s.kary; ++ka) {
    if (level < s.half - 1) {
      CANDMC_TRY(spc_shift(s, bidir, level + 1, dbeta));
    } else {
      const int p = s.cur;
      if (s.last_xchg) CANDMC_CUDA(cudaStreamWaitEvent(s.compute, s.last_xchg, 0));
      // DGEMM('N','T',m,n,k,alpha,A,m,B,n,dbeta,C,m) — spcannon.cxx:117,195
      CANDMC_TRY(gemm_f64('N', 'T', s.m, s.n, s.k, s.alpha, s.A[p], s.m, s.B[p], s.n, dbeta, s.C, s.ldc, s.compute));
      s.last_reader[p] = g_events.get();
      CANDMC_CHECK(s.last_reader[p] != nullptr, "event pool exhausted");
      CANDMC_CUDA(cudaEventRecord(s.last_reader[p], s.compute));
    }
    dbeta = 1.0;
    if (s.kary == 1) continue;  // every shift is the identity
    // the shift after the very last multiply only restores the reference's in-place operands; A and B are private copies
    // here, so nobody would read it
    if (level == 0 && ka == s.kary - 1) continue;
    std::vector<Xfer> xs;
    const int p = s.cur;
    for (int j = 0; j < s.half; ++j) {
      const int i = (j + level) % s.half;
      const int tA = digit(s.rank, s.kary, 2 * j), tB = digit(s.rank, s.kary, 2 * j + 1);
      const int sA = ipow(s.kary, 2 * j), sB = ipow(s.kary, 2 * j + 1);
      const int upA = s.rank + (wrapi(tA + 1, s.kary) - tA) * sA, dnA = s.rank + (wrapi(tA - 1, s.kary) - tA) * sA;
      const int upB = s.rank + (wrapi(tB + 1, s.kary) - tB) * sB, dnB = s.rank + (wrapi(tB - 1, s.kary) - tB) * sB;
      if (bidir) {  // halves 2i and 2i+1 travel in opposite directions (spcannon.cxx:139-152)
        const int64_t bA = s.m * s.k / s.ndim, bB = s.k * s.n / s.ndim;
        xs.push_back({s.A[p] + 2 * i * bA, s.A[1 - p] + 2 * i * bA, bA, upA, dnA});
        xs.push_back({s.A[p] + (2 * i + 1) * bA, s.A[1 - p] + (2 * i + 1) * bA, bA, dnA, upA});
        xs.push_back({s.B[p] + 2 * i * bB, s.B[1 - p] + 2 * i * bB, bB, upB, dnB});
        xs.push_back({s.B[p] + (2 * i + 1) * bB, s.B[1 - p] + (2 * i + 1) * bB, bB, dnB, upB});
      } else {  // the whole slice travels +1 (spcannon.cxx:217-224)
        const int64_t bA = 2 * s.m * s.k / s.ndim, bB = 2 * s.k * s.n / s.ndim;
        xs.push_back({s.A[p] + i * bA, s.A[1 - p] + i * bA, bA, upA, dnA});
        xs.push_back({s.B[p] + i * bB, s.B[1 - p] + i * bB, bB, upB, dnB});
      }
    }
    CANDMC_TRY(spc_exchange(s, xs, false));  // full-width communicator, see the note in candmc_bcast_cannon_4d
  }
  return OK;
}

}  // namespace

int candmc_spcannon(int bidir, int rank, int kary, int ndim, candmc_comm_t* world, int n, int m, int k,
                    char transp_A, double alpha, const double* A, char transp_B, double beta, const double* B,
                    double* C, void* stream) {
  CANDMC_TRY(runtime_require());
  g_events.reset();
  CANDMC_CHECK(world != nullptr, "spcannon: null communicator");
  CANDMC_CHECK(ndim >= 2 && ndim % 2 == 0 && kary >= 1, "spcannon: need an even ndim >= 2 and kary >= 1");
  CANDMC_CHECK(k % ndim == 0, "spcannon: k %% ndim != 0");  // assert(k%ndim == 0), spcannon.cxx:252
  CANDMC_CHECK(ipow(kary, ndim) == world->size, "spcannon: kary^ndim = %d but communicator has %d ranks",
               ipow(kary, ndim), world->size);
  CANDMC_CHECK(rank == world->rank, "spcannon: rank argument %d != communicator rank %d", rank, world->rank);
  CANDMC_CHECK(n >= 0 && m >= 0 && k >= 0, "spcannon: negative dimension");
  CANDMC_CHECK((is_t(transp_A) || is_n(transp_A)) && (is_t(transp_B) || is_n(transp_B)), "spcannon: bad transpose");
  if (m == 0 || n == 0) return OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tA = is_t(transp_A), tB = is_t(transp_B);
  if (k == 0) return candmc_dgemm('N', 'N', m, n, 0, alpha, A, m, B, 1, beta, C, m, stream);  // C <- beta*C
  StagedMatrix sA, sB, sC;
  CANDMC_TRY(sA.open(A, tA ? k : m, tA ? m : k, tA ? k : m, true, st));
  CANDMC_TRY(sB.open(B, tB ? n : k, tB ? k : n, tB ? n : k, true, st));
  CANDMC_TRY(sC.open(C, m, n, m, beta != 0.0, st));
  const int64_t mk = (int64_t)m * k, nk = (int64_t)n * k;
  void* wsv = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * 2 * (mk + nk + 2), &wsv));
  double* ws = static_cast<double*>(wsv);
  Spc s;
  s.world = world; s.rank = rank; s.kary = kary; s.ndim = ndim; s.half = ndim / 2;
  s.n = n; s.m = m; s.k = k; s.alpha = alpha;
  s.A[0] = ws; s.A[1] = ws + mk + (mk & 1);
  s.B[0] = s.A[1] + mk + (mk & 1); s.B[1] = s.B[0] + nk + (nk & 1);
  s.C = sC.ptr(); s.ldc = sC.ld(); s.compute = st; s.comm = runtime().comm_stream;
  // canonicalise (spcannon.cxx:262-267): A -> m x k, B -> B^T = n x k; out of place, so the caller's A/B survive
  if (k > 0) {
    if (tA) CANDMC_TRY(transpose_f64(k, m, sA.ptr(), sA.ld(), s.A[0], m, st));
    else CANDMC_TRY(lda_copy_f64(m, k, sA.ld(), m, sA.ptr(), s.A[0], st));
    if (!tB) CANDMC_TRY(transpose_f64(k, n, sB.ptr(), sB.ld(), s.B[0], n, st));
    else CANDMC_TRY(lda_copy_f64(n, k, sB.ld(), n, sB.ptr(), s.B[0], st));
  }
  // every put of the stagger and the shifts by copy engines into peer windows (the default); the largest message is a whole slice
  if (kary > 1 && k > 0) CANDMC_TRY(p2p_transport_prepare(world, 2 * std::max(mk, nk) / ndim + 2));
  CANDMC_TRY(stream_wait(s.comm, st));
  if (kary > 1 && k > 0) {
    NvtxRange nvtx_range("uni_stagger");   // spcannon.cxx:278,334
    CANDMC_TRY(spc_stagger(s, 0));
  }
  {
    NvtxRange nvtx_range(bidir ? "bdr_shift" : "uni_shift");   // spcannon.cxx:285,341
    CANDMC_TRY(spc_shift(s, bidir, 0, beta));
  }
  CANDMC_TRY(stream_wait(st, s.comm));
  CANDMC_TRY(sC.close_out(st));
  if (sA.staged() || sB.staged() || sC.staged()) CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

// ================================================================================================================
}  // extern "C"

namespace candmc {
namespace {

// Ybuf = Y panel out of its lda; on the root grid row the upper triangle is zeroed and the diagonal set to one
// (copy_lower(..., zero_square = 1) + the explicit 1.0 of update_A, alg/QR/qr_2d/qr_2d.cxx:157-163)
__global__ void pack_y_panel_kernel(const double* __restrict__ Y, int64_t lda_Y, double* __restrict__ Ybuf, int64_t mb,
                                    int64_t b, int root_row) {
  const int64_t total = mb * b;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t j = e / mb, r = e - j * mb;
    double v = Y[r + j * lda_Y];
    if (root_row) {
      if (r < j) v = 0.0;
      if (r == j) v = 1.0;
    }
    Ybuf[e] = v;
  }
}

// T = lower triangle of S with halved diagonal, zero above (compute_invT_from_Y, qr_2d.cxx:36-50; T zero-filled :235)
__global__ void tril_halve_diag_kernel(const double* __restrict__ S, double* __restrict__ T, int64_t b) {
  const int64_t total = b * b;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t j = e / b, i = e - j * b;
    T[e] = (i > j) ? S[e] : (i == j ? 0.5 * S[e] : 0.0);
  }
}

// comp_bcast_T_from_W on the root rank (qr_2d.cxx:189-195): L = W^T as a lower-triangular b x b matrix (W is the panel QR's
// upper-triangular factor, its strict lower triangle is never read), X0 = -Y1 (top b x b of the packed panel); the
// triangular solve L X = X0 that follows is compute_invT_from_W's cdtrsm('L','U','T','N', alpha = -1) (hh_recon.cxx:26-31)
__global__ void t_from_w_setup_kernel(const double* __restrict__ W, const double* __restrict__ Ybuf, int64_t ldy,
                                      double* __restrict__ L, double* __restrict__ X, int64_t b) {
  const int64_t total = b * b;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t j = e / b, i = e - j * b;
    L[e] = (i >= j) ? W[j + i * b] : 0.0;
    X[e] = (i >= j) ? -Ybuf[i + j * ldy] : 0.0;   // Y1 is unit lower-triangular (copy_lower + 1.0, :157-163)
  }
}

int upd_A_impl(const double* Y, int64_t lda_Y, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
               const double* T, candmc_comm* ccol, double* W, cudaStream_t st) {
  if (kb == 0) return OK;
  // W = Y^T A (qr_2d.cxx:259); ranks without rows contribute zeros (:262)
  if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, kb, mb, 1.0, Y, lda_Y, A, lda_A, 0.0, W, b, st));
  else CANDMC_TRY(fill_f64(W, b * kb, 0.0, st));
  if (ccol != nullptr && ccol->size > 1) CANDMC_TRY(comm_allreduce(ccol, W, W, b * kb, st));  // :265
  if (mb > 0) {
    CANDMC_TRY(trsm_llnn(b, kb, T, b, W, b, st));                                              // :271
    CANDMC_TRY(gemm_f64('N', 'N', mb, kb, b, -1.0, Y, lda_Y, W, b, 1.0, A, lda_A, st));        // :275
  }
  return OK;
}

// A <- A + Qm * (T * (Qm^T A)) on one grid column: the Yamamoto form keeps T = (Q1 - S)^-1 explicitly, so the middle step is
// a GEMM with alpha = -1 instead of a triangular solve (qr_y2d.cxx:123-169).  W, W2: b x kb scratch each.
int upd_Yamamoto_A_impl(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                        const double* T, candmc_comm* ccol, double* W, double* W2, cudaStream_t st) {
  if (kb == 0) return OK;
  if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, kb, mb, 1.0, Qm, lda_Qm, A, lda_A, 0.0, W, b, st));  // :140
  else CANDMC_TRY(fill_f64(W, b * kb, 0.0, st));                                                      // :143
  if (ccol != nullptr && ccol->size > 1) CANDMC_TRY(comm_allreduce(ccol, W, W, b * kb, st));          // :146
  if (mb > 0) {
    CANDMC_TRY(gemm_f64('N', 'N', b, kb, b, -1.0, T, b, W, b, 0.0, W2, b, st));                       // :156
    CANDMC_TRY(gemm_f64('N', 'N', mb, kb, b, -1.0, Qm, lda_Qm, W2, b, 1.0, A, lda_A, st));            // :160
  }
  return OK;
}

// aggregator::append (alg/QR/qr_2d/qr_y2d.cxx:38-62) on the device: the broadcast panel Qp (mb x b, ld = mb) goes into
// aQm at column n, row `shift`; the aggregated T grows by one block row:
//   aT[n.., 0..n] = T * ((Qp^T aQm[shift.., 0..n], summed over the grid column) * aT[0..n, 0..n]),   aT[n.., n..] = T
// (the reference's comment says -T11 (Y^T Y) T22; its code, which is followed here, has no minus sign).  A rank with no rows
// of the panel contributes zeros to the sum — the reference clears only b*b of the b*n doubles it then all-reduces (:54-56).
int aggregator_append(candmc_aggregator_t* agg, int64_t mb, int64_t b, const double* Qp, const double* T, candmc_comm* ccol,
                      cudaStream_t st) {
  CANDMC_CHECK(agg != nullptr && agg->aQm != nullptr && agg->aT != nullptr && agg->scratch != nullptr, "aggregator: not created");
  const int64_t n = agg->n;
  CANDMC_CHECK(n + b <= agg->lda_aT && agg->shift + mb <= agg->lda_aQm, "aggregator: panel does not fit (n = %lld, b = %lld, lda_aT = %lld; "
               "shift = %lld, mb = %lld, lda_aQm = %lld)", (long long)n, (long long)b, (long long)agg->lda_aT, (long long)agg->shift,
               (long long)mb, (long long)agg->lda_aQm);
  if (mb > 0) CANDMC_TRY(lda_copy_f64(mb, b, mb, agg->lda_aQm, Qp, agg->aQm + n * agg->lda_aQm + agg->shift, st));
  if (n == 0) {
    CANDMC_TRY(lda_copy_f64(b, b, b, agg->lda_aT, T, agg->aT, st));
  } else {
    double* t1 = agg->scratch;          // b x n
    double* t2 = t1 + b * n + (b * n & 1);
    if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, n, mb, 1.0, Qp, mb, agg->aQm + agg->shift, agg->lda_aQm, 0.0, t1, b, st));
    else CANDMC_TRY(fill_f64(t1, b * n, 0.0, st));
    if (ccol != nullptr && ccol->size > 1) CANDMC_TRY(comm_allreduce(ccol, t1, t1, b * n, st));
    CANDMC_TRY(gemm_f64('N', 'N', b, n, n, 1.0, t1, b, agg->aT, agg->lda_aT, 0.0, t2, b, st));
    CANDMC_TRY(gemm_f64('N', 'N', b, n, b, 1.0, T, b, t2, b, 0.0, agg->aT + n, agg->lda_aT, st));
    CANDMC_TRY(lda_copy_f64(b, b, b, agg->lda_aT, T, agg->aT + n * agg->lda_aT + n, st));
  }
  agg->n += b;
  return OK;
}

int update_Yamamoto_A_core(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                           double* T, const candmc_pview_t* pv, candmc_aggregator_t* agg, bool update, void* stream) {
  NvtxRange nvtx_range("update_Yamamoto_A");
  CANDMC_TRY(runtime_require());
  g_events.reset();
  CANDMC_CHECK(pv != nullptr && pv->crow != nullptr && pv->ccol != nullptr, "update_Yamamoto_A: null processor view");
  CANDMC_CHECK(b > 0 && m >= 0 && k >= 0 && m % b == 0 && k % b == 0, "update_Yamamoto_A: m and k must be multiples of b");
  CANDMC_CHECK(T != nullptr, "update_Yamamoto_A: null T");
  CANDMC_CHECK(update || agg != nullptr, "update_Yamamoto_A: nothing to do");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nprow = pv->ccol->size, npcol = pv->crow->size, myrow = pv->ccol->rank, mycol = pv->crow->rank;
  CANDMC_CHECK(pv->rrow >= 0 && pv->rrow < nprow && pv->rcol >= 0 && pv->rcol < npcol, "update_Yamamoto_A: bad root row/col");
  // block-cyclic local extents of the remaining matrix, qr_y2d.cxx:81-88 (same formulas as update_A)
  int64_t mb = (m / b) / nprow;
  if ((myrow + nprow - pv->rrow) % nprow < (m / b) % nprow) mb++;
  mb *= b;
  int64_t kb = (k / b) / npcol;
  if ((mycol + npcol - pv->rcol - 1) % npcol < (k / b) % npcol) kb++;
  kb *= b;
  CANDMC_CHECK(mb == 0 || mycol != pv->rcol || (Qm != nullptr && lda_Qm >= mb),
               "update_Yamamoto_A: the root column needs its Qm panel (lda_Qm >= %lld)", (long long)mb);
  if (!update) kb = 0;   // the last panel of a block column (QR_Yamamoto_2D :266-271): broadcast and append only
  // (a rank whose share of the trailing matrix is empty may hold a pointer past the end of its array, as the reference's
  // drivers do after their pointer arithmetic: it is never dereferenced)
  // Host operands — what QR_Yamamoto_2D itself holds (qr_y2d.cxx:214) — are staged for the call: the panel and T on the root
  // column, A where the rank has a share; A and T (an output on the other columns, :112) are written back before it returns.
  const bool host_call = !is_device_ptr(Qm) || !is_device_ptr(A) || !is_device_ptr(T);
  StagedMatrix sQ, sAm, sT;
  CANDMC_TRY(sQ.open((mycol == pv->rcol && mb > 0) ? Qm : nullptr, mb, b, lda_Qm, true, st));
  CANDMC_TRY(sAm.open((mb > 0 && kb > 0) ? A : nullptr, mb, kb, lda_A, true, st));
  CANDMC_TRY(sT.open(T, b, b, b, mycol == pv->rcol, st));
  Qm = sQ.ptr(); lda_Qm = sQ.ld();
  if (mb > 0 && kb > 0) { A = sAm.ptr(); lda_A = sAm.ld(); }
  T = sT.ptr();
  void* wsv = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * (mb * b + 2 * b * kb + 8), &wsv));
  double* Qbuf = static_cast<double*>(wsv);
  double* W = Qbuf + mb * b + (mb * b & 1);
  double* W2 = W + b * kb;
  // packed panel on the root column (:101-106), MPI_Bcast of the panel and of T along the grid row (:109-113)
  if (mycol == pv->rcol && mb > 0) CANDMC_TRY(lda_copy_f64(mb, b, lda_Qm, mb, Qm, Qbuf, st));
  if (mb > 0) CANDMC_TRY(comm_bcast(pv->crow, Qbuf, Qbuf, mb * b, pv->rcol, st));
  CANDMC_TRY(comm_bcast(pv->crow, T, T, b * b, pv->rcol, st));
  if (update) CANDMC_TRY(upd_Yamamoto_A_impl(Qbuf, mb, A, lda_A, mb, kb, b, T, pv->ccol, W, W2, st));
  if (agg != nullptr) CANDMC_TRY(aggregator_append(agg, mb, b, Qbuf, T, pv->ccol, st));   // :117-118
  if (host_call) {
    CANDMC_TRY(sAm.close_out(st));
    CANDMC_TRY(sT.close_out(st));
    CANDMC_CUDA(cudaStreamSynchronize(st));
  }
  return OK;
}

}  // namespace
}  // namespace candmc

extern "C" {

int candmc_update_Yamamoto_A(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                             double* T, const candmc_pview_t* pv, void* stream) {
  return update_Yamamoto_A_core(Qm, lda_Qm, A, lda_A, m, k, b, T, pv, nullptr, true, stream);
}

int candmc_update_Yamamoto_A_agg(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                                 double* T, const candmc_pview_t* pv, candmc_aggregator_t* agg, int update, void* stream) {
  return update_Yamamoto_A_core(Qm, lda_Qm, A, lda_A, m, k, b, T, pv, agg, update != 0, stream);
}

int candmc_aggregator_create(int64_t lda_aQm, int64_t lda_aT, candmc_aggregator_t* out) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(out != nullptr && lda_aQm > 0 && lda_aT > 0, "aggregator: bad sizes");
  memset(out, 0, sizeof(*out));
  out->lda_aQm = lda_aQm;
  out->lda_aT = lda_aT;
  // aQm: lda_aQm x lda_aT, aT: lda_aT x lda_aT (aggregator::aggregator, qr_y2d.cxx:13-23); scratch: two b x n blocks
  CANDMC_CUDA(cudaMalloc(&out->aQm, sizeof(double) * lda_aQm * lda_aT));
  CANDMC_CUDA(cudaMalloc(&out->aT, sizeof(double) * lda_aT * lda_aT));
  CANDMC_CUDA(cudaMalloc(&out->scratch, sizeof(double) * (2 * lda_aT * lda_aT + 2)));
  return candmc_aggregator_reset(out);
}

int candmc_aggregator_reset(candmc_aggregator_t* agg) {   // aggregator::reset, qr_y2d.cxx:25-31
  CANDMC_CHECK(agg != nullptr && agg->aQm != nullptr && agg->aT != nullptr, "aggregator: not created");
  agg->n = 0;
  agg->shift = 0;
  CANDMC_CUDA(cudaMemset(agg->aQm, 0, sizeof(double) * agg->lda_aQm * agg->lda_aT));
  CANDMC_CUDA(cudaMemset(agg->aT, 0, sizeof(double) * agg->lda_aT * agg->lda_aT));
  return OK;
}

int candmc_aggregator_shift_down(candmc_aggregator_t* agg, int64_t b) {   // aggregator::shift_down, qr_y2d.cxx:34-36
  CANDMC_CHECK(agg != nullptr && b >= 0, "aggregator: bad shift");
  agg->shift += b;
  return OK;
}

int candmc_aggregator_free(candmc_aggregator_t* agg) {
  if (agg == nullptr) return OK;
  if (agg->aQm) cudaFree(agg->aQm);
  if (agg->aT) cudaFree(agg->aT);
  if (agg->scratch) cudaFree(agg->scratch);
  memset(agg, 0, sizeof(*agg));
  return OK;
}

int candmc_upd_Yamamoto_A(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                          const double* T, candmc_comm_t* ccol, void* stream) {
  NvtxRange nvtx_range("upd_Yamamoto_A");
  CANDMC_TRY(runtime_require());
  g_events.reset();
  CANDMC_CHECK(mb >= 0 && kb >= 0 && b > 0, "upd_Yamamoto_A: bad extents");
  CANDMC_CHECK((Qm != nullptr && A != nullptr) || mb == 0 || kb == 0, "upd_Yamamoto_A: null operand");
  CANDMC_CHECK(T != nullptr || mb == 0 || kb == 0, "upd_Yamamoto_A: null T");
  if (kb == 0) return OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // host operands (the reference's QR_Yamamoto drivers, qr_y2d.cxx:113,367) are staged for the call, as in candmc_upd_A
  StagedMatrix sQ, sA, sT;
  CANDMC_TRY(sQ.open(Qm, mb, b, lda_Qm, true, st));
  CANDMC_TRY(sA.open(A, mb, kb, lda_A, true, st));
  CANDMC_TRY(sT.open(mb > 0 ? T : nullptr, b, b, b, true, st));   // (a rank without rows never reads T, :150)
  void* wsv = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * 2 * b * kb, &wsv));
  double* W = static_cast<double*>(wsv);
  CANDMC_TRY(upd_Yamamoto_A_impl(sQ.ptr(), sQ.ld(), sA.ptr(), sA.ld(), mb, kb, b, sT.ptr(), ccol, W, W + b * kb, st));
  if (mb > 0) CANDMC_TRY(sA.close_out(st));
  if (sQ.staged() || sA.staged() || sT.staged() || !is_device_ptr(Qm) || !is_device_ptr(A)) CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int candmc_update_A(const double* Y, int64_t lda_Y, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                    const double* W, const candmc_pview_t* pv, double* aggreg_Y, int64_t lda_aY, int W_is_T,
                    void* stream) {
  NvtxRange nvtx_range("Bcast_update");
  CANDMC_TRY(runtime_require());
  g_events.reset();
  CANDMC_CHECK(pv != nullptr && pv->crow != nullptr && pv->ccol != nullptr, "update_A: null processor view");
  CANDMC_CHECK(b > 0 && m >= 0 && k >= 0 && m % b == 0 && k % b == 0, "update_A: m and k must be multiples of b");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nprow = pv->ccol->size, npcol = pv->crow->size, myrow = pv->ccol->rank, mycol = pv->crow->rank;
  CANDMC_CHECK(pv->rrow >= 0 && pv->rrow < nprow && pv->rcol >= 0 && pv->rcol < npcol, "update_A: bad root row/col");
  // three forms, chosen by the same (W == NULL, W_is_T) on every rank: T from Y, W is T, or W is the panel QR's factor, which
  // only the root rank reads (upd_A :244-253) — the other ranks may hand in any non-null pointer, as QR_2D does (:309,325)
  const bool t_from_w = (W != nullptr && !W_is_T);
  const bool w_root = t_from_w && myrow == pv->rrow && mycol == pv->rcol;
  CANDMC_CHECK(!t_from_w || m >= b, "update_A: the panel factor form needs at least b rows (m=%lld, b=%lld)", (long long)m,
               (long long)b);
  // block-cyclic local extents of the remaining matrix, qr_2d.cxx:140-147
  int64_t mb = (m / b) / nprow;
  if ((myrow + nprow - pv->rrow) % nprow < (m / b) % nprow) mb++;
  mb *= b;
  int64_t kb = (k / b) / npcol;
  if ((mycol + npcol - pv->rcol - 1) % npcol < (k / b) % npcol) kb++;
  kb *= b;
  // Host operands — what QR_2D itself holds (qr_2d.cxx:325) — are staged for the call like the multiplies' blocks: the panel on
  // the root column only, W where it is read (every rank for W_is_T, the root rank for the panel factor), A and aggreg_Y where the
  // rank has rows; A and aggreg_Y are written back before the call returns.  Device pointers are used in place, asynchronously.
  StagedMatrix sYp, sAm, sWm, sAgg;
  CANDMC_TRY(sYp.open((mycol == pv->rcol && mb > 0) ? Y : nullptr, mb, b, lda_Y, true, st));
  CANDMC_TRY(sAm.open((mb > 0 && kb > 0) ? A : nullptr, mb, kb, lda_A, true, st));
  CANDMC_TRY(sWm.open((W != nullptr && (W_is_T ? mb > 0 : w_root)) ? W : nullptr, b, b, b, true, st));
  CANDMC_TRY(sAgg.open(mb > 0 ? aggreg_Y : nullptr, mb, b, lda_aY, false, st));
  // (a rank without rows stages nothing; it still returns only when its part of the collectives has run, like its peers)
  const bool any_staged = sYp.staged() || sAm.staged() || sWm.staged() || sAgg.staged() || !is_device_ptr(A) || !is_device_ptr(Y);
  Y = sYp.ptr(); lda_Y = sYp.ld();
  if (mb > 0 && kb > 0) { A = sAm.ptr(); lda_A = sAm.ld(); }
  if (sWm.staged()) W = sWm.ptr();
  if (mb > 0 && aggreg_Y != nullptr) { aggreg_Y = sAgg.ptr(); lda_aY = sAgg.ld(); }
  void* wsv = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * (mb * b + 2 * b * b + b * kb + 8), &wsv));
  double* Ybuf = static_cast<double*>(wsv);
  double* S = Ybuf + mb * b + (mb * b & 1);
  double* T = S + b * b;
  double* Wbuf = T + b * b;
  // Ybuf on the root column (:155-165), then MPI_Bcast along the grid row (:168)
  if (mycol == pv->rcol && mb > 0) {
    int64_t g = (mb * b + 255) / 256;
    const int64_t cap = static_cast<int64_t>(runtime().num_sms) * 8;
    if (g > cap) g = cap;
    pack_y_panel_kernel<<<static_cast<int>(g), 256, 0, st>>>(Y, lda_Y, Ybuf, mb, b, myrow == pv->rrow);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
  }
  if (mb > 0) CANDMC_TRY(comm_bcast(pv->crow, Ybuf, Ybuf, mb * b, pv->rcol, st));
  const double* Tuse = W;
  if (W == nullptr) {
    // T^-1 form from Y: lower triangle of sum over the grid column of Ybuf^T Ybuf, diagonal halved (:22-60)
    if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, b, mb, 1.0, Ybuf, mb, Ybuf, mb, 0.0, S, b, st));
    else CANDMC_TRY(fill_f64(S, b * b, 0.0, st));
    if (nprow > 1) CANDMC_TRY(comm_allreduce(pv->ccol, S, S, b * b, st));
    tril_halve_diag_kernel<<<static_cast<int>((b * b + 255) / 256), 256, 0, st>>>(S, T, b);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
    Tuse = T;
  } else if (t_from_w) {
    // comp_bcast_T_from_W (:179-208): the root rank solves W^T X = -Y1 and every rank receives X's lower triangle.  The
    // reference broadcasts over cworld from rank rcol + rrow*npcol (:250), which is the root only when that numbering matches
    // the grid's (its own drivers number rank = myrow + mycol*nprow); here T travels along the root's grid row and then down
    // every grid column, which needs no assumption about the world numbering.
    if (w_root) {
      t_from_w_setup_kernel<<<static_cast<int>((b * b + 255) / 256), 256, 0, st>>>(W, Ybuf, mb, S, T, b);
      CANDMC_CUDA(cudaGetLastError());
      runtime().launches++;
      CANDMC_TRY(trsm_llnn(b, b, S, b, T, b, st));
    }
    if (myrow == pv->rrow && npcol > 1) CANDMC_TRY(comm_bcast(pv->crow, T, T, b * b, pv->rcol, st));
    if (nprow > 1) CANDMC_TRY(comm_bcast(pv->ccol, T, T, b * b, pv->rrow, st));
    Tuse = T;
  }
  CANDMC_TRY(upd_A_impl(Ybuf, mb, A, lda_A, mb, kb, b, Tuse, pv->ccol, Wbuf, st));
  if (aggreg_Y != nullptr && mb > 0) CANDMC_TRY(lda_copy_f64(mb, b, mb, lda_aY, Ybuf, aggreg_Y, st));  // :172-174
  if (any_staged) {
    CANDMC_TRY(sAm.close_out(st));
    CANDMC_TRY(sAgg.close_out(st));
    CANDMC_CUDA(cudaStreamSynchronize(st));
  }
  return OK;
}

int candmc_upd_A(const double* Y, int64_t lda_Y, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                 const double* T, candmc_comm_t* ccol, void* stream) {
  NvtxRange nvtx_range("upd_A");
  CANDMC_TRY(runtime_require());
  g_events.reset();
  CANDMC_CHECK(mb >= 0 && kb >= 0 && b > 0, "upd_A: bad extents");
  CANDMC_CHECK(Y != nullptr || mb == 0, "upd_A: null Y");
  CANDMC_CHECK(A != nullptr || mb == 0 || kb == 0, "upd_A: null A");
  if (kb == 0) return OK;   // (the reference forms T before it looks at kb; nothing reads it then)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // host operands (what the reference's QR drivers own, qr_2d.cxx:447-620) are staged for the call, like the multiplies' blocks
  StagedMatrix sY, sA, sT;
  CANDMC_TRY(sY.open(Y, mb, b, lda_Y, true, st));
  CANDMC_TRY(sA.open(A, mb, kb, lda_A, true, st));
  CANDMC_TRY(sT.open(mb > 0 ? T : nullptr, b, b, b, true, st));   // (a rank without rows never reads T, :268)
  void* wsv = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * (b * kb + 2 * b * b), &wsv));
  double* W = static_cast<double*>(wsv);
  const double* Tuse = sT.ptr();
  if (T == nullptr) {
    // W == NULL in the reference (qr_2d.cxx:241-246): T^-1 from Y — lower triangle of the grid column's sum of Y^T Y, diagonal
    // halved (compute_invT_from_Y :22-60; there the root column computes and broadcasts along the row — every column holds
    // the same Ybuf rows, so here each computes its own copy and the broadcast disappears)
    double* S = W + b * kb;
    double* Tl = S + b * b;
    if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, b, mb, 1.0, sY.ptr(), sY.ld(), sY.ptr(), sY.ld(), 0.0, S, b, st));
    else CANDMC_TRY(fill_f64(S, b * b, 0.0, st));
    if (ccol != nullptr && ccol->size > 1) CANDMC_TRY(comm_allreduce(ccol, S, S, b * b, st));
    tril_halve_diag_kernel<<<static_cast<int>((b * b + 255) / 256), 256, 0, st>>>(S, Tl, b);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
    Tuse = Tl;
  }
  CANDMC_TRY(upd_A_impl(sY.ptr(), sY.ld(), sA.ptr(), sA.ld(), mb, kb, b, Tuse, ccol, W, st));
  if (mb > 0) CANDMC_TRY(sA.close_out(st));
  if (sY.staged() || sA.staged() || sT.staged() || !is_device_ptr(Y) || !is_device_ptr(A)) CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

}  // extern "C"
