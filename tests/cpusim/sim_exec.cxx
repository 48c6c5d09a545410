// tests/cpusim — kernel executor of the CPU functional simulator (TEST INFRASTRUCTURE ONLY, see sim_device.h).
// One block at a time; the threads of a block are ucontext fibers scheduled round-robin; a fiber leaves the CPU only at
// a barrier (__syncthreads, warp shuffle) or when its kernel body returns.
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <ucontext.h>

// Fiber switch.  glibc's swapcontext saves the signal mask with a system call on every switch; a kernel emulation makes
// millions of switches, so x86-64 gets the classic callee-saved-registers-and-stack-pointer switch instead.
#if defined(__x86_64__)
#define CPUSIM_FAST_SWITCH 1
extern "C" void cpusim_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl cpusim_ctx_switch
.type cpusim_ctx_switch,@function
cpusim_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cpusim_ctx_switch,.-cpusim_ctx_switch
)");
#endif

#include <map>
#include <vector>

#include "sim_device.h"
#include "sim_internal.h"

namespace cpusim {

namespace {

constexpr size_t kStack = 256 * 1024;
constexpr int kMaxThreads = 1024;

constexpr int kDmmaBatch = 64;
struct PendingDmma {
  double *d0, *d1;
};
struct Fiber {
#ifdef CPUSIM_FAST_SWITCH
  void* sp = nullptr;
#else
  ucontext_t ctx;
#endif
  ThreadState st;
  bool started = false, done = false;
  int npending = 0;                 // DMMAs logged since the last warp rendezvous
  PendingDmma pending[kDmmaBatch];
};

struct Warp {
  int alive = 0, arrived = 0;
  unsigned gen = 0;
  alignas(16) unsigned char slot[2][32][16];
  double dmma_ab[kDmmaBatch][32][2];   // operands (a, b) of every lane for each logged DMMA of the current batch
};

struct MBar {
  uint32_t count = 0, pending = 0, phase = 0;
  int64_t tx = 0;
};
struct NamedBar {
  int arrived = 0;
  unsigned gen = 0;
};

struct Block {
  int nthreads = 0, alive = 0, arrived = 0;
  unsigned gen = 0;
  std::vector<Warp> warps;
  std::map<uint32_t, MBar> mbars;
  NamedBar named[16];
  // tcgen05.mma is asynchronous: issued operations are queued and executed as LATE as the program allows — when somebody waits
  // on an mbarrier that a tcgen05.commit behind them will arrive on (or when the block ends).  A kernel that reads its
  // accumulator without waiting for the commit sees stale tensor memory, one that refills a shared-memory stage without
  // waiting for the commit has its operands overwritten before the multiply reads them: both show up as wrong results.
  struct PendingUmma {
    uint32_t tmem_d, idesc, accumulate;
    uint64_t desc_a, desc_b;
  };
  std::vector<PendingUmma> umma_queue;
  size_t umma_done = 0;                                    // operations [0, umma_done) have been executed
  std::vector<std::pair<uint32_t, size_t>> umma_commits;   // (mbarrier, number of operations issued before the commit)
  size_t commits_done = 0;
  std::vector<uint32_t> tmem;      // tensor memory of the SM the block runs on: 128 lanes x 512 columns of 32 bits
  uint32_t tmem_next = 0;          // bump allocator (columns)
  int tmem_live = 0;               // allocations not yet returned
};
constexpr uint32_t kTmemLanes = 128, kTmemCols = 512;
uint64_t g_progress = 0;   // anything that lets a waiting fiber go on: a barrier opening, an mbarrier phase completing

char* g_stacks = nullptr;
Fiber g_fibers[kMaxThreads];
#ifdef CPUSIM_FAST_SWITCH
void* g_main_sp = nullptr;
#else
ucontext_t g_main;
#endif
int g_cur = -1;
Block g_blk;
unsigned char* g_dyn = nullptr;   // dynamic shared memory of the running block, 1024-byte aligned like the hardware window
size_t g_dyn_size = 0;
constexpr uint32_t kSmemHandleBase = 4096;
const std::function<void()>* g_body = nullptr;
ThreadState g_host_state;

#ifdef CPUSIM_FAST_SWITCH
void fiber_yield() { cpusim_ctx_switch(&g_fibers[g_cur].sp, g_main_sp); }
#else
void fiber_yield() { swapcontext(&g_fibers[g_cur].ctx, &g_main); }
#endif

void fiber_entry() {
  (*g_body)();
  Fiber& f = g_fibers[g_cur];
  if (f.npending != 0) {
    fprintf(stderr, "cpusim: a thread exited with DMMAs that were never synchronised (no __syncwarp before the accumulators' use?)\n");
    abort();
  }
  f.done = true;
  // a thread that exits no longer takes part in barriers
  g_blk.alive--;
  if (g_blk.arrived > 0 && g_blk.arrived == g_blk.alive) {
    g_blk.arrived = 0;
    g_blk.gen++;
  }
  Warp& w = g_blk.warps[g_cur / 32];
  w.alive--;
  if (w.arrived > 0 && w.arrived == w.alive) {
    w.arrived = 0;
    w.gen++;
  }
  fiber_yield();   // never resumed
  abort();
}

}  // namespace

ThreadState& ts() { return g_cur >= 0 ? g_fibers[g_cur].st : g_host_state; }
void* dyn_smem() { return g_dyn; }

void sync_threads() {
  if (g_cur < 0) return;
  const unsigned gen = g_blk.gen;
  if (++g_blk.arrived == g_blk.alive) {
    g_blk.arrived = 0;
    g_blk.gen++;
    return;
  }
  while (g_blk.gen == gen) fiber_yield();
}

void sync_warp_exchange(const void* in, void* out, int src_lane, size_t bytes) {
  if (bytes > 16) {
    fprintf(stderr, "cpusim: shuffle of %zu bytes unsupported\n", bytes);
    abort();
  }
  Warp& w = g_blk.warps[g_cur / 32];
  const int lane = g_cur % 32;
  const unsigned gen = w.gen;
  memcpy(w.slot[gen & 1][lane], in, bytes);
  if (++w.arrived == w.alive) {
    w.arrived = 0;
    w.gen++;
  } else {
    while (w.gen == gen) fiber_yield();
  }
  memcpy(out, w.slot[gen & 1][src_lane], bytes);
}

// ---- emulated PTX of the GEMM kernel ------------------------------------------------------------------------------------
uint32_t smem_handle(const void* p) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  if (c < g_dyn || c >= g_dyn + g_dyn_size) {
    fprintf(stderr, "cpusim: shared-space address of a pointer outside the block's dynamic shared memory\n");
    abort();
  }
  return kSmemHandleBase + (uint32_t)(c - g_dyn);
}
void* smem_ptr(uint32_t handle) {
  if (handle < kSmemHandleBase || handle - kSmemHandleBase + 8 > g_dyn_size) {
    fprintf(stderr, "cpusim: shared-memory access outside the block's window (handle %u)\n", handle);
    abort();
  }
  return g_dyn + (handle - kSmemHandleBase);
}

static void mbar_check(MBar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1;
    b.pending = b.count;
    ++g_progress;
  }
}
static MBar& mbar_get(uint32_t bar) {
  auto it = g_blk.mbars.find(bar);
  if (it == g_blk.mbars.end()) {
    fprintf(stderr, "cpusim: mbarrier %u used before mbarrier.init\n", bar);
    abort();
  }
  return it->second;
}
void mbar_init(uint32_t bar, uint32_t count) {
  MBar b;
  b.count = b.pending = count;
  g_blk.mbars[bar] = b;
}
void mbar_arrive(uint32_t bar) {
  MBar& b = mbar_get(bar);
  if (b.pending == 0) {
    fprintf(stderr, "cpusim: more arrivals than the mbarrier expects in one phase\n");
    abort();
  }
  --b.pending;
  mbar_check(b);
}
void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  MBar& b = mbar_get(bar);
  b.tx += bytes;
  if (b.pending == 0) {
    fprintf(stderr, "cpusim: more arrivals than the mbarrier expects in one phase\n");
    abort();
  }
  --b.pending;
  mbar_check(b);
}
static void umma_complete_for(uint32_t bar);
bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  umma_complete_for(bar);   // asynchronous multiplies complete no earlier than somebody waits for them
  if (mbar_get(bar).phase != (parity & 1u)) return true;
  if (g_cur >= 0) fiber_yield();
  return mbar_get(bar).phase != (parity & 1u);
}

void tma_load_2d(uint32_t dst, const void* tensor_map, uint32_t bar, int c0, int c1) {
  SimTensorMap m;
  memcpy(&m, tensor_map, sizeof(m));
  if (m.magic != kTensorMapMagic) {
    fprintf(stderr, "cpusim: TMA load through something that is not an encoded tensor map\n");
    abort();
  }
  if (m.swizzle128 && dst % 1024 != 0 && (dst % 128 != 0)) {
    fprintf(stderr, "cpusim: TMA destination %u not 128-byte aligned\n", dst);
    abort();
  }
  for (uint32_t i1 = 0; i1 < m.box[1]; ++i1)
    for (uint32_t i0 = 0; i0 < m.box[0]; ++i0) {
      const int64_t g0 = (int64_t)c0 + i0, g1 = (int64_t)c1 + i1;
      const uint32_t es = m.elem_bytes;
      unsigned char v[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // out-of-range elements are zero-filled
      if (g0 >= 0 && g1 >= 0 && (uint64_t)g0 < m.dim[0] && (uint64_t)g1 < m.dim[1])
        memcpy(v, m.base + g0 * es + g1 * (int64_t)m.stride1, es);
      uint32_t addr = dst + (i1 * m.box[0] + i0) * es;
      if (m.swizzle128) addr ^= ((addr >> 7) & 7u) << 4;   // 16-byte chunk index XOR 128-byte line index, on the shared address
      memcpy(smem_ptr(addr), v, es);
    }
  MBar& b = mbar_get(bar);
  b.tx -= (int64_t)m.box[0] * m.box[1] * m.elem_bytes;
  mbar_check(b);
}

void tma_store_2d(const void* tensor_map, uint32_t src, int c0, int c1) {
  SimTensorMap m;
  memcpy(&m, tensor_map, sizeof(m));
  if (m.magic != kTensorMapMagic) {
    fprintf(stderr, "cpusim: TMA store through something that is not an encoded tensor map\n");
    abort();
  }
  if (src % 128 != 0) {
    fprintf(stderr, "cpusim: TMA source %u not 128-byte aligned\n", src);
    abort();
  }
  for (uint32_t i1 = 0; i1 < m.box[1]; ++i1)
    for (uint32_t i0 = 0; i0 < m.box[0]; ++i0) {
      const int64_t g0 = (int64_t)c0 + i0, g1 = (int64_t)c1 + i1;
      const uint32_t es = m.elem_bytes;
      uint32_t addr = src + (i1 * m.box[0] + i0) * es;
      if (m.swizzle128) addr ^= ((addr >> 7) & 7u) << 4;
      if (g0 >= 0 && g1 >= 0 && (uint64_t)g0 < m.dim[0] && (uint64_t)g1 < m.dim[1])   // out-of-range elements are not written
        memcpy(const_cast<char*>(m.base) + g0 * es + g1 * (int64_t)m.stride1, smem_ptr(addr), es);
    }
}

void warp_allgather16(const void* in16, void* out32x16) {
  Warp& w = g_blk.warps[g_cur / 32];
  const int lane = g_cur % 32;
  const unsigned gen = w.gen;
  if (w.alive != 32) {   // nobody can have left: every lane still has this operation in front of it
    fprintf(stderr, "cpusim: warp-collective operation with exited lanes\n");
    abort();
  }
  memcpy(w.slot[gen & 1][lane], in16, 16);
  if (++w.arrived == w.alive) {
    w.arrived = 0;
    w.gen++;
    ++g_progress;
  } else {
    while (w.gen == gen) fiber_yield();
  }
  memcpy(out32x16, w.slot[gen & 1], 32 * 16);
}

void dmma_flush() {
  Fiber& f = g_fibers[g_cur];
  if (f.npending == 0) return;
  Warp& w = g_blk.warps[g_cur / 32];
  if (w.alive != 32) {
    fprintf(stderr, "cpusim: warp-collective operation with exited lanes\n");
    abort();
  }
  const int lane = g_cur % 32, n = f.npending;
  // rendezvous: everybody's operands of this batch are in w.dmma_ab once all 32 lanes are here (same count on every lane —
  // mma.sync is warp-uniform by definition, a mismatch is a bug in the kernel)
  unsigned gen = w.gen;
  int cnt = n;
  memcpy(w.slot[gen & 1][lane], &cnt, sizeof(cnt));
  if (++w.arrived == w.alive) {
    w.arrived = 0;
    w.gen++;
    ++g_progress;
  } else {
    while (w.gen == gen) fiber_yield();
  }
  for (int l = 0; l < 32; ++l) {
    int other;
    memcpy(&other, w.slot[gen & 1][l], sizeof(other));
    if (other != n) {
      fprintf(stderr, "cpusim: lanes of a warp issued different numbers of DMMAs (%d vs %d)\n", n, other);
      abort();
    }
  }
  const int row = lane >> 2, col = 2 * (lane & 3);
  for (int i = 0; i < n; ++i) {
    double d0 = *f.pending[i].d0, d1 = *f.pending[i].d1;
    for (int k = 0; k < 4; ++k) {
      const double a = w.dmma_ab[i][row * 4 + k][0];
      d0 = __builtin_fma(a, w.dmma_ab[i][col * 4 + k][1], d0);
      d1 = __builtin_fma(a, w.dmma_ab[i][(col + 1) * 4 + k][1], d1);
    }
    *f.pending[i].d0 = d0;
    *f.pending[i].d1 = d1;
  }
  f.npending = 0;
  // second rendezvous: nobody may log into dmma_ab for the next batch before everybody has applied this one
  gen = w.gen;
  if (++w.arrived == w.alive) {
    w.arrived = 0;
    w.gen++;
    ++g_progress;
  } else {
    while (w.gen == gen) fiber_yield();
  }
}

void dmma_defer(double* d0, double* d1, double a, double b) {
  Fiber& f = g_fibers[g_cur];
  for (int i = 0; i < f.npending; ++i)
    if (f.pending[i].d0 == d0 || f.pending[i].d1 == d1 || f.pending[i].d0 == d1 || f.pending[i].d1 == d0) {
      dmma_flush();   // this accumulator still has a logged DMMA in front of it
      break;
    }
  if (f.npending == kDmmaBatch) dmma_flush();
  Warp& w = g_blk.warps[g_cur / 32];
  const int lane = g_cur % 32;
  w.dmma_ab[f.npending][lane][0] = a;
  w.dmma_ab[f.npending][lane][1] = b;
  f.pending[f.npending].d0 = d0;
  f.pending[f.npending].d1 = d1;
  ++f.npending;
}

// ---- tcgen05: tensor memory and the TF32 UMMA of gemm_f32.cu ------------------------------------------------------------------
// Descriptors are DECODED here by the bit fields of CUTLASS's cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor),
// independently of the product's encoders in common.cuh; anything outside what this emulation implements aborts.
void tmem_alloc(uint32_t slot, uint32_t ncols) {
  if (g_cur % 32 != 0) return;   // warp-collective: lane 0 does the bookkeeping
  if (ncols < 32 || ncols > kTmemCols || (ncols & (ncols - 1)) != 0 || g_blk.tmem_next + ncols > kTmemCols) {
    fprintf(stderr, "cpusim: tcgen05.alloc of %u columns (%u in use)\n", ncols, g_blk.tmem_next);
    abort();
  }
  if (g_blk.tmem.empty()) g_blk.tmem.assign((size_t)kTmemLanes * kTmemCols, 0xCDCDCDCDu);   // contents undefined
  const uint32_t base = g_blk.tmem_next;   // lane 0, column `base`
  g_blk.tmem_next += ncols;
  g_blk.tmem_live++;
  memcpy(smem_ptr(slot), &base, 4);
}
void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (g_cur % 32 != 0) return;
  if (g_blk.tmem_live <= 0 || (taddr >> 16) != 0 || (taddr & 0xFFFF) + ncols > g_blk.tmem_next) {
    fprintf(stderr, "cpusim: tcgen05.dealloc of something that was not allocated\n");
    abort();
  }
  g_blk.tmem_live--;
}
static float tf32_of(uint32_t bits) {   // what the tensor core reads of a 32-bit word: sign, 8 exponent and 10 mantissa bits
  bits &= 0xFFFFE000u;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
struct DecodedSmemDesc {
  uint32_t start, lbo, sbo;
};
static DecodedSmemDesc decode_smem_desc(uint64_t d, const char* which) {
  DecodedSmemDesc o;
  o.start = (uint32_t)(d & 0x3FFF) << 4;
  o.lbo = (uint32_t)((d >> 16) & 0x3FFF) << 4;
  o.sbo = (uint32_t)((d >> 32) & 0x3FFF) << 4;
  const unsigned version = (unsigned)(d >> 46) & 3, base_offset = (unsigned)(d >> 49) & 7, lbo_mode = (unsigned)(d >> 52) & 1,
                 layout = (unsigned)(d >> 61) & 7;
  if (version != 1 || base_offset != 0 || lbo_mode != 0 || layout != 2 /* SWIZZLE_128B */ || o.sbo != 1024) {
    fprintf(stderr, "cpusim: UMMA %s descriptor outside the emulated subset (version %u base_offset %u lbo_mode %u layout %u sbo %u)\n",
            which, version, base_offset, lbo_mode, layout, o.sbo);
    abort();
  }
  return o;
}
static float umma_operand(const DecodedSmemDesc& d, uint32_t row, uint32_t k) {
  // K-major, 128-byte swizzle: 8-row groups `sbo` bytes apart, 128 bytes per row, 16-byte chunk index XOR row-in-group on the
  // ABSOLUTE shared address (so a start address advanced by 32 B inside the swizzle span selects the next K = 8 slice)
  uint32_t addr = d.start + (row / 8) * d.sbo + (row % 8) * 128 + k * 4;
  addr ^= ((addr >> 7) & 7u) << 4;
  uint32_t bits;
  memcpy(&bits, smem_ptr(addr), 4);
  return tf32_of(bits);
}
static void umma_execute(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const unsigned sparse = idesc & 7, saturate = (idesc >> 3) & 1, cfmt = (idesc >> 4) & 3, afmt = (idesc >> 7) & 7,
                 bfmt = (idesc >> 10) & 7, neg = (idesc >> 13) & 3, amaj = (idesc >> 15) & 1, bmaj = (idesc >> 16) & 1,
                 N = ((idesc >> 17) & 0x3F) << 3, M = ((idesc >> 24) & 0x1F) << 4, shift = idesc >> 30;
  if (sparse || saturate || cfmt != 1 || afmt != 2 || bfmt != 2 || neg || amaj || bmaj || M != 128 || N < 16 || N > 256 || N % 16 ||
      shift || (idesc & ((1u << 6) | (1u << 23) | (1u << 29)))) {
    fprintf(stderr, "cpusim: UMMA instruction descriptor 0x%08x outside the emulated subset (kind::tf32, F32 accumulate, K-major, M = 128)\n",
            idesc);
    abort();
  }
  const uint32_t col0 = tmem_d & 0xFFFF;
  if ((tmem_d >> 16) != 0 || col0 + N > g_blk.tmem_next) {
    fprintf(stderr, "cpusim: UMMA accumulator outside the allocated tensor memory\n");
    abort();
  }
  const DecodedSmemDesc a = decode_smem_desc(desc_a, "A"), b = decode_smem_desc(desc_b, "B");
  static thread_local std::vector<float> av, bv;
  av.resize((size_t)M * 8);
  bv.resize((size_t)N * 8);
  for (uint32_t r = 0; r < M; ++r)
    for (uint32_t k = 0; k < 8; ++k) av[r * 8 + k] = umma_operand(a, r, k);
  for (uint32_t r = 0; r < N; ++r)
    for (uint32_t k = 0; k < 8; ++k) bv[r * 8 + k] = umma_operand(b, r, k);
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      uint32_t& cell = g_blk.tmem[(size_t)m * kTmemCols + col0 + n];
      float acc = 0.0f;
      if (accumulate) memcpy(&acc, &cell, 4);
      for (uint32_t k = 0; k < 8; ++k) acc += av[m * 8 + k] * bv[n * 8 + k];   // products of TF32 values are exact in FP32
      memcpy(&cell, &acc, 4);
    }
}
void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  g_blk.umma_queue.push_back({tmem_d, idesc, accumulate, desc_a, desc_b});
}
void umma_commit(uint32_t bar) { g_blk.umma_commits.emplace_back(bar, g_blk.umma_queue.size()); }
// Completes every commit up to the LAST pending one that arrives on `bar` (operations complete in issue order, so the commits
// in front of it complete too); `bar` == 0: everything (block exit).
static void umma_complete_upto(size_t ncommits) {
  while (g_blk.commits_done < ncommits) {
    const auto& cm = g_blk.umma_commits[g_blk.commits_done];
    for (; g_blk.umma_done < cm.second; ++g_blk.umma_done) {
      const Block::PendingUmma& u = g_blk.umma_queue[g_blk.umma_done];
      umma_execute(u.tmem_d, u.desc_a, u.desc_b, u.idesc, u.accumulate);
    }
    ++g_blk.commits_done;
    mbar_arrive(cm.first);
  }
}
static void umma_complete_for(uint32_t bar) {
  size_t upto = 0;
  for (size_t i = g_blk.commits_done; i < g_blk.umma_commits.size(); ++i)
    if (g_blk.umma_commits[i].first == bar) upto = i + 1;
  if (upto) umma_complete_upto(upto);
}
void tmem_ld_32x32(uint32_t taddr, uint32_t* v) {
  const uint32_t lane0 = taddr >> 16, col0 = taddr & 0xFFFF, warp = (uint32_t)g_cur / 32, lane = (uint32_t)g_cur % 32;
  if (lane0 != 32 * (warp % 4)) {
    fprintf(stderr, "cpusim: tcgen05.ld.32x32b by warp %u on TMEM lanes %u..: a warp may only read the lane quarter (warp %% 4)\n", warp,
            lane0);
    abort();
  }
  if (col0 + 32 > g_blk.tmem_next) {
    fprintf(stderr, "cpusim: tcgen05.ld outside the allocated tensor memory\n");
    abort();
  }
  for (uint32_t j = 0; j < 32; ++j) v[j] = g_blk.tmem[(size_t)(lane0 + lane) * kTmemCols + col0 + j];
}
void smem_rw16(uint32_t handle, void* data, bool write) {
  if (handle % 16 != 0) {
    fprintf(stderr, "cpusim: misaligned 16-byte shared-memory access\n");
    abort();
  }
  void* lo = smem_ptr(handle);
  smem_ptr(handle + 8);   // bounds of the upper half
  if (write) memcpy(lo, data, 16);
  else memcpy(data, lo, 16);
}

void named_barrier(int id, int count) {
  NamedBar& nb = g_blk.named[id & 15];
  const unsigned gen = nb.gen;
  if (++nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
    ++g_progress;
    return;
  }
  while (nb.gen == gen) fiber_yield();
}

static void execute(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

void launch(cudaStream_t stream, dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
  // the launch configuration is validated when the launch is issued, the kernel runs when its stream gets to it
  const long nthreads = (long)block.x * block.y * block.z;
  if (nthreads <= 0 || nthreads > kMaxThreads || smem > 227 * 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 ||
      grid.y > 65535 || grid.z > 65535) {
    fprintf(stderr, "cpusim: invalid launch configuration grid=(%u,%u,%u) block=(%u,%u,%u) smem=%zu\n", grid.x, grid.y,
            grid.z, block.x, block.y, block.z, smem);
    abort();
  }
  stream_submit(stream, [=]() { execute(grid, block, smem, body); });
}

static void execute(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  if (g_cur >= 0) {
    fprintf(stderr, "cpusim: nested kernel launch\n");
    abort();
  }
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > kMaxThreads || smem > 227 * 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 ||
      grid.y > 65535 || grid.z > 65535) {
    fprintf(stderr, "cpusim: invalid launch configuration grid=(%u,%u,%u) block=(%u,%u,%u) smem=%zu\n", grid.x, grid.y,
            grid.z, block.x, block.y, block.z, smem);
    abort();
  }
  if (!g_stacks) {
    g_stacks = (char*)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE,
                           -1, 0);
    if (g_stacks == MAP_FAILED) {
      perror("cpusim: mmap of fiber stacks");
      abort();
    }
  }
  g_body = &body;
  if (smem + 16 > g_dyn_size) {
    free(g_dyn);
    g_dyn_size = smem + 16;
    if (posix_memalign(reinterpret_cast<void**>(&g_dyn), 1024, g_dyn_size) != 0) abort();
  }
  memset(g_dyn, 0xCD, g_dyn_size);  // dynamic shared memory is NOT zeroed on a GPU either
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blk.nthreads = g_blk.alive = nthreads;
        g_blk.arrived = 0;
        g_blk.gen = 0;
        g_blk.warps.assign((nthreads + 31) / 32, Warp());
        g_blk.mbars.clear();
        g_blk.tmem.clear();
        g_blk.tmem_next = 0;
        g_blk.tmem_live = 0;
        g_blk.umma_queue.clear();
        g_blk.umma_commits.clear();
        g_blk.umma_done = g_blk.commits_done = 0;
        for (NamedBar& nb : g_blk.named) nb = NamedBar();
        for (int t = 0; t < nthreads; ++t) {
          Fiber& f = g_fibers[t];
          f.started = f.done = false;
          f.npending = 0;
          f.st.thread.x = t % block.x;
          f.st.thread.y = (t / block.x) % block.y;
          f.st.thread.z = t / (block.x * block.y);
          f.st.block.x = bx;
          f.st.block.y = by;
          f.st.block.z = bz;
          f.st.bdim.x = block.x;
          f.st.bdim.y = block.y;
          f.st.bdim.z = block.z;
          f.st.gdim.x = grid.x;
          f.st.gdim.y = grid.y;
          f.st.gdim.z = grid.z;
          g_blk.warps[t / 32].alive++;
        }
        int remaining = nthreads;
        long idle_rounds = 0;
        while (remaining > 0) {
          int progressed = 0;
          const uint64_t progress_before = g_progress;
          for (int t = 0; t < nthreads; ++t) {
            Fiber& f = g_fibers[t];
            if (f.done) continue;
            if (!f.started) {
#ifdef CPUSIM_FAST_SWITCH
              // stack as cpusim_ctx_switch expects it: six callee-saved registers, then the address `ret` jumps to; the
              // entry function then sees the alignment of a normal call (rsp = 16n + 8)
              uintptr_t top = (reinterpret_cast<uintptr_t>(g_stacks + kStack * (t + 1))) & ~(uintptr_t)15;
              void** sp = reinterpret_cast<void**>(top);
              *--sp = nullptr;                                    // fake return address of the entry function
              *--sp = reinterpret_cast<void*>(&fiber_entry);      // ret target
              for (int r = 0; r < 6; ++r) *--sp = nullptr;        // rbp rbx r12 r13 r14 r15
              f.sp = sp;
#else
              getcontext(&f.ctx);
              f.ctx.uc_stack.ss_sp = g_stacks + kStack * t;
              f.ctx.uc_stack.ss_size = kStack;
              f.ctx.uc_link = nullptr;
              makecontext(&f.ctx, fiber_entry, 0);
#endif
              f.started = true;
            }
            const unsigned bgen = g_blk.gen, wgen = g_blk.warps[t / 32].gen;
            g_cur = t;
#ifdef CPUSIM_FAST_SWITCH
            cpusim_ctx_switch(&g_main_sp, f.sp);
#else
            swapcontext(&g_main, &f.ctx);
#endif
            g_cur = -1;
            if (f.done) {
              --remaining;
              ++progressed;
            } else if (bgen != g_blk.gen || wgen != g_blk.warps[t / 32].gen) {
              ++progressed;
            }
          }
          // a full round in which nobody finished and no barrier opened twice in a row = divergent barrier (deadlock)
          if (g_progress != progress_before) ++progressed;
          idle_rounds = progressed ? 0 : idle_rounds + 1;
          if (idle_rounds > 2) {
            fprintf(stderr, "cpusim: block (%u,%u,%u) deadlocked at a barrier (%d threads waiting, %d alive)\n", bx, by, bz,
                    g_blk.arrived, g_blk.alive);
            abort();
          }
        }
        if (g_blk.umma_done != g_blk.umma_queue.size() && g_blk.commits_done == g_blk.umma_commits.size()) {
          fprintf(stderr, "cpusim: block (%u,%u,%u) exited with tcgen05.mma operations that were never committed\n", bx, by, bz);
          abort();
        }
        if (g_blk.tmem_live != 0) {
          fprintf(stderr, "cpusim: block (%u,%u,%u) exited with tensor memory still allocated\n", bx, by, bz);
          abort();
        }
      }
  g_body = nullptr;
}

}  // namespace cpusim
