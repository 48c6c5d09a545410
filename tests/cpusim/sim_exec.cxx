// tests/cpusim — kernel executor of the CPU functional simulator (TEST INFRASTRUCTURE ONLY, see sim_device.h).
// One block at a time; the threads of a block are ucontext fibers scheduled round-robin; a fiber leaves the CPU only at
// a barrier (__syncthreads, warp shuffle) or when its kernel body returns.
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <vector>

#include "sim_device.h"
#include "sim_internal.h"

namespace cpusim {

namespace {

constexpr size_t kStack = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Fiber {
  ucontext_t ctx;
  ThreadState st;
  bool started = false, done = false;
};

struct Warp {
  int alive = 0, arrived = 0;
  unsigned gen = 0;
  alignas(16) unsigned char slot[2][32][16];
};

struct Block {
  int nthreads = 0, alive = 0, arrived = 0;
  unsigned gen = 0;
  std::vector<Warp> warps;
};

char* g_stacks = nullptr;
Fiber g_fibers[kMaxThreads];
ucontext_t g_main;
int g_cur = -1;
Block g_blk;
std::vector<unsigned char> g_dyn;
const std::function<void()>* g_body = nullptr;
ThreadState g_host_state;

void fiber_yield() { swapcontext(&g_fibers[g_cur].ctx, &g_main); }

void fiber_entry() {
  (*g_body)();
  Fiber& f = g_fibers[g_cur];
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
  swapcontext(&f.ctx, &g_main);
}

}  // namespace

ThreadState& ts() { return g_cur >= 0 ? g_fibers[g_cur].st : g_host_state; }
void* dyn_smem() { return g_dyn.data(); }

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
  g_dyn.assign(smem + 16, 0xCD);  // dynamic shared memory is NOT zeroed on a GPU either
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blk.nthreads = g_blk.alive = nthreads;
        g_blk.arrived = 0;
        g_blk.gen = 0;
        g_blk.warps.assign((nthreads + 31) / 32, Warp());
        for (int t = 0; t < nthreads; ++t) {
          Fiber& f = g_fibers[t];
          f.started = f.done = false;
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
          for (int t = 0; t < nthreads; ++t) {
            Fiber& f = g_fibers[t];
            if (f.done) continue;
            if (!f.started) {
              getcontext(&f.ctx);
              f.ctx.uc_stack.ss_sp = g_stacks + kStack * t;
              f.ctx.uc_stack.ss_size = kStack;
              f.ctx.uc_link = nullptr;
              makecontext(&f.ctx, fiber_entry, 0);
              f.started = true;
            }
            const unsigned bgen = g_blk.gen, wgen = g_blk.warps[t / 32].gen;
            g_cur = t;
            swapcontext(&g_main, &f.ctx);
            g_cur = -1;
            if (f.done) {
              --remaining;
              ++progressed;
            } else if (bgen != g_blk.gen || wgen != g_blk.warps[t / 32].gen) {
              ++progressed;
            }
          }
          // a full round in which nobody finished and no barrier opened twice in a row = divergent barrier (deadlock)
          idle_rounds = progressed ? 0 : idle_rounds + 1;
          if (idle_rounds > 2) {
            fprintf(stderr, "cpusim: block (%u,%u,%u) deadlocked at a barrier (%d threads waiting, %d alive)\n", bx, by, bz,
                    g_blk.arrived, g_blk.alive);
            abort();
          }
        }
      }
  g_body = nullptr;
}

}  // namespace cpusim
