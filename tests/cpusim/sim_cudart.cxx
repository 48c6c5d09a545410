// tests/cpusim — the CUDA runtime calls the product's host code makes, re-implemented on host memory with SYNCHRONOUS
// streams (TEST INFRASTRUCTURE ONLY, see sim_device.h).  Declarations come from the real <cuda_runtime.h>, so the
// signatures cannot drift.  Extras a GPU does not give for free:
//   * every "device" allocation is filled with signalling garbage (NaN bit pattern) — code that relies on zeroed
//     cudaMalloc memory shows up as NaNs;
//   * 64-byte canaries on both sides of every allocation, verified at cudaFree, at every synchronize and at exit;
//   * a host<->device registry: cudaMemcpy* with a stated direction is checked against where the pointers really live,
//     cudaPointerGetAttributes answers from it (so the staging logic of the C ABI is exercised as on a GPU).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <deque>
#include <functional>
#include <map>
#include <vector>

#include "sim_internal.h"

namespace {

constexpr size_t kGuard = 64;
constexpr unsigned char kCanary = 0xA5;

struct Alloc {
  size_t bytes;
  bool device;     // false: pinned host
  bool shm;        // backed by a POSIX shared-memory object (large device allocations): can be exported with cudaIpcGetMemHandle
  bool peer;       // another process's allocation mapped here by cudaIpcOpenMemHandle (its owner checks the canaries)
  char name[48];
};
constexpr size_t kShmThreshold = 4 * 1024;   // device allocations from this size on are shared-memory objects (exportable)
constexpr size_t kShmHeader = 256;
std::map<uintptr_t, Alloc> g_allocs;   // user base -> allocation
cudaError_t g_last = cudaSuccess;
int g_device = 0;

struct SimEvent {
  double ms = 0.0;
  uint64_t gen_enq = 0;    // cudaEventRecord calls issued
  uint64_t gen_done = 0;   // ... and executed
};

enum TaskKind { T_RUN, T_RECORD, T_WAIT, T_NCCL, T_WAITVAL };
struct Task {
  TaskKind kind = T_RUN;
  uint64_t seq = 0;
  std::function<void()> fn;
  SimEvent* ev = nullptr;
  uint64_t gen = 0;
  cpusim::NcclBatch* batch = nullptr;
  bool started = false;
  const uint32_t* addr = nullptr;   // T_WAITVAL: cuStreamWaitValue32
  uint32_t value = 0;
  unsigned wflags = 0;
};

bool waitval_satisfied(const uint32_t* addr, uint32_t value, unsigned flags) {
  const uint32_t v = __atomic_load_n(addr, __ATOMIC_ACQUIRE);
  switch (flags & 3u) {
    case CU_STREAM_WAIT_VALUE_GEQ: return (int32_t)(v - value) >= 0;
    case CU_STREAM_WAIT_VALUE_EQ: return v == value;
    case CU_STREAM_WAIT_VALUE_AND: return (v & value) != 0;
    default: return (~(v | value)) != 0;   // NOR
  }
}
struct SimStream {
  int id = 0;
  std::deque<Task> q;
};
enum Policy { P_SYNC, P_FIFO, P_LIFO, P_RANDOM };
Policy g_policy = P_SYNC;
uint64_t g_rng = 0x9E3779B97F4A7C15ull;
SimStream g_null;
std::vector<SimStream*> g_streams{&g_null};
uint64_t g_seq = 0;
int g_next_stream = 1;
bool g_draining = false;
uint64_t g_executed = 0, g_reordered = 0, g_max_seq_run = 0;   // tasks run / run after a task that was enqueued later

struct PolicyInit {
  PolicyInit() {
    const char* e = getenv("CPUSIM_SCHED");
    if (!e || !strcmp(e, "sync")) return;
    if (!strcmp(e, "fifo")) g_policy = P_FIFO;
    else if (!strcmp(e, "lifo")) g_policy = P_LIFO;
    else if (!strncmp(e, "random", 6)) {
      g_policy = P_RANDOM;
      if (e[6] == ':') g_rng ^= strtoull(e + 7, nullptr, 10) * 0xD1B54A32D192ED03ull;
    } else {
      fprintf(stderr, "cpusim: unknown CPUSIM_SCHED=%s (sync | fifo | lifo | random:<seed>)\n", e);
      abort();
    }
  }
} g_policy_init;

SimStream* S(cudaStream_t s) {
  // 0, cudaStreamLegacy (1) and cudaStreamPerThread (2) are one queue here; the product's own streams are all created
  // cudaStreamNonBlocking, so the legacy stream's implicit synchronisation with blocking streams never applies
  return reinterpret_cast<uintptr_t>(s) <= 2 ? &g_null : reinterpret_cast<SimStream*>(s);
}

void push(SimStream* st, Task&& t) {
  t.seq = g_seq++;
  st->q.push_back(std::move(t));
}

double now_ms() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

const Alloc* find_alloc(const void* p, uintptr_t* base = nullptr) {
  uintptr_t a = (uintptr_t)p;
  auto it = g_allocs.upper_bound(a);
  if (it == g_allocs.begin()) return nullptr;
  --it;
  if (a < it->first + (it->second.bytes ? it->second.bytes : 1)) {
    if (base) *base = it->first;
    return &it->second;
  }
  return nullptr;
}

void check_canaries(const char* when) {
  for (auto& kv : g_allocs) {
    if (kv.second.peer) continue;
    const unsigned char* lo = (const unsigned char*)kv.first - kGuard;
    const unsigned char* hi = (const unsigned char*)kv.first + kv.second.bytes;
    for (size_t i = 0; i < kGuard; ++i)
      if (lo[i] != kCanary || hi[i] != kCanary) {
        fprintf(stderr, "cpusim: OUT-OF-BOUNDS WRITE %s allocation %p (%zu bytes, %s), detected at %s\n",
                lo[i] != kCanary ? "below" : "above", (void*)kv.first, kv.second.bytes,
                kv.second.device ? "device" : "pinned host", when);
        abort();
      }
  }
}

void fill_garbage(unsigned char* user, size_t bytes) {
  // signalling-NaN-ish garbage: 0x7FF4A5A5A5A5A5A5
  uint64_t pat = 0x7FF4A5A5A5A5A5A5ull;
  for (size_t i = 0; i + 8 <= bytes; i += 8) memcpy(user + i, &pat, 8);
  for (size_t i = bytes & ~(size_t)7; i < bytes; ++i) user[i] = 0xA7;
}

// A test that dies with abort() cannot unlink its shared-memory objects: whoever comes next removes what dead processes left.
void remove_stale_objects() {
  DIR* d = opendir("/dev/shm");
  if (!d) return;
  while (dirent* e = readdir(d)) {
    int pid = 0;
    if (sscanf(e->d_name, "cpusim.mem.%d.", &pid) != 1 && sscanf(e->d_name, "cpusim.%d.", &pid) != 1) continue;
    if (pid > 0 && kill(pid, 0) != 0 && errno == ESRCH) {
      char name[300];
      snprintf(name, sizeof(name), "/%s", e->d_name);
      shm_unlink(name);
    }
  }
  closedir(d);
}

void* sim_alloc(size_t bytes, bool device) {
  static bool cleaned = false;
  if (!cleaned) {
    cleaned = true;
    remove_stale_objects();
  }
  Alloc a;
  memset(&a, 0, sizeof(a));
  a.bytes = bytes;
  a.device = device;
  unsigned char* user = nullptr;
  if (device && bytes >= kShmThreshold) {
    // a named shared-memory object, so that another rank-process can map it (CUDA IPC); same layout as the heap case
    static int counter = 0;
    snprintf(a.name, sizeof(a.name), "/cpusim.mem.%d.%d", (int)getpid(), counter++);
    const size_t total = kShmHeader + bytes + kGuard;
    int fd = shm_open(a.name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)total) != 0) {
      if (fd >= 0) close(fd);
      return nullptr;
    }
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
      shm_unlink(a.name);
      return nullptr;
    }
    a.shm = true;
    user = (unsigned char*)m + kShmHeader;
  } else {
    void* raw = nullptr;
    if (posix_memalign(&raw, 256, bytes + 2 * kGuard + 256) != 0) return nullptr;
    // keep the user pointer 256-byte aligned like cudaMalloc: guard sits in the 256 bytes in front of it
    user = (unsigned char*)raw + 256;
  }
  memset(user - kGuard, kCanary, kGuard);
  memset(user + bytes, kCanary, kGuard);
  if (device) fill_garbage(user, bytes);
  g_allocs[(uintptr_t)user] = a;
  return user;
}

cudaError_t sim_free(void* p, bool device) {
  if (!p) return cudaSuccess;
  auto it = g_allocs.find((uintptr_t)p);
  if (it == g_allocs.end() || it->second.device != device || it->second.peer) {
    fprintf(stderr, "cpusim: %s of a pointer that is not a live %s allocation: %p\n", device ? "cudaFree" : "cudaFreeHost",
            device ? "device" : "pinned", p);
    abort();
  }
  check_canaries(device ? "cudaFree" : "cudaFreeHost");
  memset(p, 0xDD, it->second.bytes);  // use-after-free shows up as garbage
  if (it->second.shm) {
    shm_unlink(it->second.name);
    munmap((unsigned char*)p - kShmHeader, kShmHeader + it->second.bytes + kGuard);
  } else {
    free((unsigned char*)p - 256);
  }
  g_allocs.erase(it);
  return cudaSuccess;
}

void check_range(const void* p, size_t bytes, bool want_device, const char* what) {
  if (bytes == 0) return;
  uintptr_t base = 0;
  const Alloc* a = find_alloc(p, &base);
  if (want_device) {
    if (!a || !a->device) {
      fprintf(stderr, "cpusim: %s: %p is not device memory\n", what, p);
      abort();
    }
  } else if (a && a->device) {
    fprintf(stderr, "cpusim: %s: %p is device memory where a host pointer is required\n", what, p);
    abort();
  }
  if (a && (uintptr_t)p + bytes > base + a->bytes) {
    fprintf(stderr, "cpusim: %s: range %p + %zu runs past the end of its allocation (%p + %zu)\n", what, p, bytes,
            (void*)base, a->bytes);
    abort();
  }
}

void check_copy(void* dst, const void* src, size_t dst_span, size_t src_span, cudaMemcpyKind kind, const char* what) {
  switch (kind) {
    case cudaMemcpyHostToDevice:
      check_range(dst, dst_span, true, what);
      check_range(src, src_span, false, what);
      break;
    case cudaMemcpyDeviceToHost:
      check_range(dst, dst_span, false, what);
      check_range(src, src_span, true, what);
      break;
    case cudaMemcpyDeviceToDevice:
      check_range(dst, dst_span, true, what);
      check_range(src, src_span, true, what);
      break;
    case cudaMemcpyHostToHost:
      check_range(dst, dst_span, false, what);
      check_range(src, src_span, false, what);
      break;
    default:
      break;  // cudaMemcpyDefault: anything goes
  }
}

double now_s2() { return now_ms() * 1e-3; }

// Runs queued work until pred() holds.  Exactly one runnable stream head is executed per iteration (chosen by the policy);
// every NCCL batch that has started is progressed a step per iteration, so batches on different streams overlap as they
// do on a GPU.
void drain(const std::function<bool()>& pred, const char* why) {
  if (g_policy == P_SYNC) return;
  if (g_draining) {
    fprintf(stderr, "cpusim: synchronisation (%s) from inside a queued task\n", why);
    abort();
  }
  g_draining = true;
  static double timeout = getenv("CPUSIM_TIMEOUT") ? atof(getenv("CPUSIM_TIMEOUT")) : 60.0;
  static std::vector<cpusim::NcclBatch*> active;
  double last = now_s2();
  long idle = 0;
  while (!pred()) {
    bool any = false;
    SimStream* pick = nullptr;
    int ncand = 0;
    bool external_wait = false;
    for (SimStream* st : g_streams) {
      if (st->q.empty()) continue;
      Task& t = st->q.front();
      if (t.kind == T_NCCL && t.started) continue;
      if (t.kind == T_WAIT && t.ev->gen_done < t.gen) continue;
      if (t.kind == T_WAITVAL && !waitval_satisfied(t.addr, t.value, t.wflags)) {
        external_wait = true;   // another process (or a later task) has to write the value
        continue;
      }
      ++ncand;
      if (!pick) {
        pick = st;
      } else if (g_policy == P_FIFO) {
        if (t.seq < pick->q.front().seq) pick = st;
      } else if (g_policy == P_LIFO) {
        if (t.seq > pick->q.front().seq) pick = st;
      } else {  // reservoir sampling
        g_rng ^= g_rng << 13;
        g_rng ^= g_rng >> 7;
        g_rng ^= g_rng << 17;
        if (g_rng % (uint64_t)ncand == 0) pick = st;
      }
    }
    if (pick) {
      Task& t = pick->q.front();
      ++g_executed;
      if (t.seq < g_max_seq_run) ++g_reordered;
      else g_max_seq_run = t.seq;
      if (t.kind == T_NCCL) {
        t.started = true;
        active.push_back(t.batch);
      } else {
        Task run = std::move(t);
        pick->q.pop_front();
        if (run.kind == T_RUN) {
          run.fn();
        } else if (run.kind == T_RECORD) {
          if (run.gen > run.ev->gen_done) run.ev->gen_done = run.gen;
          run.ev->ms = now_ms();
        }
      }
      any = true;
    }
    if (!active.empty()) {
      any |= cpusim::nccl_progress(active);
      for (size_t i = 0; i < active.size();) {
        if (!cpusim::nccl_done(active[i])) {
          ++i;
          continue;
        }
        cpusim::NcclBatch* b = active[i];
        active.erase(active.begin() + i);
        for (SimStream* st : g_streams)
          if (!st->q.empty() && st->q.front().kind == T_NCCL && st->q.front().batch == b) st->q.pop_front();
        cpusim::nccl_finish(b);
        any = true;
      }
    }
    if (any) {
      idle = 0;
      last = now_s2();
      continue;
    }
    if (!pick && active.empty() && !external_wait) {
      fprintf(stderr, "cpusim: %s can never complete: every queued stream waits on an event that nobody will record\n", why);
      abort();
    }
    if (++idle < 200) {
      sched_yield();
    } else {
      timespec ts = {0, 50000};
      nanosleep(&ts, nullptr);
      if ((idle & 1023) == 0 && now_s2() - last > timeout) {
        fprintf(stderr, "cpusim nccl: NO PROGRESS for %.0f s in %s — deadlock in the communication schedule?  Active operations:\n",
                timeout, why);
        for (cpusim::NcclBatch* b : active) cpusim::nccl_describe(b);
        for (SimStream* st : g_streams)
          if (!st->q.empty() && st->q.front().kind == T_WAITVAL)
            fprintf(stderr, "  stream %d waits for *%p (now %u) to reach %u\n", st->id, (const void*)st->q.front().addr,
                    *st->q.front().addr, st->q.front().value);
        abort();
      }
    }
  }
  g_draining = false;
}

void drain_stream(cudaStream_t s) {
  SimStream* st = S(s);
  drain([st]() { return st->q.empty(); }, "a stream synchronisation");
}
void drain_all(const char* why) {
  drain([]() {
    for (SimStream* st : g_streams)
      if (!st->q.empty()) return false;
    return true;
  }, why);
}

struct AtExit {
  ~AtExit() {
    check_canaries("process exit");
    for (auto& kv : g_allocs)
      if (kv.second.shm && !kv.second.peer) shm_unlink(kv.second.name);   // nothing of ours stays behind in /dev/shm
    if (getenv("CPUSIM_VERBOSE"))
      fprintf(stderr, "cpusim: %llu queued tasks executed, %llu of them after a task that was enqueued later\n",
              (unsigned long long)g_executed, (unsigned long long)g_reordered);
  }
} g_at_exit;

// cuStreamWaitValue32: a stream memory operation — the stream goes on when *addr satisfies the condition
CUresult fake_stream_wait_value32(CUstream st, CUdeviceptr addr, cuuint32_t value, unsigned int flags);

CUresult fake_encode_tiled(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* gdim,
                           const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estride, CUtensorMapInterleave il,
                           CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  static_assert(sizeof(cpusim::SimTensorMap) <= sizeof(CUtensorMap), "descriptor does not fit");
  // the driver's own argument rules for what the product encodes (2-D FP64 tiles)
  if ((dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 && dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT32) || rank != 2 || il != CU_TENSOR_MAP_INTERLEAVE_NONE)
    return CUDA_ERROR_INVALID_VALUE;
  const uint32_t es = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT64 ? 8 : 4;
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0 || gstride[0] % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
  if (box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256 || estride[0] != 1 || estride[1] != 1) return CUDA_ERROR_INVALID_VALUE;
  if (sw == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * es > 128) return CUDA_ERROR_INVALID_VALUE;
  // extents: each dimension 1 .. 2^32, row pitch below 2^40 bytes and not shorter than a row, box no larger than the tensor's limits allow
  for (int d = 0; d < 2; ++d)
    if (gdim[d] == 0 || gdim[d] > (1ull << 32)) return CUDA_ERROR_INVALID_VALUE;
  if (gstride[0] >= (1ull << 40) || gstride[0] < gdim[0] * es) return CUDA_ERROR_INVALID_VALUE;
  if ((box[0] * es) % 16 != 0) return CUDA_ERROR_INVALID_VALUE;   // inner box extent: whole 16-byte units
  if (sw != CU_TENSOR_MAP_SWIZZLE_128B && sw != CU_TENSOR_MAP_SWIZZLE_NONE) return CUDA_ERROR_INVALID_VALUE;
  memset(out, 0, sizeof(*out));
  cpusim::SimTensorMap m;
  m.magic = cpusim::kTensorMapMagic;
  m.base = static_cast<const char*>(base);
  m.dim[0] = gdim[0];
  m.dim[1] = gdim[1];
  m.stride1 = gstride[0];
  m.box[0] = box[0];
  m.box[1] = box[1];
  m.swizzle128 = sw == CU_TENSOR_MAP_SWIZZLE_128B;
  m.elem_bytes = es;
  memcpy(out, &m, sizeof(m));
  return CUDA_SUCCESS;
}

}  // namespace

namespace cpusim {

void stream_submit(cudaStream_t s, std::function<void()> fn) {
  if (g_policy == P_SYNC) {
    fn();
    return;
  }
  Task t;
  t.kind = T_RUN;
  t.fn = std::move(fn);
  push(S(s), std::move(t));
}

void stream_submit_nccl(cudaStream_t s, NcclBatch* b) {
  if (g_policy == P_SYNC) {
    run_batch_blocking(b);
    return;
  }
  Task t;
  t.kind = T_NCCL;
  t.batch = b;
  push(S(s), std::move(t));
}

}  // namespace cpusim

namespace {
CUresult fake_stream_wait_value32(CUstream st, CUdeviceptr addr, cuuint32_t value, unsigned int flags) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(addr);
  check_range(p, 4, true, "cuStreamWaitValue32");
  if (g_policy == P_SYNC) {   // everything before it has executed; wait for the writer (another process) right here
    static double timeout = getenv("CPUSIM_TIMEOUT") ? atof(getenv("CPUSIM_TIMEOUT")) : 60.0;
    const double t0 = now_ms();
    long spins = 0;
    while (!waitval_satisfied(p, value, flags)) {
      if (++spins < 200) {
        sched_yield();
      } else {
        timespec ts = {0, 50000};
        nanosleep(&ts, nullptr);
        if ((now_ms() - t0) * 1e-3 > timeout) {
          fprintf(stderr, "cpusim: cuStreamWaitValue32 NO PROGRESS for %.0f s: *%p is %u, waiting for %u\n", timeout, (const void*)p, *p,
                  value);
          abort();
        }
      }
    }
    return CUDA_SUCCESS;
  }
  Task t;
  t.kind = T_WAITVAL;
  t.addr = p;
  t.value = value;
  t.wflags = flags;
  push(S(reinterpret_cast<cudaStream_t>(st)), std::move(t));
  return CUDA_SUCCESS;
}
}  // namespace

extern "C" {

// helpers for the Python side of the tests (simtorch.py)
void* cpusim_malloc(size_t bytes) { return sim_alloc(bytes, true); }
void cpusim_check(void) {   // what torch.cuda.synchronize() becomes
  drain_all("torch.cuda.synchronize");
  check_canaries("cpusim_check");
}
void cpusim_require_device_range(const void* p, size_t bytes, const char* what) { check_range(p, bytes, true, what); }
void cpusim_sched_stats(unsigned long long* executed, unsigned long long* reordered) {
  *executed = g_executed;
  *reordered = g_reordered;
}
int cpusim_is_device(const void* p) {
  const Alloc* a = find_alloc(p);
  return a && a->device;
}

cudaError_t cudaGetDeviceCount(int* n) {
  *n = 16;  // one simulated device per rank-process of a node
  return cudaSuccess;
}
cudaError_t cudaGetDevice(int* d) {
  *d = g_device;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) {
  g_device = d;
  return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* prop, int) {
  memset(prop, 0, sizeof(*prop));
  snprintf(prop->name, sizeof(prop->name), "cpusim (functional simulator, not a GPU)");
  prop->major = 10;
  prop->minor = 0;
  const char* s = getenv("CPUSIM_SMS");
  prop->multiProcessorCount = s ? atoi(s) : 3;
  prop->sharedMemPerBlockOptin = 227 * 1024;
  prop->totalGlobalMem = (size_t)8 << 30;
  return cudaSuccess;
}
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cpusim error"; }
cudaError_t cudaGetLastError(void) {
  cudaError_t e = g_last;
  g_last = cudaSuccess;
  return e;
}
cudaError_t cudaDeviceSynchronize(void) {
  drain_all("cudaDeviceSynchronize");
  check_canaries("cudaDeviceSynchronize");
  return cudaSuccess;
}

cudaError_t cudaMalloc(void** p, size_t bytes) {
  *p = sim_alloc(bytes, true);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
  if (p) drain_all("cudaFree");  // cudaFree synchronises the device
  return sim_free(p, true);
}
cudaError_t cudaMallocHost(void** p, size_t bytes) {
  *p = sim_alloc(bytes, false);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMallocHost(p, bytes); }
cudaError_t cudaFreeHost(void* p) { return sim_free(p, false); }

cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* attr, const void* p) {
  memset(attr, 0, sizeof(*attr));
  const Alloc* a = find_alloc(p);
  attr->type = !a ? cudaMemoryTypeUnregistered : (a->device ? cudaMemoryTypeDevice : cudaMemoryTypeHost);
  attr->device = g_device;
  attr->devicePointer = a ? const_cast<void*>(p) : nullptr;   // pinned host memory is addressable from the device (UVA)
  attr->hostPointer = a && a->device ? nullptr : const_cast<void*>(p);
  return cudaSuccess;
}

static bool pageable(const void* p) { return find_alloc(p) == nullptr; }

// CUDA's rules for "async" copies that touch pageable host memory: host->device waits for the stream, stages the source
// and may return before the DMA has landed (modelled: stream sync, then the copy happens now); device->pageable host
// returns only when the copy has completed (modelled: enqueue, then stream sync).  Pinned and device memory: truly async.
static void submit_copy(cudaStream_t st, void* dst, const void* src, cudaMemcpyKind kind, std::function<void()> body) {
  const bool src_pageable = pageable(src), dst_pageable = pageable(dst);
  if (src_pageable) {
    drain_stream(st);
    body();
    return;
  }
  cpusim::stream_submit(st, std::move(body));
  if (dst_pageable) drain_stream(st);
  (void)kind;
}

cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  check_copy(dst, src, bytes, bytes, kind, "cudaMemcpyAsync");
  submit_copy(st, dst, src, kind, [=]() { memmove(dst, src, bytes); });
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  check_copy(dst, src, bytes, bytes, kind, "cudaMemcpy");
  cpusim::stream_submit(nullptr, [=]() { memmove(dst, src, bytes); });
  drain_stream(nullptr);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind kind, cudaStream_t st) {
  if (width == 0 || height == 0) return cudaSuccess;
  if (width > dpitch || width > spitch) {
    fprintf(stderr, "cpusim: cudaMemcpy2DAsync: width %zu exceeds a pitch (%zu, %zu)\n", width, dpitch, spitch);
    g_last = cudaErrorInvalidPitchValue;
    return cudaErrorInvalidPitchValue;
  }
  check_copy(dst, src, dpitch * (height - 1) + width, spitch * (height - 1) + width, kind, "cudaMemcpy2DAsync");
  submit_copy(st, dst, src, kind, [=]() {
    for (size_t r = 0; r < height; ++r) memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
  });
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t st) {
  check_range(p, bytes, true, "cudaMemsetAsync");
  cpusim::stream_submit(st, [=]() { memset(p, v, bytes); });
  return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) {
  cudaMemsetAsync(p, v, bytes, nullptr);
  drain_stream(nullptr);
  return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  SimStream* st = new SimStream();
  st->id = g_next_stream++;
  g_streams.push_back(st);
  *s = reinterpret_cast<cudaStream_t>(st);
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
// priorities only bias the hardware's block scheduler; the simulator's stream order policies ignore them
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int) { return cudaStreamCreateWithFlags(s, flags); }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) {
  if (lo) *lo = 0;
  if (hi) *hi = -5;
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  drain_stream(s);  // CUDA lets pending work finish before the stream goes away
  SimStream* st = S(s);
  g_streams.erase(std::find(g_streams.begin(), g_streams.end(), st));
  delete st;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  drain_stream(s);
  check_canaries("cudaStreamSynchronize");
  return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
  if (!e) {
    fprintf(stderr, "cpusim: cudaStreamWaitEvent on a null event\n");
    abort();
  }
  SimEvent* ev = reinterpret_cast<SimEvent*>(e);
  if (g_policy == P_SYNC || ev->gen_enq == 0) return cudaSuccess;  // nothing recorded yet: a no-op, as in CUDA
  Task t;
  t.kind = T_WAIT;
  t.ev = ev;
  t.gen = ev->gen_enq;  // the most recent record at the time of THIS call
  push(S(s), std::move(t));
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = reinterpret_cast<cudaEvent_t>(new SimEvent());
  return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  SimEvent* ev = reinterpret_cast<SimEvent*>(e);
  if (ev->gen_done >= ev->gen_enq) delete ev;  // else: queued tasks still point at it; CUDA defers the release too (leaked here)
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  SimEvent* ev = reinterpret_cast<SimEvent*>(e);
  ev->gen_enq++;
  if (g_policy == P_SYNC) {
    ev->gen_done = ev->gen_enq;
    ev->ms = now_ms();
    return cudaSuccess;
  }
  Task t;
  t.kind = T_RECORD;
  t.ev = ev;
  t.gen = ev->gen_enq;
  push(S(s), std::move(t));
  return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t e) {
  SimEvent* ev = reinterpret_cast<SimEvent*>(e);
  return ev->gen_done >= ev->gen_enq ? cudaSuccess : cudaErrorNotReady;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  SimEvent* ev = reinterpret_cast<SimEvent*>(e);
  drain([&]() { return ev->gen_done >= ev->gen_enq; }, "cudaEventSynchronize");
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  SimEvent *ea = reinterpret_cast<SimEvent*>(a), *eb = reinterpret_cast<SimEvent*>(b);
  if (ea->gen_enq == 0 || eb->gen_enq == 0) return cudaErrorInvalidResourceHandle;
  if (ea->gen_done < ea->gen_enq || eb->gen_done < eb->gen_enq) return cudaErrorNotReady;
  *ms = (float)(eb->ms - ea->ms);
  return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* value, cudaDeviceAttr attr, int) {
  *value = (attr == cudaDevAttrCanFlushRemoteWrites) ? 1 : 0;   // the only one the product asks for
  return cudaSuccess;
}

// ---- CUDA IPC: large device allocations are shared-memory objects, the handle is the object's name ------------------------
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  auto it = g_allocs.find((uintptr_t)p);
  if (it == g_allocs.end() || !it->second.shm || it->second.peer) {
    g_last = cudaErrorInvalidValue;   // not the base of an exportable allocation
    return cudaErrorInvalidValue;
  }
  memset(h, 0, sizeof(*h));
  static_assert(sizeof(h->reserved) >= sizeof(it->second.name), "handle too small");
  memcpy(h->reserved, it->second.name, sizeof(it->second.name));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  // fault injection: CPUSIM_IPC_FAIL_RANK=r makes every mapping attempt of rank r fail (the library must then agree, on all
  // ranks of the communicator, to stay on NCCL)
  if (const char* fr = getenv("CPUSIM_IPC_FAIL_RANK")) {
    const char* me = getenv("RANK");
    if (me && atoi(me) == atoi(fr)) {
      g_last = cudaErrorInvalidValue;
      return cudaErrorInvalidValue;
    }
  }
  char name[64];
  memcpy(name, h.reserved, sizeof(name));
  name[63] = 0;
  int fd = shm_open(name, O_RDWR, 0600);
  struct stat sb;
  if (fd < 0 || fstat(fd, &sb) != 0) {
    if (fd >= 0) close(fd);
    g_last = cudaErrorInvalidValue;
    return cudaErrorInvalidValue;
  }
  void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) {
    g_last = cudaErrorMemoryAllocation;
    return cudaErrorMemoryAllocation;
  }
  Alloc a;
  memset(&a, 0, sizeof(a));
  a.bytes = (size_t)sb.st_size - kShmHeader - kGuard;
  a.device = true;
  a.shm = true;
  a.peer = true;
  snprintf(a.name, sizeof(a.name), "%s", name);
  unsigned char* user = (unsigned char*)m + kShmHeader;
  g_allocs[(uintptr_t)user] = a;
  *p = user;
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
  auto it = g_allocs.find((uintptr_t)p);
  if (it == g_allocs.end() || !it->second.peer) {
    g_last = cudaErrorInvalidValue;
    return cudaErrorInvalidValue;
  }
  munmap((unsigned char*)p - kShmHeader, kShmHeader + it->second.bytes + kGuard);
  g_allocs.erase(it);
  return cudaSuccess;
}

cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  if (strcmp(symbol, "cuTensorMapEncodeTiled") == 0) {
    *fn = reinterpret_cast<void*>(&fake_encode_tiled);
    if (q) *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  if (strcmp(symbol, "cuStreamWaitValue32") == 0) {
    *fn = reinterpret_cast<void*>(&fake_stream_wait_value32);
    if (q) *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorInvalidValue;
}

}  // extern "C"
