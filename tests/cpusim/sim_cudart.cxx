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

#include <map>

namespace {

constexpr size_t kGuard = 64;
constexpr unsigned char kCanary = 0xA5;

struct Alloc {
  size_t bytes;
  bool device;  // false: pinned host
};
std::map<uintptr_t, Alloc> g_allocs;   // user base -> allocation
cudaError_t g_last = cudaSuccess;
int g_device = 0;

struct SimStream {
  int id;
};
struct SimEvent {
  double ms;
  bool recorded;
};
int g_next_stream = 1;

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

void* sim_alloc(size_t bytes, bool device) {
  void* raw = nullptr;
  size_t total = bytes + 2 * kGuard + 256;
  if (posix_memalign(&raw, 256, total) != 0) return nullptr;
  // keep the user pointer 256-byte aligned like cudaMalloc: guard sits in the 256 bytes in front of it
  unsigned char* user = (unsigned char*)raw + 256;
  memset(user - kGuard, kCanary, kGuard);
  memset(user + bytes, kCanary, kGuard);
  if (device) {
    // signalling-NaN-ish garbage: 0x7FF4A5A5A5A5A5A5
    uint64_t pat = 0x7FF4A5A5A5A5A5A5ull;
    for (size_t i = 0; i + 8 <= bytes; i += 8) memcpy(user + i, &pat, 8);
    for (size_t i = bytes & ~(size_t)7; i < bytes; ++i) user[i] = 0xA7;
  }
  g_allocs[(uintptr_t)user] = Alloc{bytes, device};
  return user;
}

cudaError_t sim_free(void* p, bool device) {
  if (!p) return cudaSuccess;
  auto it = g_allocs.find((uintptr_t)p);
  if (it == g_allocs.end() || it->second.device != device) {
    fprintf(stderr, "cpusim: %s of a pointer that is not a live %s allocation: %p\n", device ? "cudaFree" : "cudaFreeHost",
            device ? "device" : "pinned", p);
    abort();
  }
  check_canaries(device ? "cudaFree" : "cudaFreeHost");
  memset(p, 0xDD, it->second.bytes);  // use-after-free shows up as garbage
  g_allocs.erase(it);
  free((unsigned char*)p - 256);
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

struct AtExit {
  ~AtExit() { check_canaries("process exit"); }
} g_at_exit;

CUresult fake_encode_tiled(CUtensorMap* out, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  memset(out, 0, sizeof(*out));
  return CUDA_SUCCESS;
}

}  // namespace

extern "C" {

// helpers for the Python side of the tests (simtorch.py)
void* cpusim_malloc(size_t bytes) { return sim_alloc(bytes, true); }
void cpusim_check(void) { check_canaries("cpusim_check"); }
void cpusim_require_device_range(const void* p, size_t bytes, const char* what) { check_range(p, bytes, true, what); }
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
  check_canaries("cudaDeviceSynchronize");
  return cudaSuccess;
}

cudaError_t cudaMalloc(void** p, size_t bytes) {
  *p = sim_alloc(bytes, true);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) { return sim_free(p, true); }
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
  attr->devicePointer = a && a->device ? const_cast<void*>(p) : nullptr;
  attr->hostPointer = a && a->device ? nullptr : const_cast<void*>(p);
  return cudaSuccess;
}

cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t) {
  check_copy(dst, src, bytes, bytes, kind, "cudaMemcpyAsync");
  memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  check_copy(dst, src, bytes, bytes, kind, "cudaMemcpy");
  memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind kind, cudaStream_t) {
  if (width == 0 || height == 0) return cudaSuccess;
  if (width > dpitch || width > spitch) {
    fprintf(stderr, "cpusim: cudaMemcpy2DAsync: width %zu exceeds a pitch (%zu, %zu)\n", width, dpitch, spitch);
    g_last = cudaErrorInvalidPitchValue;
    return cudaErrorInvalidPitchValue;
  }
  check_copy(dst, src, dpitch * (height - 1) + width, spitch * (height - 1) + width, kind, "cudaMemcpy2DAsync");
  for (size_t r = 0; r < height; ++r) memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) {
  check_range(p, bytes, true, "cudaMemsetAsync");
  memset(p, v, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) { return cudaMemsetAsync(p, v, bytes, nullptr); }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  SimStream* st = new SimStream{g_next_stream++};
  *s = reinterpret_cast<cudaStream_t>(st);
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete reinterpret_cast<SimStream*>(s);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) {
  check_canaries("cudaStreamSynchronize");
  return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t e, unsigned) {
  if (!e) {
    fprintf(stderr, "cpusim: cudaStreamWaitEvent on a null event\n");
    abort();
  }
  return cudaSuccess;  // synchronous streams: whatever the event covers has already happened
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = reinterpret_cast<cudaEvent_t>(new SimEvent{0.0, false});
  return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete reinterpret_cast<SimEvent*>(e);
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  SimEvent* ev = reinterpret_cast<SimEvent*>(e);
  ev->ms = now_ms();
  ev->recorded = true;
  return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  SimEvent *ea = reinterpret_cast<SimEvent*>(a), *eb = reinterpret_cast<SimEvent*>(b);
  if (!ea->recorded || !eb->recorded) return cudaErrorInvalidResourceHandle;
  *ms = (float)(eb->ms - ea->ms);
  return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }

cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  if (strcmp(symbol, "cuTensorMapEncodeTiled") == 0) {
    *fn = reinterpret_cast<void*>(&fake_encode_tiled);
    if (q) *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorInvalidValue;
}

}  // extern "C"
