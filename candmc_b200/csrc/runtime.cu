// candmc_b200 — runtime: device binding, error string, workspace, TMA descriptor encoding.
#include "runtime.h"
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace candmc {

namespace {
thread_local char g_err[1024] = "";
Runtime g_rt;
constexpr unsigned kTileCounters = 1024;
// The split-K scratch (partial tiles + per-tile arrival counters) is one buffer per process.  Launches on one stream are
// ordered by the stream; a launch on ANOTHER stream (the LU seam's GEMM stream next to a caller's stream, two caller streams)
// first waits for the last launch that used the scratch, so two split-K GEMMs never mix their partials or counters.
cudaEvent_t g_splitk_last = nullptr;
cudaStream_t g_splitk_stream = nullptr;
bool g_splitk_used = false;
}  // namespace

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

Runtime& runtime() { return g_rt; }

int runtime_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_last_error("candmc_b200: no CUDA device visible (%s); this library has no CPU path",
                   e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return ERR_NODEVICE;
  }
  if (device < 0) CANDMC_CUDA(cudaGetDevice(&device));
  CANDMC_CHECK(device < count, "candmc_init: device %d out of range (%d visible)", device, count);
  if (g_rt.initialized && g_rt.device == device) return OK;
  CANDMC_CHECK(!g_rt.initialized, "candmc_init: already bound to device %d (one GPU per process)", g_rt.device);
  CANDMC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CANDMC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("candmc_b200: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major,
                   prop.minor);
    return ERR_NODEVICE;
  }
  g_rt.device = device;
  g_rt.num_sms = prop.multiProcessorCount;
  g_rt.cc_major = prop.major;
  g_rt.cc_minor = prop.minor;
  // The communication stream gets the highest priority: when a persistent GEMM ends and both the next GEMM's CTAs and an
  // NCCL kernel's (a panel broadcast, a shift, a slab of the depth sum) are pending, the few communication CTAs are placed
  // first and the GEMM — whose dynamic tile scheduler does not care how many of its CTAs are resident — takes the rest.
  int prio_lo = 0, prio_hi = 0;
  CANDMC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CANDMC_CUDA(cudaStreamCreateWithPriority(&g_rt.comm_stream, cudaStreamNonBlocking, prio_hi));
  CANDMC_CUDA(cudaStreamCreateWithFlags(&g_rt.aux_stream, cudaStreamNonBlocking));
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  CANDMC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  CANDMC_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess,
               "candmc_init: driver does not export cuTensorMapEncodeTiled");
  g_rt.pfn_encode_tiled = fn;
  CANDMC_CUDA(cudaMalloc(&g_rt.tile_counters, sizeof(int) * kTileCounters));
  CANDMC_CUDA(cudaMemset(g_rt.tile_counters, 0, sizeof(int) * kTileCounters));
  // Tuning switches for callers that cannot call the candmc_set_* functions (the reference's unmodified mains running as
  // drop-ins): same meaning as the setters, read once when the process binds to its GPU.
  auto env_int = [](const char* name, long long* out) {
    const char* e = getenv(name);
    if (e == nullptr || *e == 0) return false;
    *out = atoll(e);
    return true;
  };
  long long v;
  if (env_int("CANDMC_PANEL_TRANSPORT", &v)) g_rt.panel_transport = (v != 0);
  if (env_int("CANDMC_FUSED_REDUCE", &v)) {
    g_rt.fused_reduce = (v != 0);
    g_rt.fused_reduce_grids = (v >= 2);
  }
  if (env_int("CANDMC_BG_CTAS", &v) && v >= 0 && v <= 64) g_rt.bg_max_ctas = (int)v;
  if (env_int("CANDMC_MIN_KCHUNK", &v) && v >= 2) g_rt.min_kchunk = v;
  if (env_int("CANDMC_MERGE_LAST_PANEL", &v)) g_rt.merge_panels = (v != 0) ? 1 : 0;
  if (env_int("CANDMC_MERGE_PANELS", &v) && v >= 0 && v <= 3) g_rt.merge_panels = (int)v;
  if (env_int("CANDMC_EARLY_C_DOWNLOAD", &v)) g_rt.early_c_download = (v != 0);
  if (env_int("CANDMC_SKIP_UNUSED_UPLOADS", &v)) g_rt.skip_unused_uploads = (v != 0);
  if (env_int("CANDMC_CHECK_PEER_ARGS", &v)) g_rt.check_peer_args = (v != 0);
  g_rt.initialized = true;
  return OK;
}

int runtime_require() {
  if (g_rt.initialized) return OK;
  return runtime_init(-1);
}

int runtime_finalize() {
  if (!g_rt.initialized) return OK;
  if (g_rt.workspace) cudaFree(g_rt.workspace);
  if (g_rt.stage_pool) cudaFree(g_rt.stage_pool);
  if (g_rt.tile_counters) cudaFree(g_rt.tile_counters);
  if (g_rt.sm_slots) cudaFree(g_rt.sm_slots);
  if (g_rt.splitk_part) cudaFree(g_rt.splitk_part);
  if (g_rt.splitk_sem) cudaFree(g_rt.splitk_sem);
  if (g_splitk_last) cudaEventDestroy(g_splitk_last);
  g_splitk_last = nullptr;
  g_splitk_stream = nullptr;
  g_splitk_used = false;
  if (g_rt.comm_stream) cudaStreamDestroy(g_rt.comm_stream);
  if (g_rt.aux_stream) cudaStreamDestroy(g_rt.aux_stream);
  g_rt = Runtime();
  return OK;
}

int workspace_get(size_t bytes, void** out) {
  CANDMC_TRY(runtime_require());
  if (bytes > g_rt.workspace_bytes) {
    if (g_rt.workspace) {
      CANDMC_CUDA(cudaDeviceSynchronize());
      CANDMC_CUDA(cudaFree(g_rt.workspace));
      g_rt.workspace = nullptr;
      g_rt.workspace_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&g_rt.workspace, bytes);
    if (e != cudaSuccess) {
      set_last_error("workspace: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      return ERR_NOMEM;
    }
    g_rt.workspace_bytes = bytes;
  }
  *out = g_rt.workspace;
  return OK;
}

int stage_pool_get(size_t bytes, void** out) {
  CANDMC_TRY(runtime_require());
  if (bytes > g_rt.stage_pool_bytes) {
    if (g_rt.stage_pool) {
      CANDMC_CUDA(cudaDeviceSynchronize());
      CANDMC_CUDA(cudaFree(g_rt.stage_pool));
      g_rt.stage_pool = nullptr;
      g_rt.stage_pool_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&g_rt.stage_pool, bytes);
    if (e != cudaSuccess) {
      set_last_error("staging pool: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      return ERR_NOMEM;
    }
    g_rt.stage_pool_bytes = bytes;
  }
  *out = g_rt.stage_pool;
  return OK;
}

int splitk_release(cudaStream_t stream) {
  if (g_splitk_last == nullptr) CANDMC_CUDA(cudaEventCreateWithFlags(&g_splitk_last, cudaEventDisableTiming));
  CANDMC_CUDA(cudaEventRecord(g_splitk_last, stream));
  g_splitk_stream = stream;
  g_splitk_used = true;
  return OK;
}

int splitk_buffers(int64_t part_elems, double** part, int** sem, cudaStream_t stream) {
  if (g_splitk_used && g_splitk_stream != stream) CANDMC_CUDA(cudaStreamWaitEvent(stream, g_splitk_last, 0));
  if (g_rt.splitk_sem == nullptr) {
    CANDMC_CUDA(cudaMalloc(&g_rt.splitk_sem, sizeof(int) * 4096));
    CANDMC_CUDA(cudaMemset(g_rt.splitk_sem, 0, sizeof(int) * 4096));
  }
  if ((size_t)part_elems > g_rt.splitk_part_elems) {
    if (g_rt.splitk_part) {
      CANDMC_CUDA(cudaDeviceSynchronize());
      CANDMC_CUDA(cudaFree(g_rt.splitk_part));
      g_rt.splitk_part = nullptr;
      g_rt.splitk_part_elems = 0;
    }
    cudaError_t e = cudaMalloc(&g_rt.splitk_part, sizeof(double) * part_elems);
    if (e != cudaSuccess) {
      set_last_error("split-K scratch: cudaMalloc(%lld doubles) failed: %s", (long long)part_elems, cudaGetErrorString(e));
      return ERR_NOMEM;
    }
    g_rt.splitk_part_elems = (size_t)part_elems;
  }
  *part = g_rt.splitk_part;
  *sem = g_rt.splitk_sem;
  return OK;
}

int next_sm_slots(int** out, cudaStream_t stream) {
  constexpr unsigned kRing = 64;
  if (g_rt.sm_slots == nullptr) CANDMC_CUDA(cudaMalloc(&g_rt.sm_slots, sizeof(int) * kSmSlotInts * kRing));
  int* c = g_rt.sm_slots + static_cast<size_t>(g_rt.sm_slots_seq++ % kRing) * kSmSlotInts;
  CANDMC_CUDA(cudaMemsetAsync(c, 0, sizeof(int) * kSmSlotInts, stream));
  *out = c;
  return OK;
}

int next_tile_counter(int** out, cudaStream_t stream) {
  int* c = g_rt.tile_counters + (g_rt.tile_counter_seq++ % kTileCounters);
  CANDMC_CUDA(cudaMemsetAsync(c, 0, sizeof(int), stream));
  *out = c;
  return OK;
}

namespace {
struct ProfRec {
  cudaEvent_t e0, e1;
  double flops;
};
std::vector<ProfRec> g_prof;
}  // namespace

int profile_begin_launch(cudaStream_t stream, double flops) {
  ProfRec r;
  CANDMC_CUDA(cudaEventCreate(&r.e0));
  CANDMC_CUDA(cudaEventCreate(&r.e1));
  r.flops = flops;
  CANDMC_CUDA(cudaEventRecord(r.e0, stream));
  g_prof.push_back(r);
  return OK;
}
int profile_end_launch(cudaStream_t stream) {
  CANDMC_CUDA(cudaEventRecord(g_prof.back().e1, stream));
  return OK;
}
int profile_reset() {
  for (ProfRec& r : g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return OK;
}
int profile_collect(int64_t* launches, double* total_ms, double* total_flops) {
  CANDMC_CUDA(cudaDeviceSynchronize());
  double ms = 0, fl = 0;
  int64_t cnt = 0;
  for (ProfRec& r : g_prof) {
    if (r.flops == 0.0) continue;   // a mark (start of a distributed call), not a launch
    float t = 0;
    CANDMC_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    fl += r.flops;
    ++cnt;
  }
  *launches = cnt;
  *total_ms = ms;
  *total_flops = fl;
  return OK;
}

int profile_timeline(double* start_ms, double* end_ms, int64_t cap, int64_t* n) {
  CANDMC_CUDA(cudaDeviceSynchronize());
  int64_t cnt = 0;
  for (ProfRec& r : g_prof) {
    if (cnt >= cap) break;
    float a = 0, b = 0;
    CANDMC_CUDA(cudaEventElapsedTime(&a, g_prof.front().e0, r.e0));
    CANDMC_CUDA(cudaEventElapsedTime(&b, g_prof.front().e0, r.e1));
    start_ms[cnt] = a;
    end_ms[cnt] = b;
    ++cnt;
  }
  *n = cnt;
  return OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

namespace {
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base, int64_t dim0, int64_t dim1,
                int64_t ld, int box0, int box1, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(dim0 > 0 && dim1 > 0, "tensor map: empty operand");
  cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(g_rt.pfn_encode_tiled)(
      out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d) dims=%lldx%lld ld=%lld box=%dx%d ptr=%p elem=%d", (int)r,
                   (long long)dim0, (long long)dim1, (long long)ld, box0, box1, base, elem_bytes);
    return ERR_CUDA;
  }
  return OK;
}
}  // namespace

int encode_tmap_f64(CUtensorMap* out, const double* base, int64_t dim0, int64_t dim1, int64_t ld, int box0,
                    int box1) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, base, dim0, dim1, ld, box0, box1);
}

int encode_tmap_f64_linear(CUtensorMap* out, const double* base, int64_t dim0, int64_t dim1, int64_t ld, int box0,
                           int box1) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, base, dim0, dim1, ld, box0, box1, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int encode_tmap_f32(CUtensorMap* out, const float* base, int64_t dim0, int64_t dim1, int64_t ld, int box0, int box1) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dim0, dim1, ld, box0, box1);
}

}  // namespace candmc
