// candmc_b200 — shared device/host helpers for the sm_100a kernels.
// Everything here is private to candmc_b200/csrc; the public surface is include/candmc_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace candmc {

// ---- status codes (mirrored in include/candmc_b200.h) ----------------------------------------
enum : int {
  OK = 0,
  ERR_INVALID = 1,   // bad argument (the reference would `assert`/ABORT, util.h:127-138)
  ERR_CUDA = 2,      // CUDA runtime / driver error
  ERR_NCCL = 3,      // NCCL error
  ERR_NOMEM = 4,     // workspace allocation failed
  ERR_NODEVICE = 5,  // no sm_100 device visible: the product path has no CPU fallback
};

void set_last_error(const char* fmt, ...);
const char* last_error();

#define CANDMC_CUDA(call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::candmc::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                               cudaGetErrorString(e__));                                    \
      return ::candmc::ERR_CUDA;                                                            \
    }                                                                                       \
  } while (0)

#define CANDMC_CHECK(cond, ...)                                                             \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::candmc::set_last_error(__VA_ARGS__);                                                \
      return ::candmc::ERR_INVALID;                                                         \
    }                                                                                       \
  } while (0)

#define CANDMC_TRY(expr)                                                                    \
  do {                                                                                      \
    int s__ = (expr);                                                                       \
    if (s__ != ::candmc::OK) return s__;                                                    \
  } while (0)

// ---- device PTX helpers ----------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 2-D tiled TMA load global -> shared, completion on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// FP64 tensor-core atom: D(8x8) = A(8x4, row) * B(4x8, col) + C.  SASS: DMMA.8x8x4.
// Lane L holds A[L>>2][L&3], B[L&3][L>>2], C/D[L>>2][2*(L&3)+{0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

#endif  // __CUDACC__

}  // namespace candmc
