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

// ---- tcgen05 descriptor encoders (plain integer arithmetic: also compiled by the CPU test build, which decodes them
// independently by CUTLASS's bit-field definitions) ---------------------------------------------------------------------
#ifdef __CUDACC__
#define CANDMC_HOSTDEV __host__ __device__
#else
#define CANDMC_HOSTDEV
#endif

// Shared-memory matrix descriptor of a K-major operand tile whose rows are 128-byte lines under the 128-byte swizzle (what a
// TMA box of 32 floats x rows with CU_TENSOR_MAP_SWIZZLE_128B writes): start address >> 4 in bits [0,14), leading byte offset
// in [16,30) (unused for swizzled K-major; 1 as CuTe's make_umma_desc sets it), stride byte offset = 8 rows x 128 B = 1024 >> 4
// in [32,46), descriptor version 1 in [46,48), base offset 0, layout type SWIZZLE_128B = 2 in [61,64)
// (cute/arch/mma_sm100_desc.hpp, UMMA::SmemDescriptor).  tests/test_cutlass_descriptors.py compares this and the instruction
// descriptor below bit for bit with what CuTe itself builds for the same tile.
CANDMC_HOSTDEV inline uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

// Instruction descriptor of kind::tf32, FP32 accumulate, both operands K-major (UMMA::InstrDescriptor): c_format F32 = 1 in
// bits [4,6), a_format / b_format TF32 = 2 in [7,10) / [10,13), a_major / b_major K = 0 in [15] / [16], N >> 3 in [17,23),
// M >> 4 in [24,29).
CANDMC_HOSTDEV constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

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

// 2-D tiled TMA store shared -> global (SASS: UTMASTG) as part of the thread's current bulk group; out-of-range elements
// of the box are not written.  The source must have been made visible to the async proxy (fence_proxy_async) first.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// returns once at most N of this thread's bulk groups still have to READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// ... and once at most N have not completed entirely (global writes done)
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// Asks L2 to fetch `bytes` (a multiple of 16) from the 16-byte aligned global address `p`; no destination, no completion.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(p)), "r"(bytes) : "memory");
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// named barrier 1 among the first N threads of the CTA (the consumer warps of the GEMM kernels)
template <int N>
__device__ __forceinline__ void consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ unsigned sm_id() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void nanosleep_ns(unsigned ns) { asm volatile("nanosleep.u32 %0;" ::"r"(ns)); }

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}

// FP64 tensor-core atom: D(8x8) = A(8x4, row) * B(4x8, col) + C.  SASS: DMMA.8x8x4.
// Lane L holds A[L>>2][L&3], B[L&3][L>>2], C/D[L>>2][2*(L&3)+{0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// ---- 5th-generation tensor cores (tcgen05) for the FP32 path: accumulators in TMEM, operands in swizzled shared memory.
// Spellings as in CUTLASS's cute/arch/{mma_sm100_umma,copy_sm100,tmem_allocator_sm100}.hpp and cutlass/arch/barrier.h.

// Whole warp: allocate `ncols` (power of two >= 32) TMEM columns; the base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

// Whole warp (the one that allocated).
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One thread: D[tmem] (+)= A[smem] * B[smem]^T, M x N x 8 TF32 (SASS: UTCMMA).  `accumulate` = 0 overwrites D.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One thread: the mbarrier gets one arrival once every MMA this thread issued so far has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Whole warp w of an aligned group of four (w = warp index % 4): lane i reads TMEM lane 32*w + i, columns [col, col + 32) of
// the accumulator at `taddr` (lane in bits [16,32), column in [0,16)) into v[0..31]; returns after tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16-byte shared-memory accesses by shared-space address (the operand-splitting warps of the FP32 kernel)
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#endif  // __CUDACC__

}  // namespace candmc
