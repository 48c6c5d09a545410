// tests/cpusim — CPU functional simulator of the CUDA execution model, TEST INFRASTRUCTURE ONLY.
//
// Force-included (g++ -include) in front of every translated candmc_b200/csrc/*.cu so that the product's host schedules
// and its simple (non-TMA, non-DMMA) kernels can be EXECUTED in the CPU test suite: a kernel launch runs every thread of
// every block as a fiber (ucontext), __syncthreads / warp shuffles are fiber barriers, streams are synchronous.  The hot
// GEMM kernel (TMA + DMMA, inline PTX) cannot be simulated and is replaced by plain loops (sim_gemm.cxx); what this buys is
// a CPU run of everything AROUND it — index arithmetic of the pack kernels, staging, stream/event ordering logic, NCCL call
// sequences on real multi-process grids — before GPU minutes are spent.  The product library never loads or links any of
// this (tests/test_boundary.py checks), and nothing here is a fallback: libcandmc_b200.so still fails without a B200.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <time.h>

#include <functional>

namespace cpusim {

struct Idx {
  unsigned x = 0, y = 0, z = 0;
};
struct ThreadState {
  Idx thread, block, bdim, gdim;
};
ThreadState& ts();        // state of the fiber that is running
void* dyn_smem();         // dynamic shared memory of the running block
void sync_threads();      // block barrier
void sync_warp_exchange(const void* in, void* out, int src_lane, size_t bytes);  // one warp shuffle
void launch(cudaStream_t stream, dim3 grid, dim3 block, size_t smem, std::function<void()> body);
// ---- the PTX the hot GEMM kernel is written in (common.cuh), emulated ------------------------------------------------
uint32_t smem_handle(const void* p);   // "shared-space address": offset inside the block's dynamic shared memory (+4096)
void* smem_ptr(uint32_t handle);
void mbar_init(uint32_t bar, uint32_t count);
void mbar_arrive(uint32_t bar);
void mbar_expect_tx(uint32_t bar, uint32_t bytes);   // arrive.expect_tx
bool mbar_try_wait(uint32_t bar, uint32_t parity);   // yields to the other fibers when the phase is not complete yet
void tma_load_2d(uint32_t dst, const void* tensor_map, uint32_t bar, int c0, int c1);  // tiled, 128 B swizzle, OOB zero fill
void tma_store_2d(const void* tensor_map, uint32_t src, int c0, int c1);   // tiled store, executed when it is issued
void warp_allgather16(const void* in16, void* out32x16);  // every lane contributes 16 bytes, all get all
void named_barrier(int id, int count);                 // bar.sync id, count
// tcgen05 (gemm_f32.cu): tensor memory, the TF32 UMMA executed when it is issued, commit = immediate mbarrier arrival
void tmem_alloc(uint32_t slot, uint32_t ncols);
void tmem_dealloc(uint32_t taddr, uint32_t ncols);
void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate);
void umma_commit(uint32_t bar);
void tmem_ld_32x32(uint32_t taddr, uint32_t* v);
void smem_rw16(uint32_t handle, void* data, bool write);
void dmma_defer(double* d0, double* d1, double a, double b);   // log one DMMA.8x8x4 of the calling lane
void dmma_flush();                                     // warp rendezvous: apply every logged DMMA

}  // namespace cpusim

#define threadIdx (::cpusim::ts().thread)
#define blockIdx (::cpusim::ts().block)
#define blockDim (::cpusim::ts().bdim)
#define gridDim (::cpusim::ts().gdim)
#define warpSize 32

#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#undef __forceinline__
#define __forceinline__ inline
// a __shared__ array is shared by the fibers of a block, which run one block at a time on one OS thread: `static` is it
#undef __shared__
#define __shared__ static

// kernel arguments are evaluated and copied when the launch is issued (as CUDA does), the body runs when the stream gets to it
namespace cpusim {
template <class K, class... Args>
static inline void launch_args(cudaStream_t stream, dim3 grid, dim3 block, size_t smem, K kern, Args... args) {
  launch(stream, grid, block, smem, [=]() { kern(args...); });
}
}  // namespace cpusim
#define CPUSIM_LAUNCH(kern, grid, block, smem, stream, ...) \
  ::cpusim::launch_args((cudaStream_t)(stream), dim3(grid), dim3(block), (size_t)(smem), kern, ##__VA_ARGS__)

static inline void __syncthreads() { ::cpusim::sync_threads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { ::cpusim::dmma_flush(); }

template <class T>
static inline T __ldg(const T* p) {
  return *p;
}
template <class T>
static inline T __ldcg(const T* p) {
  return *p;
}
template <class T>
static inline void __stcg(T* p, T v) {
  *p = v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  T out;
  ::cpusim::sync_warp_exchange(&v, &out, (int)((::cpusim::ts().thread.x & 31u) ^ (unsigned)lane_mask), sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src_lane) {
  T out;
  ::cpusim::sync_warp_exchange(&v, &out, src_lane & 31, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  T out;
  int lane = (int)(::cpusim::ts().thread.x & 31u);
  ::cpusim::sync_warp_exchange(&v, &out, lane + (int)delta < 32 ? lane + (int)delta : lane, sizeof(T));
  return out;
}
// one OS thread per process executes all fibers: plain read-modify-write is atomic here
template <class T>
static inline T atomicAdd(T* p, T v) {
  T old = *p;
  *p = old + v;
  return old;
}
static inline unsigned atomicInc(unsigned* p, unsigned lim) {
  unsigned old = *p;
  *p = old >= lim ? 0 : old + 1;
  return old;
}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
template <class T>
static inline T min(T a, T b) {
  return a < b ? a : b;
}
template <class T>
static inline T max(T a, T b) {
  return a > b ? a : b;
}
static inline unsigned __float_as_uint(float f) {
  unsigned v;
  memcpy(&v, &f, 4);
  return v;
}
static inline float __uint_as_float(unsigned v) {
  float f;
  memcpy(&f, &v, 4);
  return f;
}
static inline long long __double_as_longlong(double d) {
  long long v;
  memcpy(&v, &d, 8);
  return v;
}
static inline double __longlong_as_double(long long v) {
  double d;
  memcpy(&d, &v, 8);
  return d;
}
// the typed overload nvcc's cuda_runtime.h provides for kernels
template <class T>
static inline cudaError_t cudaFuncSetAttribute(T* f, cudaFuncAttribute a, int v) {
  return ::cudaFuncSetAttribute(reinterpret_cast<const void*>(f), a, v);
}

// common.cuh's device helpers (there under #ifdef __CUDACC__), same names and signatures, on the emulation above
namespace candmc {
static inline uint32_t smem_u32(const void* p) { return ::cpusim::smem_handle(p); }
static inline void mbar_init(uint64_t* bar, uint32_t count) { ::cpusim::mbar_init(smem_u32(bar), count); }
static inline void fence_mbar_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { ::cpusim::mbar_expect_tx(smem_u32(bar), bytes); }
static inline void mbar_arrive(uint64_t* bar) { ::cpusim::mbar_arrive(smem_u32(bar)); }
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return ::cpusim::mbar_try_wait(smem_u32(bar), parity); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
static inline void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  ::cpusim::tma_load_2d(smem_u32(dst), tmap, smem_u32(bar), c0, c1);
}
static inline void tma_store_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
  ::cpusim::tma_store_2d(tmap, smem_u32(src), c0, c1);
}
static inline void tma_store_commit() {}
template <int N>
static inline void tma_store_wait_read() {}
template <int N>
static inline void tma_store_wait_all() {}
static inline void tma_prefetch_desc(const CUtensorMap*) {}
static inline void l2_prefetch_bulk(const void*, uint32_t) {}   // a cache hint: nothing to emulate
static inline double lds_f64(uint32_t addr) { return *static_cast<const double*>(::cpusim::smem_ptr(addr)); }
template <int N>
static inline void consumer_barrier() { ::cpusim::named_barrier(1, N); }
static inline unsigned sm_id() { return ::cpusim::ts().block.x % 148u; }   // blocks run one after the other: any id will do
static inline unsigned long long global_timer_ns() {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (unsigned long long)t.tv_sec * 1000000000ull + (unsigned long long)t.tv_nsec;
}
static inline void nanosleep_ns(unsigned) {}
static inline double2 lds_f64x2(uint32_t addr) {
  double2 v;
  memcpy(&v, ::cpusim::smem_ptr(addr), 16);
  return v;
}
static inline void tmem_alloc(uint32_t* slot, uint32_t ncols) { ::cpusim::tmem_alloc(smem_u32(slot), ncols); }
static inline void tmem_dealloc(uint32_t taddr, uint32_t ncols) { ::cpusim::tmem_dealloc(taddr, ncols); }
static inline void tcgen05_fence_before() {}
static inline void tcgen05_fence_after() {}
static inline void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  ::cpusim::umma_tf32(tmem_d, desc_a, desc_b, idesc, accumulate);
}
static inline void umma_commit(uint64_t* bar) { ::cpusim::umma_commit(smem_u32(bar)); }
static inline void tmem_ld_32x32(uint32_t taddr, uint32_t* v) { ::cpusim::tmem_ld_32x32(taddr, v); }
static inline float4 lds_f32x4(uint32_t addr) {
  float4 v;
  ::cpusim::smem_rw16(addr, &v, false);
  return v;
}
static inline void sts_f32x4(uint32_t addr, float4 v) { ::cpusim::smem_rw16(addr, &v, true); }
// mma.sync.aligned.m8n8k4.row.col.f64: lane L holds A[L>>2][L&3], B[L&3][L>>2], C/D[L>>2][2*(L&3)+{0,1}].
// A rendezvous of 32 fibers per DMMA would dominate the run time, so the emulation defers: a call only logs its operands
// and where its accumulators live; the warp meets once per batch — when an accumulator comes round again (the next k-step)
// or at __syncwarp(), which the kernel reaches before anything reads the accumulators — and every lane then applies the
// logged DMMAs in call order.
static inline void dmma884(double& d0, double& d1, double a, double b) { ::cpusim::dmma_defer(&d0, &d1, a, b); }
}  // namespace candmc
