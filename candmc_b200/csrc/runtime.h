// candmc_b200 — per-process runtime state (device binding, streams, workspace, TMA descriptor encoding).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <nvtx3/nvToolsExt.h>

namespace candmc {

// NVTX range named after the reference's TAU/CTF_Timer region it stands for (d25_summa_gemm at d25_summa.cxx:122,
// d2_topo_bcast_gemm at summa.cxx:58, uni_stagger / bdr_shift / uni_shift / DGEMM at spcannon.cxx:116-341, Bcast_update at
// qr_2d.cxx:167): a profiler timeline of the GPU build lines up with the reference's own timer table.  Header-only NVTX 3;
// costs nothing when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Runtime {
  bool initialized = false;
  int device = -1;
  int num_sms = 0;
  int cc_major = 0, cc_minor = 0;
  bool force_generic = false;        // test hook: route dgemm through the CUDA-core kernel
  unsigned long long merged_chunked = 0, merged_plain = 0;   // merged last-panel launches (summa_sweep): chunk-major B / plain B
  unsigned long long transport_sends = 0;   // panel chunks shipped by copy engines (transport.h)
  unsigned long long launches = 0;   // kernels launched by this library (bench.py's gpu_launches)
  unsigned long long fused_launches = 0;   // fused GEMM + depth-sum launches enqueued (FusedEpochGuard, ipc.h)
  cudaStream_t comm_stream = nullptr;   // NCCL panel traffic
  cudaStream_t aux_stream = nullptr;    // second compute/copy stream
  void* workspace = nullptr;            // grow-only device scratch
  size_t workspace_bytes = 0;
  void* stage_pool = nullptr;           // grow-only staging buffer for chunk-wise uploads of host operands
  size_t stage_pool_bytes = 0;
  void* pfn_encode_tiled = nullptr;     // cuTensorMapEncodeTiled via cudaGetDriverEntryPoint
  int* tile_counters = nullptr;         // ring of device counters for the GEMM's dynamic tile scheduler
  unsigned tile_counter_seq = 0;
  bool fused_reduce_grids = false;      // ... also on q x q x c grids (validated so far on 1 x 1 x c only)
  bool check_peer_args = false;         // d25_summa on c > 1: verify that the depth ranks agree on the kind of C pointer (candmc_set_check_peer_args)
  bool skip_unused_uploads = true;      // host operands: a layer's rank uploads only the blocks its panels use (candmc_set_skip_unused_uploads)
  bool panel_transport = true;          // SUMMA panels and Cannon shifts by copy engines into peer windows instead of NCCL kernels (transport.h)
  bool b_first_chunk_early = false;     // pageable host B: upload the first k-chunk's rows ahead of the rest (opt-in)
  bool early_c_download = true;         // host C: finalise + download column slabs under the last multiplies (candmc_set_early_c_download)
  bool fused_reduce = true;             // depth all-reduce fused into the last GEMM's epilogue over peer memory
  bool host_gather = true;              // pinned host B blocks: k-chunks gathered straight out of host memory by the pack kernel (candmc_set_host_gather)
  bool transpose_tma = true;            // transpose through TMA loads / stores (candmc_debug_transpose_tma(0): the LDG/STG kernel)
  bool splitk = true;                   // cut small-tile-count GEMMs along k too
  int gemm_tile_n = 0;                  // CTA tile columns of the DMMA GEMM: 0 automatic, 128 (one CTA per SM) or 64 (two per SM); candmc_debug_gemm_tile
  int* sm_slots = nullptr;              // ring of per-SM arrival counters for launches with two CTAs per SM
  unsigned sm_slots_seq = 0;
  bool prefetch_c = true;               // beta != 0 GEMMs prefetch their C tiles into L2 under the last k-tiles (candmc_debug_prefetch_c)
  double* splitk_part = nullptr;        // split-K partial tiles (grow-only) and per-tile arrival counters
  size_t splitk_part_elems = 0;
  int* splitk_sem = nullptr;
  bool static_schedule = false;         // debug: round-robin tile schedule instead of the atomic counter
  bool profile = false;                 // bracket every DMMA GEMM launch with events (bench.py's roofline leg)
  int gemm_reserve_sms = 0;             // SMs a GEMM launch leaves free (set by schedules that overlap NCCL traffic)
  int bg_max_ctas = 2;                  // CTA cap of the communicators used for traffic overlapped with GEMMs (0 = no cap)
  int host_pipeline_panels = 0;         // cut of that pipeline: n > 0 uniform panels / first-panel k-chunks, -1 graduated, 0 automatic
                                        // (graduated at n, k >= 8192, else 8 uniform; host_pipeline_cut in mm_algs.cu)
  int64_t host_pipeline_min = 2048;     // smallest n for which host operands on a 1x1 grid are streamed panel-wise
  int64_t min_kchunk = 1024;            // smallest k-chunk the SUMMA pipeline cuts a panel into
  int merge_panels = 2;                 // SUMMA sweeps: k-chunks multiplied in merged launches — 2 (default): as few launches as the data's arrival allows;
                                        // 0: one launch per chunk; 1 / 3: earlier experiments (last panel only / doubling groups)
};

Runtime& runtime();

// Binds to `device` (or the current device if < 0), verifies compute capability 10.x, creates streams.
int runtime_init(int device);
// Lazily initialises on the current device; fails with ERR_NODEVICE when no usable GPU exists.
int runtime_require();
int runtime_finalize();
// Grow-only scratch; contents undefined.  Not thread safe (one rank = one host thread, as in the reference).
int workspace_get(size_t bytes, void** out);

// Scratch for split-K GEMM launches: one buffer per process.  splitk_buffers makes `stream` wait for the last split-K launch
// of any OTHER stream; splitk_release(stream) is called right after the launch that uses the scratch.
int splitk_buffers(int64_t part_elems, double** part, int** sem, cudaStream_t stream);
int splitk_release(cudaStream_t stream);

// GEMM launch profiling (candmc_profile_*): events around each TMA+DMMA launch on its own stream.
int profile_begin_launch(cudaStream_t stream, double flops);
int profile_end_launch(cudaStream_t stream);
int profile_reset();
int profile_collect(int64_t* launches, double* total_ms, double* total_flops);
int profile_timeline(double* start_ms, double* end_ms, int64_t cap, int64_t* n);

// Second grow-only device buffer: staging of host operands that are uploaded while the multiply runs.
int stage_pool_get(size_t bytes, void** out);

// Hands out a zeroed device counter (memset is enqueued on `stream`) for one GEMM launch.
int next_tile_counter(int** out, cudaStream_t stream);
// ... and a zeroed array of kSmSlotInts per-SM counters (launches with two CTAs per SM: which CTA of an SM came second)
constexpr int kSmSlotInts = 256;
int next_sm_slots(int** out, cudaStream_t stream);

// 2-D FP64 tensor map with 128 B swizzle: dim0 contiguous (dim0 x dim1 elements, leading dimension `ld`),
// box = box0 x box1 elements (box0 * 8 bytes must be <= 128).
int encode_tmap_f64(CUtensorMap* out, const double* base, int64_t dim0, int64_t dim1, int64_t ld, int box0,
                    int box1);

// The same without swizzle: the box lands (or is read) row after row, box0 x box1 doubles dense (box0 <= 256).
int encode_tmap_f64_linear(CUtensorMap* out, const double* base, int64_t dim0, int64_t dim1, int64_t ld, int box0,
                           int box1);

// The same for FP32 operands (box0 * 4 bytes <= 128; base 16-byte aligned, ld a multiple of 4).
int encode_tmap_f32(CUtensorMap* out, const float* base, int64_t dim0, int64_t dim1, int64_t ld, int box0, int box1);

// ---- kernels' host entry points (device pointers only) ---------------------------------------
// FP32 GEMM on tcgen05 (gemm_f32.cu).  Operands that are not K-major and TMA-readable in place are packed into the library
// workspace first, so calls on different streams must not overlap.
int gemm_f32(char transa, char transb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
             const float* B, int64_t ldb, float beta, float* C, int64_t ldc, cudaStream_t stream);
int gemm_f64(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
             int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
             cudaStream_t stream);
struct FusedParams;
// GEMM whose epilogue also performs the depth all-reduce over peer memory (fused == nullptr: plain GEMM)
int gemm_f64_fused(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                   int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream,
                   const FusedParams* fused);
// C = alpha * op(A) * B + beta * C with B in the SUMMA pipeline's chunk-major layout: k / b_kc consecutive chunks, chunk t the
// rows [t * b_kc, (t+1) * b_kc) of B stored as a b_kc x n matrix with leading dimension b_kc.  One launch over all of k (no
// per-chunk epilogue or tail wave); TMA path only — gemm_f64_bchunked_ok says whether the operands qualify.
bool gemm_f64_bchunked_ok(const double* A, int64_t lda, const double* B, int64_t n, int64_t k, int64_t b_kc);
int gemm_f64_bchunked(char transa, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                      const double* B, int64_t b_kc, double beta, double* C, int64_t ldc, cudaStream_t stream);
// The same for the column range [col0, col0 + ncols) of every chunk (chunks are b_kc x chunk_cols): C is the m x ncols block
// the range belongs to.  What a sweep uses to finish C column slab by column slab in launches that still cover all of k.
int gemm_f64_bchunked_cols(char transa, int64_t m, int64_t ncols, int64_t k, double alpha, const double* A, int64_t lda,
                           const double* B, int64_t b_kc, int64_t chunk_cols, int64_t col0, double beta, double* C, int64_t ldc,
                           cudaStream_t stream);
int gemm_f64_ex(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream,
                const FusedParams* fused, int64_t b_kc, int64_t b_cs);
int lda_copy_f64(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                 cudaStream_t stream);
// ... with at most max_ctas CTAs (sources that are not HBM: pinned host memory read over PCIe)
int lda_copy_f64_capped(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                        cudaStream_t stream, int max_ctas);
int lda_axpby_f64(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B, double a,
                  double b, cudaStream_t stream);
int transpose_f64(int64_t rows, int64_t cols, const double* A, int64_t lda, double* B, int64_t ldb,
                  cudaStream_t stream);
int fill_f64(double* X, int64_t count, double value, cudaStream_t stream);
// sum_k parts[k] -> out (count doubles each; parts may alias out for k==0)
int drand48_fill_f64(double* X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0, int64_t col0, int64_t n_global,
                     int which, cudaStream_t stream);
int frob_diff_f64(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t nrow, int64_t ncol,
                  double* out2 /* device: {||X-Y||_F^2, ||Y||_F^2} */, cudaStream_t stream);

}  // namespace candmc
