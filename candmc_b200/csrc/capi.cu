// candmc_b200 — extern "C" entry points for the local kernels (see include/candmc_b200.h).
// The distributed entry points live in comm.cu / mm_algs.cu.
#include "../../include/candmc_b200.h"
#include "common.cuh"
#include "runtime.h"
#include "staging.h"

using namespace candmc;

namespace {
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
inline bool tr(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }
}  // namespace

extern "C" {

int candmc_version(void) { return CANDMC_B200_VERSION; }
const char* candmc_last_error(void) { return last_error(); }
int candmc_init(int device) { return runtime_init(device); }
int candmc_finalize(void) { return runtime_finalize(); }

int candmc_device_sm_count(int* out) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(out != nullptr, "candmc_device_sm_count: null output");
  *out = runtime().num_sms;
  return OK;
}

unsigned long long candmc_launch_count(void) { return runtime().launches; }

int candmc_debug_force_generic_gemm(int on) {
  runtime().force_generic = (on != 0);
  return OK;
}

int candmc_set_fused_reduce(int on) {
  runtime().fused_reduce = (on != 0);
  runtime().fused_reduce_grids = (on >= 2);
  return OK;
}

int candmc_set_b_first_chunk_early(int on) {
  runtime().b_first_chunk_early = (on != 0);
  return OK;
}

int candmc_set_early_c_download(int on) {
  runtime().early_c_download = (on != 0);
  return OK;
}

int candmc_set_skip_unused_uploads(int on) {
  runtime().skip_unused_uploads = (on != 0);
  return OK;
}

int candmc_set_check_peer_args(int on) {
  runtime().check_peer_args = (on != 0);
  return OK;
}

int candmc_debug_splitk(int on) {
  runtime().splitk = (on != 0);
  return OK;
}

int candmc_debug_gemm_tile(int tile_n) {
  CANDMC_CHECK(tile_n == 0 || tile_n == 64 || tile_n == 128, "candmc_debug_gemm_tile: 0 (automatic), 64 or 128");
  runtime().gemm_tile_n = tile_n;
  return OK;
}

int candmc_debug_transpose_tma(int on) {
  runtime().transpose_tma = (on != 0);
  return OK;
}

int candmc_debug_prefetch_c(int on) {
  runtime().prefetch_c = (on != 0);
  return OK;
}

int candmc_debug_gemm_reserve_sms(int sms) {
  CANDMC_CHECK(sms >= 0 && sms <= 64, "candmc_debug_gemm_reserve_sms: 0..64");
  runtime().gemm_reserve_sms = sms;
  return OK;
}

int candmc_debug_static_schedule(int on) {
  runtime().static_schedule = (on != 0);
  return OK;
}

int candmc_profile_enable(int on) {
  runtime().profile = (on != 0);
  return profile_reset();
}

int candmc_profile_gemm_stats(int64_t* launches, double* total_ms, double* total_flops) {
  CANDMC_CHECK(launches && total_ms && total_flops, "candmc_profile_gemm_stats: null output");
  return profile_collect(launches, total_ms, total_flops);
}

int candmc_profile_gemm_timeline(double* start_ms, double* end_ms, int64_t cap, int64_t* n) {
  CANDMC_CHECK(start_ms && end_ms && n, "candmc_profile_gemm_timeline: null output");
  return profile_timeline(start_ms, end_ms, cap, n);
}

int candmc_set_background_ctas(int max_ctas) {
  CANDMC_CHECK(max_ctas >= 0 && max_ctas <= 64, "candmc_set_background_ctas: 0..64");
  runtime().bg_max_ctas = max_ctas;
  return OK;
}

int candmc_dgemm_chunked_b(char transa, int64_t m, int64_t n, int64_t k, int64_t kc, double alpha, const double* A, int64_t lda,
                            const double* B, double beta, double* C, int64_t ldc, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(m >= 0 && n >= 0 && k > 0 && kc > 0, "dgemm_chunked_b: bad dimensions");
  if (m == 0 || n == 0) return OK;
  CANDMC_CHECK(is_device_ptr(A) && is_device_ptr(B) && is_device_ptr(C), "dgemm_chunked_b: device pointers only");
  return gemm_f64_bchunked(transa, m, n, k, alpha, A, lda, B, kc, beta, C, ldc, as_stream(stream));
}

int candmc_dgemm(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream) {
  CANDMC_TRY(runtime_require());
  cudaStream_t st = as_stream(stream);
  if (m <= 0 || n <= 0 || (is_device_ptr(A) && is_device_ptr(B) && is_device_ptr(C)))
    return gemm_f64(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, st);
  // host operands: stage (packed) -> multiply -> copy C back; synchronous like the reference's cdgemm
  CANDMC_CHECK(m >= 0 && n >= 0 && k >= 0, "dgemm: negative dimension");
  StagedMatrix sA, sB, sC;
  CANDMC_TRY(sA.open(A, tr(transa) ? k : m, tr(transa) ? m : k, lda, true, st));
  CANDMC_TRY(sB.open(B, tr(transb) ? n : k, tr(transb) ? k : n, ldb, true, st));
  CANDMC_TRY(sC.open(C, m, n, ldc, beta != 0.0, st));
  CANDMC_TRY(gemm_f64(transa, transb, m, n, k, alpha, sA.ptr(), sA.ld(), sB.ptr(), sB.ld(), beta, sC.ptr(),
                      sC.ld(), st));
  CANDMC_TRY(sC.close_out(st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int candmc_lda_cpy(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                   void* stream) {
  CANDMC_TRY(runtime_require());
  cudaStream_t st = as_stream(stream);
  const bool dA = is_device_ptr(A), dB = is_device_ptr(B);
  if (nrow <= 0 || ncol <= 0) return OK;
  CANDMC_CHECK(lda_A >= nrow && lda_B >= nrow, "lda_cpy: leading dimension smaller than nrow");
  if (dA && dB) return lda_copy_f64(nrow, ncol, lda_A, lda_B, A, B, st);
  // any host side: a strided copy IS the staging copy — one 2-D DMA, no kernel needed
  cudaMemcpyKind kind = dA ? cudaMemcpyDeviceToHost : (dB ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost);
  CANDMC_CUDA(cudaMemcpy2DAsync(B, lda_B * 8, A, lda_A * 8, nrow * 8, ncol, kind, st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int candmc_lda_cpy_scaled(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                          double a, double b, void* stream) {
  CANDMC_TRY(runtime_require());
  cudaStream_t st = as_stream(stream);
  if (nrow <= 0 || ncol <= 0) return OK;
  if (is_device_ptr(A) && is_device_ptr(B)) return lda_axpby_f64(nrow, ncol, lda_A, lda_B, A, B, a, b, st);
  StagedMatrix sA, sB;
  CANDMC_TRY(sA.open(A, nrow, ncol, lda_A, true, st));
  CANDMC_TRY(sB.open(B, nrow, ncol, lda_B, true, st));
  CANDMC_TRY(lda_axpby_f64(nrow, ncol, sA.ld(), sB.ld(), sA.ptr(), sB.ptr(), a, b, st));
  CANDMC_TRY(sB.close_out(st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int candmc_transpose(int64_t rows, int64_t cols, const double* A, int64_t lda, double* B, int64_t ldb,
                     void* stream) {
  CANDMC_TRY(runtime_require());
  cudaStream_t st = as_stream(stream);
  if (rows <= 0 || cols <= 0) return OK;
  if (is_device_ptr(A) && is_device_ptr(B)) return transpose_f64(rows, cols, A, lda, B, ldb, st);
  StagedMatrix sA, sB;
  CANDMC_TRY(sA.open(A, rows, cols, lda, true, st));
  CANDMC_TRY(sB.open(B, cols, rows, ldb, false, st));
  CANDMC_TRY(transpose_f64(rows, cols, sA.ptr(), sA.ld(), sB.ptr(), sB.ld(), st));
  CANDMC_TRY(sB.close_out(st));
  CANDMC_CUDA(cudaStreamSynchronize(st));
  return OK;
}

int candmc_fill_drand48(double* X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0, int64_t col0,
                        int64_t n_global, int which, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(is_device_ptr(X), "candmc_fill_drand48: X must be a device pointer");
  return drand48_fill_f64(X, nrow, ncol, ld, row0, col0, n_global, which, as_stream(stream));
}

int candmc_frob_diff(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t nrow, int64_t ncol,
                     double* out_host2, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(is_device_ptr(X) && is_device_ptr(Y), "candmc_frob_diff: X and Y must be device pointers");
  cudaStream_t st = as_stream(stream);
  double* d = nullptr;
  CANDMC_CUDA(cudaMalloc(&d, 2 * sizeof(double)));
  int rc = frob_diff_f64(X, ldx, Y, ldy, nrow, ncol, d, st);
  if (rc == OK) {
    cudaError_t e = cudaMemcpyAsync(out_host2, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_last_error("candmc_frob_diff: %s", cudaGetErrorString(e));
      rc = ERR_CUDA;
    }
  }
  cudaFree(d);
  return rc;
}

}  // extern "C"
