// tests/cpusim — stand-ins for the two translation units of the product that cannot be simulated (TEST INFRASTRUCTURE
// ONLY, see sim_device.h):
//   gemm_f64.cu  (TMA + DMMA kernel, inline PTX)  ->  the same host entry points, same argument checks, plain loops;
//   ipc.cu       (CUDA IPC peer windows + the fused depth all-reduce)  ->  "not available", which is the product's own
//                documented condition for falling back to ncclAllReduce.
// Operand ranges are checked against the simulator's allocation registry, so a wrong pointer / leading dimension / extent
// handed to the GEMM by a host schedule aborts with a message instead of silently reading a neighbour.
#include "common.cuh"
#include "ipc.h"
#include "runtime.h"
#include "sim_internal.h"

#include <algorithm>
#include <vector>

extern "C" void cpusim_require_device_range(const void* p, size_t bytes, const char* what);

namespace candmc {

namespace {
bool is_trans(char c) { return c == 'T' || c == 't' || c == 'C' || c == 'c'; }
bool is_notrans(char c) { return c == 'N' || c == 'n'; }
size_t span(int64_t rows, int64_t cols, int64_t ld) { return rows && cols ? sizeof(double) * ((cols - 1) * ld + rows) : 0; }
}  // namespace

int gemm_f64(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
             const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream) {
  return gemm_f64_fused(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream, nullptr);
}

int gemm_f64_fused(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream,
                   const FusedParams* fused) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(fused == nullptr, "cpusim: the fused GEMM + depth all-reduce cannot be simulated");
  CANDMC_CHECK(is_trans(transa) || is_notrans(transa), "dgemm: bad transa '%c'", transa);
  CANDMC_CHECK(is_trans(transb) || is_notrans(transb), "dgemm: bad transb '%c'", transb);
  const bool tA = is_trans(transa), tB = is_trans(transb);
  CANDMC_CHECK(m >= 0 && n >= 0 && k >= 0, "dgemm: negative dimension m=%lld n=%lld k=%lld", (long long)m, (long long)n,
               (long long)k);
  const int64_t rowsA = tA ? k : m, colsA = tA ? m : k, rowsB = tB ? n : k, colsB = tB ? k : n;
  CANDMC_CHECK(lda >= (rowsA > 1 ? rowsA : 1), "dgemm: lda=%lld < %lld", (long long)lda, (long long)rowsA);
  CANDMC_CHECK(ldb >= (rowsB > 1 ? rowsB : 1), "dgemm: ldb=%lld < %lld", (long long)ldb, (long long)rowsB);
  CANDMC_CHECK(ldc >= (m > 1 ? m : 1), "dgemm: ldc=%lld < %lld", (long long)ldc, (long long)m);
  if (m == 0 || n == 0) return OK;
  cpusim_require_device_range(C, span(m, n, ldc), "dgemm C");
  if (k > 0 && alpha != 0.0) {
    cpusim_require_device_range(A, span(rowsA, colsA, lda), "dgemm A");
    cpusim_require_device_range(B, span(rowsB, colsB, ldb), "dgemm B");
  }
  if ((k == 0 || alpha == 0.0) && beta == 1.0) return OK;
  cpusim::stream_submit(stream, [=]() {
  // op(A) packed m x k, then column-by-column axpy (vectorisable inner loop); summation order over k is 0..k-1
    std::vector<double> Ap, acc((size_t)m);
    const double* Aop = A;
    int64_t lda_op = lda;
    if (tA && k > 0 && alpha != 0.0) {
      Ap.resize((size_t)m * k);
      for (int64_t p = 0; p < k; ++p)
        for (int64_t i = 0; i < m; ++i) Ap[i + p * m] = A[p + i * lda];
      Aop = Ap.data();
      lda_op = m;
    }
    for (int64_t j = 0; j < n; ++j) {
      std::fill(acc.begin(), acc.end(), 0.0);
      if (alpha != 0.0)
        for (int64_t p = 0; p < k; ++p) {
          const double bv = tB ? B[j + p * ldb] : B[p + j * ldb];
          const double* a = Aop + p * lda_op;
          for (int64_t i = 0; i < m; ++i) acc[i] += a[i] * bv;
        }
      double* c = C + j * ldc;
      for (int64_t i = 0; i < m; ++i)
        c[i] = beta == 0.0 ? alpha * acc[i] : alpha * acc[i] + beta * c[i];  // beta == 0 never reads C, as in the kernel
    }
  });
  runtime().launches++;
  return OK;
}

// ---- ipc.cu ------------------------------------------------------------------------------------------------------------
int window_create(candmc_comm*, size_t, PeerWindow** out) {
  *out = nullptr;
  set_last_error("cpusim: CUDA IPC is not simulated");
  return ERR_CUDA;
}
void window_destroy(PeerWindow*) {}
int fused_ctx_get(candmc_comm*, int64_t, FusedCtx** out) {
  *out = nullptr;  // "the fused path cannot be used (the caller then uses ncclAllReduce)", ipc.h
  return OK;
}
void fused_params_next(FusedCtx*, int, FusedParams*) {}
int fused_finish(FusedCtx*, int, const FusedParams&, double*, int64_t, cudaStream_t) {
  set_last_error("cpusim: fused_finish without a fused context");
  return ERR_INVALID;
}

}  // namespace candmc
