// tests/cpusim — the GEMM entry points of the simulator build (TEST INFRASTRUCTURE ONLY, see sim_device.h): same argument
// checks as gemm_f64.cu, then either plain loops (default: fast, what the schedule tests need) or the product's own kernel
// on the PTX emulation (CPUSIM_GEMM=device, and always for the fused GEMM + depth all-reduce).
// Operand ranges are checked against the simulator's allocation registry, so a wrong pointer / leading dimension / extent
// handed to the GEMM by a host schedule aborts with a message instead of silently reading a neighbour.
#include "common.cuh"
#include "ipc.h"
#include "runtime.h"
#include "sim_internal.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

extern "C" void cpusim_require_device_range(const void* p, size_t bytes, const char* what);

namespace candmc {

// the product's own gemm_f64.cu, compiled into the simulator with its entry points renamed (Makefile): TMA, mbarriers, the
// swizzled fragment loads and the DMMA atoms run on the PTX emulation of sim_exec.cxx.  Slow (every DMMA is a warp-wide fiber
// rendezvous), so it serves the tests that are about the kernel itself: CPUSIM_GEMM=device.
int gemm_f64_fused_device(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream,
                          const FusedParams* fused);

namespace {
bool use_device_kernel() {
  static const bool on = getenv("CPUSIM_GEMM") && !strcmp(getenv("CPUSIM_GEMM"), "device");
  return on;
}
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
  // the fused GEMM + depth all-reduce only exists as the kernel's epilogue over peer memory: always the emulated kernel
  if (fused != nullptr) return gemm_f64_fused_device(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream, fused);
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
  if (use_device_kernel()) return gemm_f64_fused_device(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream, nullptr);
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

}  // namespace candmc
