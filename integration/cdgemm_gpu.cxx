/* integration/cdgemm_gpu.cxx — the reference's cdgemm (alg/shared/lapack.h:10-16, lapack.cxx:425-434) on a B200.
 *
 * Every local multiply of the reference — CANMM's panels, the CAQR / LU / SE trailing updates, the serial checks of its unit
 * tests — goes through this one wrapper around Fortran dgemm_.  Linked in front of alg/shared/lapack.cxx's definition
 * (-Wl,--allow-multiple-definition with this object first, or drop cdgemm from lapack.cxx), it sends products of at least
 * CANDMC_CDGEMM_MIN_FLOP flop (default 2*256^3) to candmc_dgemm; smaller ones stay on the host BLAS, where the copies would
 * cost more than the multiply.  The operands stay in host memory: candmc_dgemm stages host pointers for the call (device
 * pointers are used in place — a caller that keeps its matrices in HBM pays no copies, which is what bench.py measures).
 * Same by-value argument list, same Fortran semantics; no CUDA in this file.  One process per GPU: device = LOCAL_RANK
 * (CANDMC_SEAM_DEVICE overrides).
 */
#include <stdio.h>
#include <stdlib.h>

#include "candmc_b200.h"

#ifndef CANDMC_SEAM_DGEMM
#define CANDMC_SEAM_DGEMM dgemm_   /* the Fortran symbol of the BLAS the reference is linked with (alg/shared/lapack.cxx:36-67) */
#endif
extern "C" void CANDMC_SEAM_DGEMM(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*,
                                  const double*, const int*, const double*, double*, const int*);

namespace {
double g_min_flop = -1.0;
bool g_ready = false;
unsigned long long g_offloaded = 0, g_host = 0;

void seam_init() {
  const char* e = getenv("CANDMC_CDGEMM_MIN_FLOP");
  g_min_flop = e ? atof(e) : 2.0 * 256 * 256 * 256;
  const char* d = getenv("CANDMC_SEAM_DEVICE");
  if (!d) d = getenv("LOCAL_RANK");
  const int rc = candmc_init(d ? atoi(d) : 0);
  if (rc != CANDMC_OK) {
    fprintf(stderr, "cdgemm_gpu: candmc_init failed: %s\n", candmc_last_error());
    abort();
  }
  g_ready = true;
  if (getenv("CANDMC_SEAM_VERBOSE"))
    atexit([] { fprintf(stderr, "cdgemm_gpu: %llu products in the library, %llu on the host BLAS\n", g_offloaded, g_host); });
}
}  // namespace

void cdgemm(char transa, char transb, int m, int n, int k, const double a, const double* A, int lda, const double* B, int ldb, double b,
            double* C, int ldc) {
  if (!g_ready) seam_init();
  if (2.0 * m * n * k < g_min_flop || m <= 0 || n <= 0) {
    ++g_host;
    CANDMC_SEAM_DGEMM(&transa, &transb, &m, &n, &k, &a, A, &lda, B, &ldb, &b, C, &ldc);
    return;
  }
  ++g_offloaded;
  const int rc = candmc_dgemm(transa, transb, m, n, k, a, A, lda, B, ldb, b, C, ldc, NULL);
  if (rc != CANDMC_OK) {
    fprintf(stderr, "cdgemm_gpu: candmc_dgemm(%c,%c,%d,%d,%d) failed: %s\n", transa, transb, m, n, k, candmc_last_error());
    abort();
  }
}
