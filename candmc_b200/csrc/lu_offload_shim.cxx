// candmc_b200 — the reference's LU offload entry points (alg/LU/lu_offload.h:19-101) as thin C++ wrappers over the
// C ABI (candmc_off_*).  Built into its own small library, libcandmc_lu_offload.so, so that a program which brings
// its own MPI and BLAS (the reference's LU drivers do) links nothing but these sixteen symbols.
#include <stdio.h>
#include <unistd.h>

#include "../../include/candmc/lu_offload.h"
#include "../../include/candmc_b200.h"

namespace {
void check(int status, const char* what) {
  if (status == CANDMC_OK) return;
  fflush(stdout);
  fprintf(stderr, "candmc_b200: %s failed (%d): %s\n", what, status, candmc_last_error());
  fflush(stderr);
  _exit(134);
}
}  // namespace

double* get_mat_handle(OFF_MAT omat) {
  double* p = nullptr;
  check(candmc_off_host_mirror((int)omat, &p), "get_mat_handle");
  return p;
}

void set_mic_rank(int mic_rank) { check(candmc_off_set_device(mic_rank), "set_mic_rank"); }

void wait_gemm() { check(candmc_off_wait_gemm(), "wait_gemm"); }

void offload_gemm_A(char tA, char tB, int m, int n, int k, double alpha, int offset_A, OFF_MAT omat_A, int lda_A,
                    int offset_B, OFF_MAT omat_B, int lda_B, double beta, int offset_C, OFF_MAT omat_C, int lda_C) {
  check(candmc_off_gemm(tA, tB, m, n, k, alpha, offset_A, (int)omat_A, lda_A, offset_B, (int)omat_B, lda_B, beta,
                        offset_C, (int)omat_C, lda_C),
        "offload_gemm_A");
}

void download_lda_cpy(int nrow, int ncol, int lda_A, int lda_B, int offset_A, double* B, OFF_MAT omat_A) {
  check(candmc_off_download(nrow, ncol, lda_A, lda_B, offset_A, B, (int)omat_A), "download_lda_cpy");
}

void upload_lda_cpy(int nrow, int ncol, int lda_A, int lda_B, double const* A, int offset_B, OFF_MAT omat_B) {
  check(candmc_off_upload(nrow, ncol, lda_A, lda_B, A, offset_B, (int)omat_B), "upload_lda_cpy");
}

void offload_sparse_rw(int nrow, int ncol, int lda_B, double* A, int lda_A, int* offsets_transfer, OFF_MAT omat_B,
                       char rw) {
  check(candmc_off_sparse_rw(nrow, ncol, lda_B, A, lda_A, offsets_transfer, (int)omat_B, rw), "offload_sparse_rw");
}

void alloc_A(int64_t size, double* ptr) { check(candmc_off_alloc(CANDMC_OFF_A, size, ptr), "alloc_A"); }
void alloc_L(int64_t size) { check(candmc_off_alloc(CANDMC_OFF_L, size, nullptr), "alloc_L"); }
void alloc_U(int64_t size) { check(candmc_off_alloc(CANDMC_OFF_U, size, nullptr), "alloc_U"); }
void alloc_transfer(int64_t size) { check(candmc_off_alloc_transfer(size), "alloc_transfer"); }
void free_offload_A() { check(candmc_off_free(CANDMC_OFF_A), "free_offload_A"); }
void free_offload_L() { check(candmc_off_free(CANDMC_OFF_L), "free_offload_L"); }
void free_offload_U() { check(candmc_off_free(CANDMC_OFF_U), "free_offload_U"); }
void free_offload_transfer() { check(candmc_off_free_transfer(), "free_offload_transfer"); }
