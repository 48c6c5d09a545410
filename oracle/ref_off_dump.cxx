/* ref_off_dump — TEST INFRASTRUCTURE (oracle/).  Runs an operation script through the UNMODIFIED reference LU offload
 * seam (alg/LU/lu_offload.cxx compiled with -DOFFLOAD and no accelerator, i.e. its host fallback) and dumps everything
 * the host can observe, so that tests/golden/ can pin the plain-C restatement (oracle_off_*) and the CUDA path to it.
 *
 *   ref_off_dump <script.txt> <out.bin>
 *
 * Script lines (all numbers decimal; see tests/off_script.py for the generator and the other two interpreters):
 *   alloc <mat> <size>
 *   fill  <mat> <seed>                              write v(seed,i) through get_mat_handle, as lu_25d_pvt.cxx:1600-1602 does
 *   up    <nrow> <ncol> <lda_A> <lda_B> <off_B> <mat> <seed>          host source = v(seed, 0 .. lda_A*ncol)
 *   down  <nrow> <ncol> <lda_A> <lda_B> <off_A> <mat>                 host target prefilled with -7 -> output record
 *   gemm  <tA> <tB> <m> <n> <k> <alpha> <offA> <matA> <ldA> <offB> <matB> <ldB> <beta> <offC> <matC> <ldC>
 *   wait
 *   sp    <rw> <nrow> <ncol> <lda_B> <lda_A> <mat> <seed> <off_0> ... host rows = v(seed, 0 .. nrow*lda_A) -> output record (r, s)
 * At the end every allocated matrix is appended whole.  Output records: int64 count, then count doubles.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../oracle/candmc_oracle.h" /* oracle_off_value only (the data generator) */
#include "alg/LU/lu_offload.h"

static void put(FILE* f, const double* x, int64_t n) {
  fwrite(&n, sizeof(n), 1, f);
  fwrite(x, sizeof(double), (size_t)n, f);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s script.txt out.bin\n", argv[0]);
    return 2;
  }
  FILE* in = fopen(argv[1], "r");
  FILE* out = fopen(argv[2], "wb");
  if (!in || !out) {
    perror("open");
    return 2;
  }
  int64_t size[3] = {-1, -1, -1};
  char op[16];
  while (fscanf(in, "%15s", op) == 1) {
    if (!strcmp(op, "alloc")) {
      int mat;
      long long sz;
      if (fscanf(in, "%d %lld", &mat, &sz) != 2) return 3;
      if (mat == 0) alloc_A(sz, NULL);
      else if (mat == 1) alloc_L(sz);
      else alloc_U(sz);
      size[mat] = sz;
    } else if (!strcmp(op, "fill")) {
      int mat;
      unsigned long long seed;
      if (fscanf(in, "%d %llu", &mat, &seed) != 2) return 3;
      double* h = get_mat_handle((OFF_MAT)mat);
      for (int64_t i = 0; i < size[mat]; i++) h[i] = oracle_off_value(seed, (uint64_t)i);
    } else if (!strcmp(op, "up")) {
      int nrow, ncol, lda_A, lda_B, off_B, mat;
      unsigned long long seed;
      if (fscanf(in, "%d %d %d %d %d %d %llu", &nrow, &ncol, &lda_A, &lda_B, &off_B, &mat, &seed) != 7) return 3;
      std::vector<double> A((size_t)lda_A * ncol + 1);
      for (size_t i = 0; i < A.size(); i++) A[i] = oracle_off_value(seed, i);
      upload_lda_cpy(nrow, ncol, lda_A, lda_B, A.data(), off_B, (OFF_MAT)mat);
    } else if (!strcmp(op, "down")) {
      int nrow, ncol, lda_A, lda_B, off_A, mat;
      if (fscanf(in, "%d %d %d %d %d %d", &nrow, &ncol, &lda_A, &lda_B, &off_A, &mat) != 6) return 3;
      std::vector<double> B((size_t)lda_B * ncol + 1, -7.0);
      download_lda_cpy(nrow, ncol, lda_A, lda_B, off_A, B.data(), (OFF_MAT)mat);
      put(out, B.data(), (int64_t)B.size());
    } else if (!strcmp(op, "gemm")) {
      char tA[4], tB[4];
      int m, n, k, offA, matA, ldA, offB, matB, ldB, offC, matC, ldC;
      double alpha, beta;
      if (fscanf(in, "%3s %3s %d %d %d %lf %d %d %d %d %d %d %lf %d %d %d", tA, tB, &m, &n, &k, &alpha, &offA, &matA,
                 &ldA, &offB, &matB, &ldB, &beta, &offC, &matC, &ldC) != 16)
        return 3;
      offload_gemm_A(tA[0], tB[0], m, n, k, alpha, offA, (OFF_MAT)matA, ldA, offB, (OFF_MAT)matB, ldB, beta, offC,
                     (OFF_MAT)matC, ldC);
    } else if (!strcmp(op, "wait")) {
      wait_gemm();
    } else if (!strcmp(op, "sp")) {
      char rw[4];
      int nrow, ncol, lda_B, lda_A, mat;
      unsigned long long seed;
      if (fscanf(in, "%3s %d %d %d %d %d %llu", rw, &nrow, &ncol, &lda_B, &lda_A, &mat, &seed) != 7) return 3;
      std::vector<int> offs((size_t)nrow + 1);
      for (int i = 0; i < nrow; i++)
        if (fscanf(in, "%d", &offs[i]) != 1) return 3;
      std::vector<double> A((size_t)nrow * lda_A + 1);
      for (size_t i = 0; i < A.size(); i++) A[i] = oracle_off_value(seed, i);
      offload_sparse_rw(nrow, ncol, lda_B, A.data(), lda_A, offs.data(), (OFF_MAT)mat, rw[0]);
      if (rw[0] != 'w') put(out, A.data(), (int64_t)A.size());
    } else {
      fprintf(stderr, "ref_off_dump: unknown op '%s'\n", op);
      return 3;
    }
  }
  wait_gemm();
  for (int mat = 0; mat < 3; mat++) {
    if (size[mat] < 0) continue;
    std::vector<double> whole((size_t)size[mat] + 1, -7.0);
    if (size[mat] > 0) download_lda_cpy((int)size[mat], 1, (int)size[mat], (int)size[mat], 0, whole.data(), (OFF_MAT)mat);
    put(out, whole.data(), size[mat]);
  }
  fclose(out);
  fclose(in);
  return 0;
}
