/* candmc/topo_pdgemm_algs.h — the SUMMA / 2.5D / 4D-Cannon entry points with the reference's exact signatures
 * (alg/MM/topo_pdgemm/topo_pdgemm_algs.h:6-59; the USE_MIC extra ints are dropped).  Matrices may be host pointers
 * (what the reference's tests pass) or device pointers (zero-copy); results match the reference to rel. Frobenius
 * 10*n*eps.  Violations abort like the reference's ASSERT/ABORT. */
#ifndef CANDMC_TOPO_PDGEMM_ALGS_H
#define CANDMC_TOPO_PDGEMM_ALGS_H

#include "comm.h"

typedef struct ctb_args {
  char trans_A;
  char trans_B;
  int64_t n;
  int64_t lda_A;
  int64_t lda_B;
  int64_t lda_C;
  int64_t buffer_size;
  int ovp;
} ctb_args_t;

void summa(ctb_args_t const* args, double const* mat_A, double const* mat_B, double* mat_C, double* buffer,
           CommData_t cdt_row, CommData_t cdt_col);

void d25_summa(ctb_args_t const* args, double* mat_A, double* mat_B, double* mat_C, double* buffer, CommData_t cdt_row,
               CommData_t cdt_col, CommData_t cdt_kdir);

void d25_summa_ovp(ctb_args_t const* args, double* mat_A, double* mat_B, double* mat_C, double* buffer,
                   CommData_t cdt_row, CommData_t cdt_col, CommData_t cdt_kdir);

void bcast_cannon_4d(ctb_args_t const* args, double* mat_A, double* mat_B, double* mat_C, double* buffer,
                     CommData_t cdt_x1, CommData_t cdt_y1, CommData_t cdt_x2, CommData_t cdt_y2);

#endif
