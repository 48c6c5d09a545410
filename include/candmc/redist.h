/* candmc/redist.h — block-cyclic <-> blocked redistribution on a pview grid.
 *
 * The reference never converts between the two layouts: its multiplies (alg/MM) take blocked matrices
 * (test/MM/topo_pdgemm_unit.cxx:250-256) while its QR / SE drivers hold ScaLAPACK-style block-cyclic ones
 * (test/QR/test_qr_2d.cxx:87-94, alg/SE/dmatrix.cxx:194-203), each with its own test driver.  These two calls are the
 * bridge, so that a block-cyclic caller can use the B200 multiplies.  They follow the conventions of update_A
 * (alg/QR/qr_2d/qr_2d.h:74-85): pview carries the row / column communicators and the current roots.
 * Device pointers; (m/nprow) x (n/npcol) local pieces; m % (nb*nprow) == 0 and n % (nb*npcol) == 0.
 */
#ifndef CANDMC_REDIST_H
#define CANDMC_REDIST_H

#include "../candmc_b200.h"
#include "comm.h"

inline void cyclic_to_blocked(int64_t m, int64_t n, int64_t nb, double const* A_cyc, int64_t lda_cyc, double* A_blk,
                              int64_t lda_blk, pview const* pv) {
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_redistribute(0, m, n, nb, A_cyc, lda_cyc, A_blk, lda_blk, &c, 0), "cyclic_to_blocked");
}

inline void blocked_to_cyclic(int64_t m, int64_t n, int64_t nb, double const* A_blk, int64_t lda_blk, double* A_cyc,
                              int64_t lda_cyc, pview const* pv) {
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_redistribute(1, m, n, nb, A_blk, lda_blk, A_cyc, lda_cyc, &c, 0), "blocked_to_cyclic");
}

#endif
