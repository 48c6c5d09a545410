/* candmc/full_to_band.h — the trailing update of the symmetric full -> band reduction on the GPU.
 *
 * The reference's sym_full2band (alg/SE/full_to_band.cxx:28-250, declared in alg/SE/CANSE.h:17-22) does, per level, a panel
 * QR on the host (QR_2D_pipe, :96) and then the update of the trailing matrix (:103-243).  This header gives the update a
 * name with the reference's own argument conventions, so that a maintainer replaces those lines of the reference by one
 * call; the panel QR and the recursion stay where they are.  `pv` is the view as on ENTRY to sym_full2band for this level
 * (pv->rrow not yet rotated, :90); A is the reference's `A` argument (the level's working corner) in device memory, Y the
 * aggregated panel (mb x b, ld lda_Y, device memory).  Requires b / b_sub and (n - b) / b_sub to be multiples of the grid
 * dimension (see candmc_b200.h).
 */
#ifndef CANDMC_FULL_TO_BAND_H
#define CANDMC_FULL_TO_BAND_H

#include "../candmc_b200.h"
#include "comm.h"

inline void sym_full2band_update(double* A, int64_t lda_A, int64_t n, int64_t b, int64_t b_sub, pview const* pv, double const* Y,
                                 int64_t lda_Y) {
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_sym_full2band_update(A, lda_A, n, b, b_sub, &c, pv->cdiag.cm, Y, lda_Y, 0), "sym_full2band_update");
}

#endif
