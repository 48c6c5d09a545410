/* candmc/qr_2d.h — the CAQR trailing-matrix updates on the GPU, under the reference's own names and argument lists.
 *
 * Replaces, for DEVICE-resident operands, the declarations of alg/QR/qr_2d/qr_2d.h:74-108 (update_A, upd_A) and
 * alg/QR/qr_2d/qr_2d.h (update_Yamamoto_A, upd_Yamamoto_A; definitions alg/QR/qr_2d/qr_y2d.cxx:68-169).  The panel
 * factorisation (TSQR + Householder reconstruction, hh_recon_qr, qr_2d.cxx:311-313) stays on the host in the reference; a
 * maintainer who keeps the matrix in HBM uploads the panel (Y, and W — b x b) and calls these instead of the host routines.
 *
 *   update_A   W == NULL            T is formed from Y            (compute_invT_from_Y, qr_2d.cxx:22-60)
 *              W != NULL, W_is_T    W is the lower-triangular T
 *              W != NULL, !W_is_T   W is the panel QR's upper-triangular factor on the root rank (what QR_2D itself passes,
 *                                   qr_2d.cxx:325); T = lower(-W^-T Y1)   (comp_bcast_T_from_W, qr_2d.cxx:179-208)
 * Every pointer is a device pointer; the calls are asynchronous on the library's default stream (0) like the other
 * candmc/ wrappers, and abort through candmc_shim_check on an argument error as the reference does through ABORT.
 */
#ifndef CANDMC_QR_2D_H
#define CANDMC_QR_2D_H

#include <cstdio>

#include "../candmc_b200.h"
#include "comm.h"
#include "util.h"

class aggregator;   /* alg/QR/qr_2d/qr_y2d.h: only ever passed as NULL here */

inline void candmc_qr2d_unsupported(bool bad, const char* what) {
  if (!bad) return;
  std::fprintf(stderr, "candmc_b200: %s\n", what);
  ABORT;
}

inline void update_A(double const* Y, int64_t lda_Y, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b, double const* W,
                     pview* pv, double* aggreg_Y, int64_t lda_aY, bool W_is_T = false) {
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_update_A(Y, lda_Y, A, lda_A, m, k, b, W, &c, aggreg_Y, lda_aY, W_is_T ? 1 : 0, 0), "update_A");
}

/* upd_A (qr_2d.cxx:224-282) after the panel has been broadcast: W == NULL and the panel-factor form need the whole grid view
 * and go through update_A above; this overload is the W_is_T form the pipelined drivers use (qr_2d.cxx:447-620). */
inline void upd_A(double const* Ybuf, int64_t lda_Y, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b, double const* T,
                  pview* pv, bool W_is_T = true) {
  candmc_qr2d_unsupported(!W_is_T || T == nullptr, "upd_A: only the W_is_T form is offered here; use update_A for the other two");
  candmc_shim_check(candmc_upd_A(Ybuf, lda_Y, A, lda_A, mb, kb, b, T, pv->ccol.cm, 0), "upd_A");
}

inline void update_Yamamoto_A(double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b, double* T,
                              pview* pv, aggregator* agg) {
  candmc_qr2d_unsupported(agg != nullptr, "update_Yamamoto_A: the aggregator is not supported (agg must be NULL)");
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_update_Yamamoto_A(Qm, lda_Qm, A, lda_A, m, k, b, T, &c, 0), "update_Yamamoto_A");
}

inline void upd_Yamamoto_A(double const* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                           double const* T, pview* pv) {
  candmc_shim_check(candmc_upd_Yamamoto_A(Qm, lda_Qm, A, lda_A, mb, kb, b, T, pv->ccol.cm, 0), "upd_Yamamoto_A");
}

#endif
