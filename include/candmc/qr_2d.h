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
 * Device pointers are used in place and the calls are asynchronous on the library's default stream (0) like the other candmc/
 * wrappers; update_A, upd_A, update_Yamamoto_A and upd_Yamamoto_A also take HOST pointers — the reference's own callers hold
 * host matrices — which are staged for the call and written back before it returns (the aggregator's arrays: device memory).
 * Argument errors abort through candmc_shim_check as the reference does through ABORT.
 */
#ifndef CANDMC_QR_2D_H
#define CANDMC_QR_2D_H

#include <cstdio>

#include "../candmc_b200.h"
#include "comm.h"
#include "util.h"

/* The reference's aggregator (alg/QR/qr_2d/qr_y2d.h:4-46, qr_y2d.cxx:13-62) with the same members and methods; aQm and aT are
 * DEVICE arrays (zero-filled by the constructor, as the reference's are), n and shift live on the host.  append() happens
 * inside update_Yamamoto_A as on the reference's side (:117-118); for the last panel of a block column, which
 * QR_Yamamoto_2D only broadcasts and appends (:266-271), there is append_last_Yamamoto_panel below. */
class aggregator {
 public:
  int64_t lda_aQm;
  int64_t lda_aT;
  int64_t shift;
  double* aQm;
  double* aT;
  int64_t n;

  aggregator(int64_t lda_aQm_, int64_t lda_aT_) {
    candmc_shim_check(candmc_aggregator_create(lda_aQm_, lda_aT_, &impl_), "aggregator");
    sync_out();
  }
  ~aggregator() { candmc_aggregator_free(&impl_); }
  aggregator(const aggregator&) = delete;
  aggregator& operator=(const aggregator&) = delete;
  void reset() {
    candmc_shim_check(candmc_aggregator_reset(&impl_), "aggregator::reset");
    sync_out();
  }
  void shift_down(int64_t b) {
    sync_in();
    candmc_shim_check(candmc_aggregator_shift_down(&impl_, b), "aggregator::shift_down");
    sync_out();
  }
  candmc_aggregator_t* handle() {   /* what the C ABI takes; picks up n / shift a caller may have changed directly */
    sync_in();
    return &impl_;
  }
  void sync_out() {
    lda_aQm = impl_.lda_aQm; lda_aT = impl_.lda_aT; shift = impl_.shift; n = impl_.n; aQm = impl_.aQm; aT = impl_.aT;
  }

 private:
  void sync_in() { impl_.shift = shift; impl_.n = n; }
  candmc_aggregator_t impl_;
};

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

/* upd_A (qr_2d.cxx:224-282) after the panel has been broadcast: the W_is_T form the pipelined drivers use (:447-620) and the
 * W == NULL form (T^-1 from the aggregated Y, QR_2D_2D :873 — formed on the device).  The panel-factor form (W != NULL,
 * W_is_T == false) needs the whole grid view and goes through update_A above.  Unlike the reference's declaration the default
 * of W_is_T is true here: with device pointers a non-null W can only be T. */
inline void upd_A(double const* Ybuf, int64_t lda_Y, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b, double const* W,
                  pview* pv, bool W_is_T = true) {
  candmc_qr2d_unsupported(W != nullptr && !W_is_T, "upd_A: the panel-factor form (W_is_T == false) is offered by update_A");
  candmc_shim_check(candmc_upd_A(Ybuf, lda_Y, A, lda_A, mb, kb, b, W, pv->ccol.cm, 0), "upd_A");
}

inline void update_Yamamoto_A(double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b, double* T,
                              pview* pv, aggregator* agg) {
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_update_Yamamoto_A_agg(Qm, lda_Qm, A, lda_A, m, k, b, T, &c, agg ? agg->handle() : nullptr, 1, 0),
                    "update_Yamamoto_A");
  if (agg) agg->sync_out();
}

/* The last panel of a block column (QR_Yamamoto_2D, qr_y2d.cxx:266-271): MPI_Bcast of Qm and T along the grid row, then
 * agg->append — no trailing matrix left to update.  m, b and pv as for update_Yamamoto_A. */
inline void append_last_Yamamoto_panel(double* Qm, int64_t lda_Qm, int64_t m, int64_t b, double* T, pview* pv, aggregator* agg) {
  candmc_pview_t c = {pv->rrow, pv->rcol, pv->crow.cm, pv->ccol.cm, pv->cworld.cm};
  candmc_shim_check(candmc_update_Yamamoto_A_agg(Qm, lda_Qm, Qm, lda_Qm, m, 0, b, T, &c, agg->handle(), 0, 0), "aggregator::append");
  agg->sync_out();
}

inline void upd_Yamamoto_A(double const* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                           double const* T, pview* pv) {
  candmc_shim_check(candmc_upd_Yamamoto_A(Qm, lda_Qm, A, lda_A, mb, kb, b, T, pv->ccol.cm, 0), "upd_Yamamoto_A");
}

#endif
