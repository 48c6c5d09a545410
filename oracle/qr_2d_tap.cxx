// qr_2d_tap — TEST INFRASTRUCTURE (oracle/).  The reference's test/QR/test_qr_2d.cxx hard-wires `#define PIPE_ON` and so only
// ever drives QR_2D_pipe (trailing updates in the W_is_T form).  Linked in front of the reference's objects, this definition of
// QR_2D_pipe sends the same unmodified test through QR_2D_2D instead — what its #else branch calls (test_qr_2d.cxx:132) — so that
// the other two forms of upd_A run under the reference's own ||A - QR|| criterion as well: T from the panel QR's factor
// (QR_2D -> update_A :170,:250) and, with an outer block QR_TAP_B2 < min(m, k), T from the aggregated Y (W == NULL, :873).
#include <stdlib.h>

#include "CANDMC.h"

void QR_2D_pipe(double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b, pview* pv, double* last_Y, int64_t lda_lY, double* last_W,
                double* my_last_W) {
  (void)last_Y; (void)lda_lY; (void)last_W; (void)my_last_W;
  const char* e = getenv("QR_TAP_B2");
  const int64_t b2 = e ? atoll(e) : m;   // the test's default: b2 = m, one level of blocking (QR_2D over the whole matrix)
  QR_2D_2D(A, lda_A, m, k, b2, b, pv, NULL, 0);
}
