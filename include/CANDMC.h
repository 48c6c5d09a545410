/* CANDMC.h — umbrella header of the B200-native CANMM path, reachable as `#include "CANDMC.h"` like the
 * reference's include/CANDMC.h:1-37.  It covers the multiply entry points (SUMMA, 2.5D, 4D Cannon, split-dim Cannon),
 * cdgemm, lda_cpy and the processor-grid descriptors; the LU / QR / SE headers of the reference umbrella are outside
 * this library's scope. */
#ifndef __CANDMC_H__
#define __CANDMC_H__

/* SUMMA and Cannon */
#include "candmc/topo_pdgemm_algs.h"

/* Split-dimensional Cannon's algorithm */
#include "candmc/spcannon.h"

/* block-cyclic <-> blocked layout bridge (no reference counterpart; see the header) */
#include "candmc/redist.h"

/* trailing update of the symmetric full -> band reduction (the GPU half of alg/SE/full_to_band.cxx's sym_full2band) */
#include "candmc/full_to_band.h"

/* CAQR trailing-matrix updates (the GPU half of alg/QR/qr_2d: update_A / upd_A / update_Yamamoto_A / upd_Yamamoto_A) */
#include "candmc/qr_2d.h"

/* local multiply + packing */
#include "candmc/lapack.h"
#include "candmc/util.h"

#endif
