/* integration/qr_2d_upd_A_gpu.cxx — the CAQR trailing update of the REFERENCE's own 2D QR drivers on a B200.
 *
 * What a CANDMC maintainer adds to the reference build: this file is compiled against the reference's
 * headers ("CANDMC.h", its comm.h and whatever MPI it was configured with) and linked IN FRONT OF alg/QR/qr_2d/qr_2d.cxx's
 * own upd_A (alg/QR/qr_2d/qr_2d.cxx:224-282) — drop the reference's definition, or link with
 * -Wl,--allow-multiple-definition and this object first, qr_2d.cxx compiled -fPIC so that its own calls go through the
 * symbol.  Every QR driver of the reference (QR_2D via update_A :170, QR_2D_pipe :447-620, QR_2D_2D :873, QR_2D_25D :934) then
 * runs its trailing-matrix GEMM pair, the all-reduce of Y^T A over the grid column and the triangular solve in
 * libcandmc_b200.so (candmc_upd_A, include/candmc_b200.h) — and likewise upd_Yamamoto_A (alg/QR/qr_2d/qr_y2d.cxx:123-169) for
 * QR_Yamamoto_2D / QR_Yamamoto_2D_2D (candmc_upd_Yamamoto_A); the panel factorisation (TSQR + Householder reconstruction), the
 * panel broadcast and the formation of T from the panel's factor stay the reference's host code.
 *
 * The matrices stay where the reference keeps them — in host memory — so candmc_upd_A stages them for the call: this seam is
 * about running the reference's drivers unmodified, not about speed (a caller that keeps the trailing matrix in HBM uses the
 * device-pointer forms of include/candmc/qr_2d.h and pays no copies; that is what bench configuration 5 measures).
 *
 * Communicators: the library needs its own (NCCL) communicator for the grid column.  It is derived once per MPI communicator
 * of the caller: the NCCL unique id travels by MPI_Bcast over pv->cworld, candmc_comm_split cuts out the column with the
 * ranks in the caller's order.  One process per GPU, device = local rank modulo the device count (CANDMC_SEAM_DEVICE
 * overrides).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>

#include "CANDMC.h"        /* the REFERENCE's umbrella header: pview, CommData_t, MPI */
#include "candmc_b200.h"   /* the C ABI */

/* defined in alg/QR/qr_2d/qr_2d.cxx:179-208, not declared in its header */
void comp_bcast_T_from_W(int64_t b, double const* W, double const* A, int lda_A, double** pT, int root, bool is_root, MPI_Comm cm);

namespace {

void seam_check(int rc, const char* what) {
  if (rc == CANDMC_OK) return;
  fprintf(stderr, "qr_2d_upd_A_gpu: %s failed: %s\n", what, candmc_last_error());
  MPI_Abort(MPI_COMM_WORLD, rc);
}

struct SeamWorld {
  candmc_comm_t* world = NULL;
};
struct ColumnKey {
  MPI_Comm world, col;
  bool operator<(const ColumnKey& o) const {
    return memcmp(&world, &o.world, sizeof(MPI_Comm)) < 0 ||
           (memcmp(&world, &o.world, sizeof(MPI_Comm)) == 0 && memcmp(&col, &o.col, sizeof(MPI_Comm)) < 0);
  }
};
std::map<ColumnKey, candmc_comm_t*> g_columns;   /* never freed: the reference's drivers have no tear-down hook */
struct WorldKey {
  MPI_Comm world;
  bool operator<(const WorldKey& o) const { return memcmp(&world, &o.world, sizeof(MPI_Comm)) < 0; }
};
std::map<WorldKey, candmc_comm_t*> g_worlds;
bool g_initialised = false;

/* the library's communicator of my grid column, ranks ordered as in pv->ccol (collective over pv->cworld on first use) */
candmc_comm_t* column_of(pview* pv) {
  ColumnKey key;
  memset(&key, 0, sizeof(key));
  key.world = pv->cworld.cm;
  key.col = pv->ccol.cm;
  std::map<ColumnKey, candmc_comm_t*>::iterator it = g_columns.find(key);
  if (it != g_columns.end()) return it->second;
  if (!g_initialised) {
    int dev = 0;
    const char* e = getenv("CANDMC_SEAM_DEVICE");
    if (e) dev = atoi(e);
    else {
      const char* l = getenv("LOCAL_RANK");
      if (!l) l = getenv("OMPI_COMM_WORLD_LOCAL_RANK");
      if (!l) l = getenv("MINIMPI_RANK");
      dev = l ? atoi(l) : pv->cworld.rank;
    }
    seam_check(candmc_init(dev), "candmc_init");
    g_initialised = true;
  }
  WorldKey wk;
  memset(&wk, 0, sizeof(wk));
  wk.world = pv->cworld.cm;
  candmc_comm_t* world = NULL;
  std::map<WorldKey, candmc_comm_t*>::iterator wt = g_worlds.find(wk);
  if (wt != g_worlds.end()) world = wt->second;
  else {
    unsigned char id[CANDMC_UNIQUE_ID_BYTES];
    memset(id, 0, sizeof(id));
    if (pv->cworld.rank == 0) seam_check(candmc_get_unique_id(id), "candmc_get_unique_id");
    MPI_Bcast(id, (int)sizeof(id), MPI_CHAR, 0, pv->cworld.cm);
    seam_check(candmc_comm_init_rank(id, pv->cworld.np, pv->cworld.rank, &world), "candmc_comm_init_rank");
    g_worlds[wk] = world;
  }
  /* my column = the ranks that share my position in the grid row; order inside it = my position in the grid column */
  candmc_comm_t* col = NULL;
  seam_check(candmc_comm_split(world, pv->crow.rank, pv->ccol.rank, &col), "candmc_comm_split");
  int np = 0, rank = -1;
  candmc_comm_size(col, &np);
  candmc_comm_rank(col, &rank);
  if (np != pv->ccol.np || rank != pv->ccol.rank) {
    fprintf(stderr, "qr_2d_upd_A_gpu: column communicator mismatch (%d/%d vs %d/%d)\n", rank, np, pv->ccol.rank, pv->ccol.np);
    MPI_Abort(MPI_COMM_WORLD, 1);
  }
  g_columns[key] = col;
  return col;
}

}  // namespace

/* Same name, argument list and meaning as alg/QR/qr_2d/qr_2d.h:99-108. */
void upd_A(double const* Ybuf, int64_t lda_Y, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b, double const* W, pview* pv,
           bool W_is_T) {
  candmc_comm_t* ccol = column_of(pv);
  double* T_from_W = NULL;
  double const* T = W;   /* NULL: the library forms T^-1 from Y (compute_invT_from_Y, qr_2d.cxx:22-60) */
  if (W != NULL && !W_is_T) {
    /* T from the panel QR's factor: the reference's own host routine, same arguments as its upd_A passes (:250) */
    comp_bcast_T_from_W(b, W, Ybuf, (int)lda_Y, &T_from_W, pv->rcol + pv->rrow * pv->crow.np,
                        (pv->rrow == pv->ccol.rank) & (pv->rcol == pv->crow.rank), pv->cworld.cm);
    T = T_from_W;
  }
  if (getenv("CANDMC_SEAM_VERBOSE") && pv->cworld.rank == 0)
    fprintf(stderr, "qr_2d_upd_A_gpu: upd_A mb=%lld kb=%lld b=%lld form=%s\n", (long long)mb, (long long)kb, (long long)b,
            W == NULL ? "T from Y" : (W_is_T ? "W is T" : "T from W"));
  seam_check(candmc_upd_A(Ybuf, lda_Y, A, lda_A, mb, kb, b, T, ccol, NULL), "candmc_upd_A");
  if (T_from_W) free(T_from_W);
}

/* Same name, argument list and meaning as alg/QR/qr_2d/qr_y2d.h:85-93 (definition qr_y2d.cxx:123-169): the Yamamoto form,
 * A <- A + Qm * (T * (Qm^T A)) with T kept explicitly.  Serves update_Yamamoto_A (:113) and the aggregated block-column update
 * of QR_Yamamoto_2D_2D (:367); the aggregator itself stays the reference's host code. */
void upd_Yamamoto_A(double const* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b, double const* T,
                    pview* pv) {
  candmc_comm_t* ccol = column_of(pv);
  if (getenv("CANDMC_SEAM_VERBOSE") && pv->cworld.rank == 0)
    fprintf(stderr, "qr_2d_upd_A_gpu: upd_Yamamoto_A mb=%lld kb=%lld b=%lld\n", (long long)mb, (long long)kb, (long long)b);
  seam_check(candmc_upd_Yamamoto_A(Qm, lda_Qm, A, lda_A, mb, kb, b, T, ccol, NULL), "candmc_upd_Yamamoto_A");
}
