/* ref_f2b_dump — TEST INFRASTRUCTURE (oracle/).  Runs the UNMODIFIED reference sym_full2band (alg/SE/full_to_band.cxx:28-285)
 * under the mini-MPI and dumps, for every level of its recursion, what one trailing update consumes and produces:
 *     <prefix>.L<level>.r<rank>.Ain   the whole local array right after the level's panel QR   (lda = n/pr, n/pr columns)
 *     <prefix>.L<level>.r<rank>.Y     the aggregated Householder panel the QR left (mb x b, ld mb)
 *     <prefix>.L<level>.r<rank>.Aout  the whole local array when the level's update has finished (next panel not started)
 *   ref_f2b_dump <n> <b> <b_sub> <prefix>            on P = pr*pr ranks, grid as in test/SE/test_full2band.cxx:151-180
 * The tap: sym_full2band calls QR_2D_pipe (full_to_band.cxx:96); this file defines QR_2D_pipe itself (linked first, with
 * --allow-multiple-definition) as the reference's own QR_2D (same arguments, the call the reference has commented out right
 * above, :95) plus the dumps.  Nothing of the reference is modified or copied.  Matrix: symmetric, element (i, j) seeded by
 * (min, max) so that numpy can regenerate it (tests/f2b_cases.py). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "CANDMC.h"
/* qr_2d.h comes with CANDMC.h; its include guard is broken (#ifndef __QR_2D_H__ / #define __QR_QD_H__), so no second include */
#include "candmc_oracle.h"

void sym_full2band(double* A, int64_t lda_A, int64_t n, int64_t b, int64_t b_sub, pview* pv);

static std::string g_prefix;
static int g_rank = 0, g_level = 0;
static double* g_base = nullptr;
static size_t g_elems = 0;

static void dump(const std::string& what, int level, const double* x, size_t n) {
  std::string fn = g_prefix + ".L" + std::to_string(level) + ".r" + std::to_string(g_rank) + "." + what;
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f || (n && fwrite(x, sizeof(double), n, f) != n)) MPI_Abort(MPI_COMM_WORLD, 3);
  fclose(f);
}

void QR_2D_pipe(double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b, pview* pv, double* last_Y, int64_t lda_lY,
                double* last_W, double* my_last_W) {
  if (g_level > 0) dump("Aout", g_level - 1, g_base, g_elems);   // the previous level's update is complete
  QR_2D(A, lda_A, m, k, b, pv, last_Y, lda_lY);
  dump("Ain", g_level, g_base, g_elems);
  dump("Y", g_level, last_Y, (size_t)(lda_lY * k));
  g_level++;
}

int main(int argc, char** argv) {
  int myRank, numPes;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &numPes);
  MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
  if (argc < 5) MPI_Abort(MPI_COMM_WORLD, 2);
  const int64_t n = atoll(argv[1]), b = atoll(argv[2]), b_sub = atoll(argv[3]);
  g_prefix = argv[4];
  g_rank = myRank;
  const int pr = (int)lround(sqrt((double)numPes));
  if (pr * pr != numPes || n % (pr * b_sub) != 0) MPI_Abort(MPI_COMM_WORLD, 2);
  const int ipr = myRank % pr, ipc = myRank / pr;
  CommData_t cdt_glb, cdt_row, cdt_col, cdt_diag;
  SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt_glb);
  SETUP_SUB_COMM(cdt_glb, cdt_row, myRank / pr, myRank % pr, pr);
  SETUP_SUB_COMM(cdt_glb, cdt_col, myRank % pr, myRank / pr, pr);
  if (ipr == ipc) {
    SETUP_SUB_COMM(cdt_glb, cdt_diag, ipr, 0, pr);
  } else {
    SETUP_SUB_COMM(cdt_glb, cdt_diag, myRank, 1, pr);
  }
  const int64_t nl = n / pr;
  std::vector<double> loc((size_t)(nl * nl));
  for (int64_t c = 0; c < nl; c++)
    for (int64_t r = 0; r < nl; r++) {
      const int64_t gr = ((r / b_sub) * pr + ipr) * b_sub + r % b_sub;
      const int64_t gc = ((c / b_sub) * pr + ipc) * b_sub + c % b_sub;
      const int64_t lo = gr < gc ? gr : gc, hi = gr < gc ? gc : gr;
      loc[r + c * nl] = oracle_off_value(23, (uint64_t)(lo + hi * n));
    }
  g_base = loc.data();
  g_elems = loc.size();
  pview pv;
  pv.rrow = 0; pv.rcol = 0; pv.crow = cdt_row; pv.ccol = cdt_col; pv.cdiag = cdt_diag; pv.cworld = cdt_glb;
  sym_full2band(loc.data(), nl, n, b, b_sub, &pv);
  if (g_level > 0) dump("Aout", g_level - 1, g_base, g_elems);
  MPI_Finalize();
  return 0;
}
