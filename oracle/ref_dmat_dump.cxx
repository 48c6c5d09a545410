/* ref_dmat_dump — TEST INFRASTRUCTURE (oracle/).  Runs the UNMODIFIED reference DMatrix pack operations
 * (alg/SE/dmatrix.cxx:252-355,367-394,470-584) under the mini-MPI and dumps each rank's result.
 *   ref_dmat_dump <op> <nrow> <ncol> <b> <nprow> <rrow> <rcol> <factor> <sliced 0|1> <prefix>
 * op: repv | reph | rsh | tpd | fc | fr.  Grid as in test/QR/test_qr_2d.cxx:367-374 (myrow = rank % nprow, mycol = rank / nprow).
 * Elements are seeded by their GLOBAL coordinates; sliced = 1 works on slice(nprow*b, ., npcol*b, .) (lda != local rows). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "CANDMC.h"
#include "alg/SE/dmatrix.h"
#include "candmc_oracle.h"

static void dump(const char* prefix, int rank, const double* x, size_t n) {
  std::string fn = std::string(prefix) + ".r" + std::to_string(rank) + ".f64";
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f || fwrite(x, sizeof(double), n, f) != n) MPI_Abort(MPI_COMM_WORLD, 3);
  fclose(f);
}

int main(int argc, char** argv) {
  int myRank, numPes;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &numPes);
  MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
  if (argc < 11) MPI_Abort(MPI_COMM_WORLD, 2);
  const char* op = argv[1];
  int64_t nrow = atoll(argv[2]), ncol = atoll(argv[3]), b = atoll(argv[4]);
  const int nprow = atoi(argv[5]), rrow = atoi(argv[6]), rcol = atoi(argv[7]);
  const int64_t factor = atoll(argv[8]);
  const int sliced = atoi(argv[9]);
  const char* prefix = argv[10];
  const int npcol = numPes / nprow;
  const int myrow = myRank % nprow, mycol = myRank / nprow;
  CommData_t cdt_glb, cdt_row, cdt_col;
  SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt_glb);
  SETUP_SUB_COMM(cdt_glb, cdt_row, myRank / nprow, myRank % nprow, npcol);
  SETUP_SUB_COMM(cdt_glb, cdt_col, myRank % nprow, myRank / nprow, nprow);
  pview pv;
  pv.rrow = rrow; pv.rcol = rcol; pv.crow = cdt_row; pv.ccol = cdt_col; pv.cworld = cdt_glb; pv.ictxt = -1;
  DMatrix A(nrow, ncol, b, pv);
  const int64_t mr = A.get_mynrow(), mc = A.get_myncol();
  for (int64_t c = 0; c < mc; c++)
    for (int64_t r = 0; r < mr; r++) {
      const int64_t gr = ((r / b) * nprow + (myrow - rrow + nprow) % nprow) * b + r % b;
      const int64_t gc = ((c / b) * npcol + (mycol - rcol + npcol) % npcol) * b + c % b;
      A.data[r + c * A.lda] = oracle_off_value(11, (uint64_t)(gr + gc * nrow));
    }
  DMatrix X = A;
  if (sliced) X = A.slice(nprow * b, nrow - nprow * b, npcol * b, ncol - npcol * b);
  const int64_t xr = X.get_mynrow(), xc = X.get_myncol();
  if (!strcmp(op, "repv")) {
    double* rep = X.replicate_vertical();
    dump(prefix, myRank, rep, (size_t)(X.nrow * xc));
  } else if (!strcmp(op, "reph")) {
    double* rep = X.replicate_horizontal();
    dump(prefix, myRank, rep, (size_t)(X.ncol * xr));
  } else if (!strcmp(op, "rsh")) {
    const int64_t n = X.ncol * xr;
    double* cntrb = (double*)malloc(sizeof(double) * (n ? n : 1));
    for (int64_t i = 0; i < n; i++) cntrb[i] = oracle_off_value(100 + myRank, (uint64_t)i);
    X.reduce_scatter_horizontal(cntrb);
    dump(prefix, myRank, X.data, (size_t)(xr * xc));
  } else if (!strcmp(op, "tpd")) {
    DMatrix T = X.transpose_data();
    dump(prefix, myRank, T.data, (size_t)(xr * xc));
  } else if (!strcmp(op, "fc")) {
    DMatrix F = X.foldcols(factor);
    dump(prefix, myRank, F.data, (size_t)(xr * xc));
  } else if (!strcmp(op, "fr")) {
    DMatrix F = X.foldrows(factor);
    dump(prefix, myRank, F.data, (size_t)(xr * xc));
  } else {
    MPI_Abort(MPI_COMM_WORLD, 2);
  }
  MPI_Finalize();
  return 0;
}
