// ref_dump — TEST INFRASTRUCTURE (oracle/).  A driver of OURS around the UNMODIFIED reference CANMM entry points:
// it builds the processor grid exactly as the reference's own tests do, fills the blocks with the reference
// unit-test generators, calls the reference routine and writes every rank's C block to <prefix>.r<rank>.f64
// (raw little-endian doubles, column-major, in the routine's local layout).  tests/golden/make_golden.py packs those
// files into the committed fixtures, and tests use them to pin both oracle/candmc_oracle.c and the CUDA path.
//
//   ref_dump d25   <n> <c_rep> <ovp 0|1> <prefix>      d25_summa / d25_summa_ovp  (grid + data: test/MM/topo_pdgemm_unit.cxx:178-283)
//   ref_dump summa <n> <prefix>                         summa                      (grid + data: :349-486, ctb_unit)
//   ref_dump dcn   <n> <x2_np> <ovp> <prefix>           bcast_cannon_4d            (grid + data: :13-175, dcn_unit); x2_np must be 1
//   ref_dump d25t  <n> <c_rep> <ovp> <tA> <tB> <prefix> | summat <n> <tA> <tB> <prefix> | dcnt <n> <x2_np> <ovp> <tA> <tB> <prefix>
//                                                       the same three with trans_A / trans_B set (the reference hands the flags to
//                                                       its local dgemm only: every block product is op(A block) * op(B block))
//   ref_dump upda  <m> <k> <b> <nprow> <rrow> <rcol> <prefix>   update_A, W == NULL (alg/QR/qr_2d/qr_2d.cxx:124-177)
//   ref_dump updw  <m> <k> <b> <nprow> <rrow> <rcol> <prefix>   update_A, W = the panel QR's upper-triangular factor, W_is_T == false
//                                                                (what QR_2D hands in, qr_2d.cxx:325; T by comp_bcast_T_from_W :179-208)
//   ref_dump updy  <m> <k> <b> <nprow> <rrow> <rcol> <prefix>   update_Yamamoto_A, agg == NULL (alg/QR/qr_2d/qr_y2d.cxx:68-120)
//   ref_dump updyagg <m> <k> <b> <nprow> <rrow> <rcol> <prefix>   update_Yamamoto_A WITH an aggregator over the k/b panels of an m x k
//                                                                block column, driven the way QR_Yamamoto_2D drives it (qr_y2d.cxx:171-277)
//   ref_dump spc   <bidir> <ndim> <seed> <n> <m> <k> <alpha> <beta> <tB N|T> <prefix>   kput_cannon / kuni_cannon (test/MM/test_spc.cxx:36-114)
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "CANDMC.h"

static void dump(const char* prefix, int rank, const double* C, size_t count) {
  std::string fn = std::string(prefix) + ".r" + std::to_string(rank) + ".f64";
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f || fwrite(C, sizeof(double), count, f) != count) {
    fprintf(stderr, "ref_dump: cannot write %s\n", fn.c_str());
    MPI_Abort(MPI_COMM_WORLD, 3);
  }
  fclose(f);
}

static double* alloc_d(size_t n) {
  void* p = NULL;
  if (posix_memalign(&p, ALIGN_BYTES, (n ? n : 1) * sizeof(double)) != 0) MPI_Abort(MPI_COMM_WORLD, 4);
  return (double*)p;
}

static int run_d25(int myRank, int numPes, int64_t n, int c_rep, int ovp, const char* prefix, bool use_summa, char tA = 'N',
                   char tB = 'N') {
  const int num_pes_dim = (int)sqrt((double)(numPes / c_rep));
  if (num_pes_dim * num_pes_dim * c_rep != numPes || n % num_pes_dim != 0) {
    if (myRank == 0) fprintf(stderr, "ref_dump: grid mismatch\n");
    return 2;
  }
  const int64_t b = n / num_pes_dim;
  int layerRank, intraLayerRank, myRow, myCol;
  CommData_t cdt_row, cdt_col, cdt_kdir;
  RSETUP_KDIR_COMM(myRank, numPes, c_rep, cdt_kdir, layerRank, intraLayerRank);
  RSETUP_LAYER_COMM(num_pes_dim, layerRank, intraLayerRank, cdt_row, cdt_col, myRow, myCol);
  double* mat_A = alloc_d(b * b);
  double* mat_B = alloc_d(b * b);
  double* mat_C = alloc_d(b * b);
  double* buffer = alloc_d(5 * b * b);
  memset(buffer, 0, 5 * b * b * sizeof(double));  // the reference needs a zeroed buffer
  memset(mat_C, 0, b * b * sizeof(double));
  for (int64_t i = 0; i < b; i++)
    for (int64_t j = 0; j < b; j++) {
      srand48((myCol * b + i) * n + myRow * b + j);
      mat_A[i * b + j] = drand48();
      mat_B[i * b + j] = drand48();
    }
  ctb_args_t p;
  p.n = n;
  p.lda_A = b;
  p.lda_B = b;
  p.lda_C = b;
  p.buffer_size = 5 * b * b * sizeof(double);
  p.trans_A = tA;   // the reference hands the flags to the local dgemm only (summa.cxx:97, d25_summa.cxx:185)
  p.trans_B = tB;
  p.ovp = ovp;
  if (use_summa)
    summa(&p, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col);
  else if (ovp)
    d25_summa_ovp(&p, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir);
  else
    d25_summa(&p, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir);
  dump(prefix, myRank, mat_C, b * b);
  return 0;
}

static int run_dcn(int myRank, int numPes, int64_t n, int x2_np, int ovp, const char* prefix, char tA = 'N', char tB = 'N') {
  // grid + generator of dcn_unit (test/MM/topo_pdgemm_unit.cxx:13-175)
  const int x1_np = (int)sqrt((double)(numPes / (x2_np * x2_np)));
  if (x1_np * x1_np * x2_np * x2_np != numPes || x2_np != 1) {
    if (myRank == 0) fprintf(stderr, "ref_dump dcn: need x2_np == 1 (the reference deadlocks otherwise) and a square x1 grid\n");
    return 2;
  }
  const int64_t b = n / (x1_np * x2_np);
  const int x1 = myRank % x1_np, y1 = (myRank / x1_np) % x1_np;
  const int x2 = (myRank / (x1_np * x1_np)) % x2_np, y2 = myRank / (x1_np * x1_np * x2_np);
  CommData_t cdt_glb, cdt_x1, cdt_y1, cdt_x2, cdt_y2;
  SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt_glb);
  // same colour/key choice as the driver: x1-communicator = ranks sharing (y1,x2,y2), ordered by x1, etc.
  SETUP_SUB_COMM(cdt_glb, cdt_x1, x1, (y1 + x1_np * (x2 + x2_np * y2)), x1_np);
  SETUP_SUB_COMM(cdt_glb, cdt_y1, y1, (x1 + x1_np * (x2 + x2_np * y2)), x1_np);
  SETUP_SUB_COMM(cdt_glb, cdt_x2, x2, (x1 + x1_np * (y1 + x1_np * y2)), x2_np);
  SETUP_SUB_COMM(cdt_glb, cdt_y2, y2, (x1 + x1_np * (y1 + x1_np * x2)), x2_np);
  double* mat_A = alloc_d(b * b);
  double* mat_B = alloc_d(b * b);
  double* mat_C = alloc_d(b * b);
  double* buffer = alloc_d(5 * b * b);
  memset(buffer, 0, 5 * b * b * sizeof(double));
  const int64_t col0 = (int64_t)(x1 * x2_np + x2) * b, row0 = (int64_t)(y1 * x2_np + y2) * b;
  for (int64_t i = 0; i < b; i++)
    for (int64_t j = 0; j < b; j++) {
      srand48((col0 + i) * n + row0 + j);
      mat_A[i * b + j] = drand48();
      mat_B[i * b + j] = drand48();
    }
  ctb_args_t p;
  p.n = n;
  p.lda_A = b;
  p.lda_B = b;
  p.lda_C = b;
  p.buffer_size = 5 * b * b * sizeof(double);
  p.trans_A = tA;   // ... and so does bcast_cannon_4d (dual_cannon.cxx:163-166,188-194)
  p.trans_B = tB;
  p.ovp = ovp;
  bcast_cannon_4d(&p, mat_A, mat_B, mat_C, buffer, cdt_x1, cdt_y1, cdt_x2, cdt_y2);
  dump(prefix, myRank, mat_C, b * b);
  return 0;
}

static int run_spc(int rank, int numPes, int bidir, int ndim, int seed, int n, int m, int k, double alpha, double beta,
                   char tB, const char* prefix) {
  int kary = 1;
  for (; pow(kary, ndim) < numPes; kary++) {
  }
  int p = 1;
  for (int l = 0; l < ndim; l++) p *= kary;
  if (p != numPes || ndim < 2 || ndim % 2 != 0 || k % ndim != 0) {
    if (rank == 0) fprintf(stderr, "ref_dump spc: bad grid\n");
    return 2;
  }
  int khalf = 1;
  for (int i = 0; i < ndim / 2; i++) khalf *= kary;
  // single-stream generator of test_spc.cxx:66-76
  double* full_A = (double*)malloc(sizeof(double) * m * khalf * k * khalf);
  double* full_B = (double*)malloc(sizeof(double) * k * khalf * n * khalf);
  double* full_C = (double*)malloc(sizeof(double) * m * khalf * n * khalf);
  srand48(seed);
  for (int i = 0; i < m * khalf * k * khalf; i++) full_A[i] = drand48();
  for (int i = 0; i < k * khalf * n * khalf; i++) full_B[i] = drand48();
  for (int i = 0; i < m * khalf * n * khalf; i++) full_C[i] = drand48();
  int px = 0, py = 0, s = 1, tr = rank;
  for (int i = 0; i < ndim / 2; i++) {
    px += (tr % kary) * s;
    tr = tr / kary;
    py += (tr % kary) * s;
    tr = tr / kary;
    s = s * kary;
  }
  double* A = alloc_d((size_t)m * k);
  double* B = alloc_d((size_t)k * n);
  double* C = alloc_d((size_t)m * n);
  for (int i = 0; i < k; i++)
    for (int j = 0; j < m; j++) A[i * m + j] = full_A[(px * k + i) * khalf * m + (py * m + j)];
  if (tB == 'N') {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < k; j++) B[i * k + j] = full_B[(px * n + i) * khalf * k + (py * k + j)];
  } else {  // B block stored transposed (n x k, column-major): Bt[j*n + i] = B(j, i)
    for (int i = 0; i < n; i++)
      for (int j = 0; j < k; j++) B[j * n + i] = full_B[(px * n + i) * khalf * k + (py * k + j)];
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) C[i * m + j] = full_C[(px * n + i) * khalf * m + (py * m + j)];
  if (bidir)
    kput_cannon(rank, kary, ndim, MPI_COMM_WORLD, n, m, k, 'N', alpha, A, tB, beta, B, C);
  else
    kuni_cannon(rank, kary, ndim, MPI_COMM_WORLD, n, m, k, 'N', alpha, A, tB, beta, B, C);
  dump(prefix, rank, C, (size_t)m * n);
  return 0;
}

// update_A (alg/QR/qr_2d/qr_2d.cxx:124-177) with W == NULL: T is formed from Y (compute_invT_from_Y, :22-60), then upd_A.
// Grid and block-cyclic layout as in test/QR/test_qr_2d.cxx:60-94,367-374: myrow = rank % nprow, mycol = rank / nprow.
// Local row block lb of the remaining matrix is global block lb*nprow + (myrow - rrow) mod nprow; trailing column block lb
// is global block lb*npcol + (mycol - rcol - 1) mod npcol.  Elements are seeded by their global coordinates.
// with_W: the third form — W is the b x b upper-triangular factor on the root rank only (every other rank gets a poisoned
// buffer: upd_A reads W on the root alone, :250), W_is_T == false.  NB upd_A names the broadcast root rcol + rrow*npcol
// (:250) while the grid of the reference's own driver is rank = myrow + mycol*nprow (test/QR/test_qr_2d.cxx:367-374): the two
// agree only where rrow == rcol on square grids, on the zero root and on one-dimensional grids, so fixtures use those.
static int run_upda(int myRank, int numPes, int64_t m, int64_t k, int64_t b, int nprow, int rrow, int rcol,
                    const char* prefix, bool with_W = false) {
  const int npcol = numPes / nprow;
  if (nprow * npcol != numPes || m % b || k % b) return 2;
  const int myrow = myRank % nprow, mycol = myRank / nprow;
  CommData_t cdt_glb, cdt_row, cdt_col;
  SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt_glb);
  SETUP_SUB_COMM(cdt_glb, cdt_row, myRank / nprow, myRank % nprow, npcol);
  SETUP_SUB_COMM(cdt_glb, cdt_col, myRank % nprow, myRank / nprow, nprow);
  pview pv;
  pv.rrow = rrow; pv.rcol = rcol; pv.crow = cdt_row; pv.ccol = cdt_col; pv.cworld = cdt_glb;
  int64_t mb = (m / b) / nprow;
  if ((myrow + nprow - rrow) % nprow < (m / b) % nprow) mb++;
  mb *= b;
  int64_t kb = (k / b) / npcol;
  if ((mycol + npcol - rcol - 1) % npcol < (k / b) % npcol) kb++;
  kb *= b;
  double* Y = alloc_d((size_t)(mb ? mb : 1) * b);
  double* A = alloc_d((size_t)(mb ? mb : 1) * (kb ? kb : 1));
  for (int64_t j = 0; j < b; j++)
    for (int64_t r = 0; r < mb; r++) {
      const int64_t gr = ((r / b) * nprow + (myrow - rrow + nprow) % nprow) * b + r % b;
      srand48(7000 + gr * b + j);
      Y[r + j * mb] = (drand48() - .5) * 0.25;
    }
  for (int64_t cc = 0; cc < kb; cc++)
    for (int64_t r = 0; r < mb; r++) {
      const int64_t gr = ((r / b) * nprow + (myrow - rrow + nprow) % nprow) * b + r % b;
      const int64_t gc = ((cc / b) * npcol + (mycol - rcol - 1 + npcol) % npcol) * b + cc % b;
      srand48(900000 + gc * m + gr);
      A[r + cc * mb] = drand48() - .5;
    }
  if (with_W) {
    if (rcol + rrow * npcol != rrow + rcol * nprow) return 2;
    double* W = alloc_d((size_t)b * b);
    const bool root = (myrow == rrow && mycol == rcol);
    for (int64_t j = 0; j < b; j++)
      for (int64_t i = 0; i < b; i++) {
        srand48(333000 + i + j * b);
        const double v = drand48();
        W[i + j * b] = !root ? 77.0 : (i > j ? -55.0 : (i == j ? 1.0 + 0.5 * v : (v - .5) * 0.2));
      }
    update_A(Y, mb, A, mb, m, k, b, W, &pv, NULL, 0, false);
  } else {
    update_A(Y, mb, A, mb, m, k, b, NULL, &pv, NULL, 0);
  }
  dump(prefix, myRank, A, (size_t)mb * kb);
  return 0;
}

// update_Yamamoto_A (alg/QR/qr_2d/qr_y2d.cxx:68-120) with agg == NULL, same grid, layout and element seeds as run_upda.  T is
// a generic dense b x b matrix on the root column; the other columns start with zeros and must receive it (:112).
static int run_updy(int myRank, int numPes, int64_t m, int64_t k, int64_t b, int nprow, int rrow, int rcol,
                    const char* prefix) {
  const int npcol = numPes / nprow;
  if (nprow * npcol != numPes || m % b || k % b) return 2;
  const int myrow = myRank % nprow, mycol = myRank / nprow;
  CommData_t cdt_glb, cdt_row, cdt_col;
  SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt_glb);
  SETUP_SUB_COMM(cdt_glb, cdt_row, myRank / nprow, myRank % nprow, npcol);
  SETUP_SUB_COMM(cdt_glb, cdt_col, myRank % nprow, myRank / nprow, nprow);
  pview pv;
  pv.rrow = rrow; pv.rcol = rcol; pv.crow = cdt_row; pv.ccol = cdt_col; pv.cworld = cdt_glb;
  int64_t mb = (m / b) / nprow;
  if ((myrow + nprow - rrow) % nprow < (m / b) % nprow) mb++;
  mb *= b;
  int64_t kb = (k / b) / npcol;
  if ((mycol + npcol - rcol - 1) % npcol < (k / b) % npcol) kb++;
  kb *= b;
  double* Qm = alloc_d((size_t)(mb ? mb : 1) * b);
  double* A = alloc_d((size_t)(mb ? mb : 1) * (kb ? kb : 1));
  double* T = alloc_d((size_t)b * b);
  for (int64_t j = 0; j < b; j++)
    for (int64_t r = 0; r < mb; r++) {
      const int64_t gr = ((r / b) * nprow + (myrow - rrow + nprow) % nprow) * b + r % b;
      srand48(7000 + gr * b + j);
      Qm[r + j * mb] = mycol == rcol ? (drand48() - .5) * 0.25 : 77.0;  // only the root column's panel counts
    }
  for (int64_t cc = 0; cc < kb; cc++)
    for (int64_t r = 0; r < mb; r++) {
      const int64_t gr = ((r / b) * nprow + (myrow - rrow + nprow) % nprow) * b + r % b;
      const int64_t gc = ((cc / b) * npcol + (mycol - rcol - 1 + npcol) % npcol) * b + cc % b;
      srand48(900000 + gc * m + gr);
      A[r + cc * mb] = drand48() - .5;
    }
  for (int64_t j = 0; j < b; j++)
    for (int64_t i = 0; i < b; i++) {
      srand48(555000 + i + j * b);
      T[i + j * b] = mycol == rcol ? (drand48() - .5) * 0.5 : 0.0;
    }
  update_Yamamoto_A(Qm, mb, A, mb, m, k, b, T, &pv, NULL);
  dump(prefix, myRank, A, (size_t)mb * kb);
  return 0;
}

// update_Yamamoto_A with agg != NULL (alg/QR/qr_2d/qr_y2d.cxx:68-120 + aggregator::append :38-62), driven exactly as
// QR_Yamamoto_2D drives it on an m x k block column with blocks of b (:171-277) — the recursion unrolled into a loop, the panel
// factorisation (Yamamoto(), host-side TSQR: out of scope) replaced by synthetic Qm / T per step: every step updates the trailing
// columns with the step's panel, appends the panel to the aggregator, shifts the aggregator down on the root row and rotates
// the roots; the last panel is only appended (:266-271).  Dumped per rank: the local matrix (mb0 x kb0), then the aggregated
// panels aQm (lda_aQm = mb0 rows x k columns), then the aggregated aT (k x k).
static int run_updyagg(int myRank, int numPes, int64_t m, int64_t k, int64_t b, int nprow, int rrow0, int rcol0, const char* prefix) {
  const int npcol = numPes / nprow;
  if (nprow * npcol != numPes || m % b || k % b || m < k) return 2;
  const int myrow = myRank % nprow, mycol = myRank / nprow;
  CommData_t cdt_glb, cdt_row, cdt_col;
  SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt_glb);
  SETUP_SUB_COMM(cdt_glb, cdt_row, myRank / nprow, myRank % nprow, npcol);
  SETUP_SUB_COMM(cdt_glb, cdt_col, myRank % nprow, myRank / nprow, nprow);
  pview pv;
  pv.rrow = rrow0; pv.rcol = rcol0; pv.crow = cdt_row; pv.ccol = cdt_col; pv.cworld = cdt_glb;
  int64_t mb0 = (m / b) / nprow;
  if ((myrow + nprow - rrow0) % nprow < (m / b) % nprow) mb0++;
  mb0 *= b;
  int64_t kb0 = (k / b) / npcol;
  if ((mycol + npcol - rcol0) % npcol < (k / b) % npcol) kb0++;
  kb0 *= b;
  const int64_t lda_A = mb0 ? mb0 : 1;
  double* A = alloc_d((size_t)lda_A * (kb0 ? kb0 : 1));
  for (int64_t cc = 0; cc < kb0; cc++)
    for (int64_t r = 0; r < mb0; r++) {
      const int64_t gr = ((r / b) * nprow + (myrow - rrow0 + nprow) % nprow) * b + r % b;
      const int64_t gc = ((cc / b) * npcol + (mycol - rcol0 + npcol) % npcol) * b + cc % b;
      srand48(900000 + gc * m + gr);
      A[r + cc * lda_A] = drand48() - .5;
    }
  aggregator agg(lda_A, k);
  double* Aptr = A;
  int64_t ms = m, ks = k;
  for (int s = 0;; ++s, ms -= b, ks -= b) {
    int64_t mb = (ms / b) / nprow;
    if ((myrow + nprow - pv.rrow) % nprow < (ms / b) % nprow) mb++;
    mb *= b;
    double* Qm = alloc_d((size_t)(mb ? mb : 1) * b);
    double* T = alloc_d((size_t)b * b);
    for (int64_t j = 0; j < b; j++)
      for (int64_t r = 0; r < mb; r++) {   // global row of the ORIGINAL matrix: s*b + the row inside the remaining matrix
        const int64_t gr = s * b + ((r / b) * nprow + (myrow - pv.rrow + nprow) % nprow) * b + r % b;
        srand48(7000 + 131 * s + gr * b + j);
        Qm[r + j * mb] = mycol == pv.rcol ? (drand48() - .5) * 0.25 : 77.0;   // only the root column's panel counts
      }
    for (int64_t j = 0; j < b; j++)
      for (int64_t i = 0; i < b; i++) {
        srand48(555000 + 977 * s + i + j * b);
        T[i + j * b] = mycol == pv.rcol ? (drand48() - .5) * 0.5 : 0.0;
      }
    if (ks - b > 0 && ms - b > 0) {
      int64_t move_ptr = 0;
      if (pv.crow.rank == pv.rcol) move_ptr = b * lda_A;
      update_Yamamoto_A(Qm, mb, Aptr + move_ptr, lda_A, ms, ks - b, b, T, &pv, &agg);
      if (pv.ccol.rank == pv.rrow) move_ptr += b;
      if (pv.ccol.rank == pv.rrow) agg.shift_down(b);
      pv.rrow = (pv.rrow + 1) % pv.ccol.np;
      pv.rcol = (pv.rcol + 1) % pv.crow.np;
      Aptr += move_ptr;
      free(Qm);
      free(T);
    } else {
      if (ms - b >= 0) {
        MPI_Bcast(Qm, mb * b, MPI_DOUBLE, pv.rcol, pv.crow.cm);
        MPI_Bcast(T, b * b, MPI_DOUBLE, pv.rcol, pv.crow.cm);
        agg.append(mb, b, Qm, mb, T, &pv);
      }
      free(Qm);
      free(T);
      break;
    }
  }
  const size_t na = (size_t)mb0 * kb0, nq = (size_t)lda_A * k, nt = (size_t)k * k;
  double* out = alloc_d(na + nq + nt);
  for (int64_t cc = 0; cc < kb0; cc++) memcpy(out + cc * mb0, A + cc * lda_A, sizeof(double) * mb0);
  memcpy(out + na, agg.aQm, sizeof(double) * nq);
  memcpy(out + na + nq, agg.aT, sizeof(double) * nt);
  dump(prefix, myRank, out, na + nq + nt);
  return 0;
}

int main(int argc, char** argv) {
  int myRank, numPes;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &numPes);
  MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
  int rc = 2;
  if (argc >= 6 && !strcmp(argv[1], "d25"))
    rc = run_d25(myRank, numPes, atoll(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5], false);
  else if (argc >= 4 && !strcmp(argv[1], "summa"))
    rc = run_d25(myRank, numPes, atoll(argv[2]), 1, 0, argv[3], true);
  else if (argc >= 6 && !strcmp(argv[1], "dcn"))
    rc = run_dcn(myRank, numPes, atoll(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5]);
  else if (argc >= 8 && !strcmp(argv[1], "dcnt"))
    rc = run_dcn(myRank, numPes, atoll(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[7], argv[5][0], argv[6][0]);
  else if (argc >= 6 && !strcmp(argv[1], "summat"))
    rc = run_d25(myRank, numPes, atoll(argv[2]), 1, 0, argv[5], true, argv[3][0], argv[4][0]);
  else if (argc >= 8 && !strcmp(argv[1], "d25t"))
    rc = run_d25(myRank, numPes, atoll(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[7], false, argv[5][0], argv[6][0]);
  else if (argc >= 12 && !strcmp(argv[1], "spc"))
    rc = run_spc(myRank, numPes, atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]),
                 atoi(argv[7]), atof(argv[8]), atof(argv[9]), argv[10][0], argv[11]);
  else if (argc >= 9 && !strcmp(argv[1], "upda"))
    rc = run_upda(myRank, numPes, atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), atoi(argv[5]), atoi(argv[6]),
                  atoi(argv[7]), argv[8]);
  else if (argc >= 9 && !strcmp(argv[1], "updw"))
    rc = run_upda(myRank, numPes, atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), atoi(argv[5]), atoi(argv[6]),
                  atoi(argv[7]), argv[8], true);
  else if (argc >= 9 && !strcmp(argv[1], "updyagg"))
    rc = run_updyagg(myRank, numPes, atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]),
                     argv[8]);
  else if (argc >= 9 && !strcmp(argv[1], "updy"))
    rc = run_updy(myRank, numPes, atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), atoi(argv[5]), atoi(argv[6]),
                  atoi(argv[7]), argv[8]);
  else if (myRank == 0)
    fprintf(stderr, "usage: see the header of oracle/ref_dump.cxx\n");
  if (rc != 0) MPI_Abort(MPI_COMM_WORLD, rc);
  MPI_Finalize();
  return 0;
}
