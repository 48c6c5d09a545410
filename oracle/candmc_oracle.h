/* candmc_oracle — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-process restatement of the reference CANMM hot path (solomonik/CANDMC), used as the parity
 * checker for the CUDA library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
 * load this; the product (candmc_b200/, include/) never does.
 *
 * Pinning: the reference ships no golden vectors (its tests recompute the serial product with BLAS at run time and
 * compare to 1e-6 absolute, test/MM/topo_pdgemm_unit.cxx:321-333, test/MM/test_spc.cxx:116-126).  This oracle is
 * pinned against OUTPUTS OF THE UNMODIFIED REFERENCE run in the build container (oracle/_ref/ref_dump under the
 * mini-MPI shim; fixtures in tests/golden/, generator tests/golden/make_golden.py).  The local multiply in the
 * reference is a vendor `dgemm_` whose version is not pinned by the reference (configure accepts any BLAS), so
 * bit-level equality is not defined; the bar is BASELINE.json's rel-Frobenius <= 10*n*eps.
 *
 * All "distributed" routines simulate every rank of the processor grid inside one process: `blocks` arguments are
 * arrays of P pointers, one local block per world rank, in the reference's rank order.
 */
#ifndef CANDMC_ORACLE_H
#define CANDMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* glibc srand48/drand48 (the reference tests' generator) */
void oracle_srand48(uint64_t* state, int64_t seed);
double oracle_drand48(uint64_t* state);
/* element (row r, col c) of the unit-test matrices: which = 0 -> A, 1 -> B  (test/MM/topo_pdgemm_unit.cxx:250-256) */
double oracle_unit_elem(int64_t r, int64_t c, int64_t n, int which);
void oracle_fill_unit_block(double* X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0, int64_t col0, int64_t n,
                            int which);

/* local kernels */
void oracle_dgemm(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                  int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
void oracle_lda_cpy(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B);
void oracle_lda_cpy_scaled(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                           double a, double b);
void oracle_transpose(int64_t rows, int64_t cols, const double* A, int64_t lda, double* B, int64_t ldb);

/* distributed multiplies; P = number of simulated ranks, blocks indexed by world rank.  Return 0 or -1 (bad grid). */
int oracle_summa(int64_t n, int q, char trans_A, char trans_B, double* const* A, int64_t lda_A, double* const* B,
                 int64_t lda_B, double* const* C, int64_t lda_C);
int oracle_d25_summa(int64_t n, int q, int c, int ovp, char trans_A, char trans_B, double* const* A, double* const* B,
                     double* const* C);
int oracle_bcast_cannon_4d_t(int64_t n, int x1_np, int x2_np, int ovp, char trans_A, char trans_B, double* const* A,
                             double* const* B, double* const* C);
int oracle_bcast_cannon_4d(int64_t n, int x1_np, int x2_np, int ovp, double* const* A, double* const* B,
                           double* const* C);
int oracle_spcannon(int bidir, int kary, int ndim, int n, int m, int k, char transp_A, double alpha, double* const* A,
                    char transp_B, double beta, double* const* B, double* const* C);
/* CAQR trailing update on one process column of nprow ranks: A_r <- A_r - Y_r * (T^-1 * sum_r(Y_r^T A_r)) */
int oracle_upd_A(int nprow, const int64_t* mb, int64_t kb, int64_t b, double* const* Y, const int64_t* lda_Y,
                 double* const* A, const int64_t* lda_A, const double* T);

/* update_A on an nprow x npcol grid (rank = myrow + mycol*nprow, test/QR/test_qr_2d.cxx:367-374) with the current root
 * (rrow, rcol): local extents follow the block-cyclic formulas of qr_2d.cxx:140-147.  Y[r]: local mb x b panel of the ranks
 * in the root column (others ignored), A[r]: local mb x kb trailing block (ld = mb).  mode 0: W == NULL (T from Y,
 * qr_2d.cxx:22-60); mode 1: W is the b x b lower-triangular T (W_is_T).  mb_out/kb_out (may be NULL) receive the extents. */
int oracle_update_A(int nprow, int npcol, int rrow, int rcol, int64_t m, int64_t k, int64_t b, double* const* Y,
                    double* const* A, const double* W, int mode, int64_t* mb_out, int64_t* kb_out);
/* update_Yamamoto_A with agg == NULL (alg/QR/qr_2d/qr_y2d.cxx:68-169) on the same grid / layout conventions as
 * oracle_update_A: Qm[r] = the root column's local mb x b panels, T = the root column's b x b matrix (every column uses it
 * after the MPI_Bcast of :112).  A_r <- A_r + Qm_r * (T * sum_rows(Qm^T A)). */
int oracle_update_Yamamoto_A(int nprow, int npcol, int rrow, int rcol, int64_t m, int64_t k, int64_t b, double* const* Qm,
                             double* const* A, const double* T);
void oracle_update_A_extents(int nprow, int npcol, int rrow, int rcol, int myrow, int mycol, int64_t m, int64_t k,
                             int64_t b, int64_t* mb, int64_t* kb);

/* ---- block-cyclic <-> blocked redistribution by the global index map, all ranks simulated.
 * Block-cyclic: local (row, col) of grid rank (myrow, mycol) is global row ((myrow - rrow) mod nprow)*nb + row%nb +
 * (row/nb)*nb*nprow, columns likewise — the generator of test/QR/test_qr_2d.cxx:87-94 (there rrow = rcol = 0), rotated
 * roots as in qr_2d.cxx:140-147 / dmatrix.cxx:194-203.  Blocked: global row myrow*(m/nprow) + row
 * (test/MM/topo_pdgemm_unit.cxx:250-256).  in/out[myrow + mycol*nprow] are the (m/nprow) x (n/npcol) local pieces, ld = m/nprow.
 * There is no reference routine that performs this conversion (the reference never mixes the two layouts), so this
 * restatement pins the LAYOUT DEFINITIONS only; see tests/test_redist.py. */
int oracle_redistribute(int to_cyclic, int64_t m, int64_t n, int64_t nb, int nprow, int npcol, int rrow, int rcol,
                        double* const* in, double* const* out);
/* global row (or column) index of local index `loc` on grid coordinate `me` of `np` in the block-cyclic layout */
int64_t oracle_cyclic_global_index(int64_t loc, int64_t nb, int me, int np, int root);

/* ---- accelerator seam of the 2.5D LU: restatement of the reference's HOST FALLBACK of
 * alg/LU/lu_offload.cxx (the #else branches: three host arrays, lda_cpy and cdgemm on them).  `mat` is 0/1/2 for
 * OFF_A/OFF_L/OFF_U (lu_offload.h:19).  Pinned by tests/golden (oracle/ref_off_dump.cxx runs the same scripts through the
 * unmodified lu_offload.cxx). */
typedef struct oracle_off {
  double* mat[3];
  int64_t size[3];
} oracle_off_t;
void oracle_off_init(oracle_off_t* o);
void oracle_off_destroy(oracle_off_t* o);
int oracle_off_alloc(oracle_off_t* o, int mat, int64_t size);                       /* lu_offload.cxx:479-531 */
double* oracle_off_handle(oracle_off_t* o, int mat);                                /* get_mat_handle :159-175 */
int oracle_off_gemm(oracle_off_t* o, char tA, char tB, int64_t m, int64_t n, int64_t k, double alpha, int64_t offset_A,
                    int mat_A, int64_t lda_A, int64_t offset_B, int mat_B, int64_t lda_B, double beta, int64_t offset_C,
                    int mat_C, int64_t lda_C);                                      /* offload_gemm_A :216-251 */
int oracle_off_upload(oracle_off_t* o, int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A,
                      int64_t offset_B, int mat_B);                                 /* upload_lda_cpy :366-392 */
int oracle_off_download(oracle_off_t* o, int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, int64_t offset_A,
                        double* B, int mat_A);                                      /* download_lda_cpy :338-364 */
int oracle_off_sparse_rw(oracle_off_t* o, int64_t nrow, int64_t ncol, int64_t lda_B, double* A, int64_t lda_A,
                         const int* offsets, int mat_B, char rw);                   /* offload_sparse_rw :424-476 */
/* test-data generator shared by the script drivers (ours, not the reference's): element idx of stream `seed` in [-0.5,0.5) */
double oracle_off_value(uint64_t seed, uint64_t idx);

#ifdef __cplusplus
}
#endif
#endif
