/* candmc_oracle.c — TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference CANMM hot path.
 * See candmc_oracle.h for scope, pinning and who may load this.  Every routine cites the reference lines
 * (solomonik/CANDMC, relative to /root/reference) whose data movement it follows; arithmetic is a plain
 * C = alpha*op(A)*op(B) + beta*C loop nest (the reference delegates it to an unpinned vendor dgemm_).
 */
#include "candmc_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * glibc srand48 / drand48 (used by test/MM/topo_pdgemm_unit.cxx:250-256 and test/MM/test_spc.cxx:66-76)
 * X0 = (seed mod 2^32) * 2^16 + 0x330E ;  X <- (0x5DEECE66D * X + 0xB) mod 2^48 ;  value = X / 2^48
 * ---------------------------------------------------------------------------------------------------------- */
void oracle_srand48(uint64_t* state, int64_t seed) { *state = (((uint64_t)seed & 0xffffffffULL) << 16) | 0x330EULL; }

double oracle_drand48(uint64_t* state) {
  *state = (0x5DEECE66DULL * *state + 0xBULL) & ((1ULL << 48) - 1);
  return (double)*state * (1.0 / 281474976710656.0);
}

double oracle_unit_elem(int64_t r, int64_t c, int64_t n, int which) {
  uint64_t s;
  double v;
  oracle_srand48(&s, c * n + r);
  v = oracle_drand48(&s);
  if (which) v = oracle_drand48(&s);
  return v;
}

void oracle_fill_unit_block(double* X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0, int64_t col0, int64_t n,
                            int which) {
  for (int64_t c = 0; c < ncol; ++c)
    for (int64_t r = 0; r < nrow; ++r) X[r + c * ld] = oracle_unit_elem(row0 + r, col0 + c, n, which);
}

/* ------------------------------------------------------------------------------------------------------------
 * cdgemm (alg/shared/lapack.cxx:425-434 -> Fortran dgemm_ semantics, column-major)
 * ---------------------------------------------------------------------------------------------------------- */
static int is_t(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }

void oracle_dgemm(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                  int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  const int ta = is_t(transa), tb = is_t(transb);
  double* At = NULL;
  if (m <= 0 || n <= 0) return;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < m; ++i) C[i + j * ldc] = (beta == 0.0) ? 0.0 : beta * C[i + j * ldc];
  if (k <= 0 || alpha == 0.0) return;
  if (ta) { /* make op(A) column-major m x k so the inner loop is unit stride */
    At = (double*)malloc(sizeof(double) * (size_t)m * (size_t)k);
    for (int64_t p = 0; p < k; ++p)
      for (int64_t i = 0; i < m; ++i) At[i + p * m] = A[p + i * lda];
    A = At;
    lda = m;
  }
  for (int64_t j = 0; j < n; ++j) {
    double* cj = C + j * ldc;
    for (int64_t p = 0; p < k; ++p) {
      const double bpj = alpha * (tb ? B[j + p * ldb] : B[p + j * ldb]);
      const double* ap = A + p * lda;
      for (int64_t i = 0; i < m; ++i) cj[i] += ap[i] * bpj;
    }
  }
  free(At);
}

/* lda_cpy (alg/shared/util.h:459-471) */
void oracle_lda_cpy(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B) {
  for (int64_t i = 0; i < ncol; ++i) memcpy(B + lda_B * i, A + lda_A * i, (size_t)nrow * sizeof(double));
}

/* scaled lda_cpy (alg/shared/util.h:484-501): B = B*b + A*a */
void oracle_lda_cpy_scaled(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                           double a, double b) {
  for (int64_t i = 0; i < ncol; ++i)
    for (int64_t j = 0; j < nrow; ++j) B[lda_B * i + j] = B[lda_B * i + j] * b + A[lda_A * i + j] * a;
}

/* out-of-place transpose, the arithmetic of TRANSPOSE/naive_transp (alg/MM/splitdim_cannon/spcannon_internal.h:66-72):
 * B(cols x rows) = A(rows x cols)^T */
void oracle_transpose(int64_t rows, int64_t cols, const double* A, int64_t lda, double* B, int64_t ldb) {
  for (int64_t c = 0; c < cols; ++c)
    for (int64_t r = 0; r < rows; ++r) B[c + r * ldb] = A[r + c * lda];
}

static double* dalloc(size_t n) {
  double* p = (double*)calloc(n ? n : 1, sizeof(double));
  if (!p) abort();
  return p;
}

/* ------------------------------------------------------------------------------------------------------------
 * summa (alg/MM/topo_pdgemm/summa.cxx:26-101).  World rank r = row*q + col (comm.h:183-195 with one layer).
 * For i < q: the rank in grid column i broadcasts its A block along its row (:59-73), the rank in grid row i
 * broadcasts its B block along its column (:74-88), every rank multiplies with beta = (i > 0) (:96-97).
 * ---------------------------------------------------------------------------------------------------------- */
int oracle_summa(int64_t n, int q, char trans_A, char trans_B, double* const* A, int64_t lda_A, double* const* B,
                 int64_t lda_B, double* const* C, int64_t lda_C) {
  if (q <= 0 || n % q != 0) return -1;
  const int64_t b = n / q;
  const int P = q * q;
  double** loc_A = (double**)malloc(sizeof(double*) * (size_t)P);
  double** loc_B = (double**)malloc(sizeof(double*) * (size_t)P);
  for (int r = 0; r < P; ++r) { /* :53-54 — inputs are copied out of their lda */
    loc_A[r] = dalloc((size_t)(b * b));
    loc_B[r] = dalloc((size_t)(b * b));
    oracle_lda_cpy(b, b, lda_A, b, A[r], loc_A[r]);
    oracle_lda_cpy(b, b, lda_B, b, B[r], loc_B[r]);
  }
  for (int i = 0; i < q; ++i)
    for (int row = 0; row < q; ++row)
      for (int col = 0; col < q; ++col) {
        const double* buf_A = loc_A[row * q + i]; /* bcast along cdt_row, root = column i */
        const double* buf_B = loc_B[i * q + col]; /* bcast along cdt_col, root = row i    */
        oracle_dgemm(trans_A, trans_B, b, b, b, 1.0, buf_A, b, buf_B, b, (i > 0) * 1.0, C[row * q + col], lda_C);
      }
  for (int r = 0; r < P; ++r) {
    free(loc_A[r]);
    free(loc_B[r]);
  }
  free(loc_A);
  free(loc_B);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * d25_summa / d25_summa_ovp (alg/MM/topo_pdgemm/d25_summa.cxx:33-281).  World rank r = layer*q*q + row*q + col
 * (comm.h:171-195).  Layer l handles panels i in [l*q/c, (l+1)*q/c) (:124,151), accumulating into a zeroed buf_C
 * with beta = (i > 0) (:200-201) — the reference relies on the buffer being zero, which this
 * restatement makes explicit.  The ovp variant defers each multiply by one step through the ovp_A/ovp_B swap
 * (:125-148); with zeroed buffers its sums are the same panels in the same order.  Finally MPI_Allreduce(SUM) over
 * the depth communicator leaves sum_l buf_C on every layer (:149,221).
 * Extension (NOT in the reference, whose q % c == 0 assert forbids it): q == 1 with c > 1 splits k across the c
 * ranks — rank l multiplies op(A)[:, l*b/c:(l+1)*b/c] * op(B)[l*b/c:(l+1)*b/c, :] — the 2-GPU configuration of SURVEY §8e
 * (there is one block per operand here, so with trans flags the result IS op(A)*op(B)).
 * ---------------------------------------------------------------------------------------------------------- */
int oracle_d25_summa(int64_t n, int q, int c, int ovp, char trans_A, char trans_B, double* const* A, double* const* B,
                     double* const* C) {
  if (q <= 0 || c <= 0 || n % q != 0) return -1;
  const int64_t b = n / q;
  const int q2 = q * q;
  const int ksplit = (q == 1 && c > 1);
  if (!ksplit && q % c != 0) return -1; /* :63 */
  if (ksplit && b % c != 0) return -1;
  double** buf_C = (double**)malloc(sizeof(double*) * (size_t)(q2 * c));
  for (int l = 0; l < c; ++l)
    for (int row = 0; row < q; ++row)
      for (int col = 0; col < q; ++col) {
        const int r = l * q2 + row * q + col;
        double* bc = buf_C[r] = dalloc((size_t)(b * b));
        if (ksplit) {
          const int64_t kb = b / c;
          /* my k-slice of op(A) and op(B): columns of a stored A / rows of a stored B, the other way round for a stored transpose */
          const double* pa = A[r] + (is_t(trans_A) ? l * kb : l * kb * b);
          const double* pb = B[r] + (is_t(trans_B) ? l * kb * b : l * kb);
          oracle_dgemm(trans_A, trans_B, b, b, kb, 1.0, pa, b, pb, b, 0.0, bc, b);
          continue;
        }
        if (!ovp) {
          for (int i = l * (q / c); i < (l + 1) * (q / c); ++i) {
            const double* pa = A[l * q2 + row * q + i]; /* root column i of my row, same layer (:153-158) */
            const double* pb = B[l * q2 + i * q + col]; /* root row i of my column         (:159-164) */
            oracle_dgemm(trans_A, trans_B, b, b, b, 1.0, pa, b, pb, b, (i > 0) * 1.0, bc, b);
          }
        } else {
          double* zero = dalloc((size_t)(b * b)); /* never-written ovp_A / ovp_B of a zeroed buffer */
          const double *ovp_A = zero, *ovp_B = zero;
          int i;
          for (i = l * (q / c); i < (l + 1) * (q / c); ++i) {
            const double* pa = A[l * q2 + row * q + i];
            const double* pb = B[l * q2 + i * q + col];
            if (i > 0) oracle_dgemm(trans_A, trans_B, b, b, b, 1.0, ovp_A, b, ovp_B, b, 1.0, bc, b); /* :136-139 */
            ovp_A = pa; /* :141-142 swap */
            ovp_B = pb;
          }
          oracle_dgemm(trans_A, trans_B, b, b, b, 1.0, ovp_A, b, ovp_B, b, (i > 0) * 1.0, bc, b); /* :147-148 */
          free(zero);
        }
      }
  for (int row = 0; row < q; ++row)
    for (int col = 0; col < q; ++col) {
      double* sum = dalloc((size_t)(b * b));
      for (int l = 0; l < c; ++l) {
        const double* bc = buf_C[l * q2 + row * q + col];
        for (int64_t e = 0; e < b * b; ++e) sum[e] += bc[e];
      }
      for (int l = 0; l < c; ++l) memcpy(C[l * q2 + row * q + col], sum, sizeof(double) * (size_t)(b * b));
      free(sum);
    }
  for (int r = 0; r < q2 * c; ++r) free(buf_C[r]);
  free(buf_C);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * bcast_cannon_4d (alg/MM/topo_pdgemm/dual_cannon.cxx:40-215), INTENDED semantics: the reference's stagger has
 * mismatched tags / waits (:116-135) and deadlocks for x2_np > 1, so for x2_np > 1 this follows
 * the algorithm the code spells out rather than an observable run:
 *   rank r = x1 + x1_np*(y1 + x1_np*(x2 + x2_np*y2))  (test/MM/topo_pdgemm_unit.cxx:53-68),
 *   stagger: A moves along x2 to (x2 - y2), B along y2 to (y2 - x2)            (:107-137),
 *   for i2 < x2_np { for i1 < x1_np { bcast A along x1 from x1==i1, B along y1 from y1==i1; multiply } (:149-195)
 *                    shift A along x2 by -1, B along y2 by -1 }                (:196-213).
 * ---------------------------------------------------------------------------------------------------------- */
static int wrap(int a, int b) { return ((a % b) + b) % b; }

int oracle_bcast_cannon_4d_t(int64_t n, int x1_np, int x2_np, int ovp, char trans_A, char trans_B, double* const* A,
                             double* const* B, double* const* C);
int oracle_bcast_cannon_4d(int64_t n, int x1_np, int x2_np, int ovp, double* const* A, double* const* B,
                           double* const* C) {
  return oracle_bcast_cannon_4d_t(n, x1_np, x2_np, ovp, 'N', 'N', A, B, C);
}

/* trans_A / trans_B go to the local multiply only (cdgemm(p->trans_A, p->trans_B, ...), dual_cannon.cxx:163-166,188-194):
 * the b x b blocks are staggered, broadcast and shifted as they are stored. */
int oracle_bcast_cannon_4d_t(int64_t n, int x1_np, int x2_np, int ovp, char trans_A, char trans_B, double* const* A,
                             double* const* B, double* const* C) {
  (void)ovp; /* only changes when the multiply is issued (:163-166,191-194), not what is summed */
  if (x1_np <= 0 || x2_np <= 0 || n % ((int64_t)x1_np * x2_np) != 0) return -1;
  const int64_t b = n / ((int64_t)x1_np * x2_np);
  const int P = x1_np * x1_np * x2_np * x2_np;
#define RK(x1, y1, x2, y2) ((x1) + x1_np * ((y1) + x1_np * ((x2) + x2_np * (y2))))
  const double** curA = (const double**)malloc(sizeof(double*) * (size_t)P);
  const double** curB = (const double**)malloc(sizeof(double*) * (size_t)P);
  const double** nxtA = (const double**)malloc(sizeof(double*) * (size_t)P);
  const double** nxtB = (const double**)malloc(sizeof(double*) * (size_t)P);
  for (int y2 = 0; y2 < x2_np; ++y2)
    for (int x2 = 0; x2 < x2_np; ++x2)
      for (int y1 = 0; y1 < x1_np; ++y1)
        for (int x1 = 0; x1 < x1_np; ++x1) { /* after the stagger I hold what (x2+y2) resp. (y2+x2) owned */
          curA[RK(x1, y1, x2, y2)] = A[RK(x1, y1, wrap(x2 + y2, x2_np), y2)];
          curB[RK(x1, y1, x2, y2)] = B[RK(x1, y1, x2, wrap(y2 + x2, x2_np))];
        }
  for (int i2 = 0; i2 < x2_np; ++i2) {
    for (int i1 = 0; i1 < x1_np; ++i1)
      for (int y2 = 0; y2 < x2_np; ++y2)
        for (int x2 = 0; x2 < x2_np; ++x2)
          for (int y1 = 0; y1 < x1_np; ++y1)
            for (int x1 = 0; x1 < x1_np; ++x1) {
              const double* mul_A = curA[RK(i1, y1, x2, y2)]; /* bcast along cdt_x1, root i1 */
              const double* mul_B = curB[RK(x1, i1, x2, y2)]; /* bcast along cdt_y1, root i1 */
              oracle_dgemm(trans_A, trans_B, b, b, b, 1.0, mul_A, b, mul_B, b, (i1 > 0 || i2 > 0) * 1.0,
                           C[RK(x1, y1, x2, y2)], b);
            }
    if (i2 < x2_np - 1) {
      for (int y2 = 0; y2 < x2_np; ++y2)
        for (int x2 = 0; x2 < x2_np; ++x2)
          for (int y1 = 0; y1 < x1_np; ++y1)
            for (int x1 = 0; x1 < x1_np; ++x1) { /* receive from +1 (:198-209) */
              nxtA[RK(x1, y1, x2, y2)] = curA[RK(x1, y1, wrap(x2 + 1, x2_np), y2)];
              nxtB[RK(x1, y1, x2, y2)] = curB[RK(x1, y1, x2, wrap(y2 + 1, x2_np))];
            }
      const double** t;
      t = curA, curA = nxtA, nxtA = t;
      t = curB, curB = nxtB, nxtB = t;
    }
  }
#undef RK
  free(curA);
  free(curB);
  free(nxtA);
  free(nxtB);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * kput_cannon / kuni_cannon (alg/MM/splitdim_cannon/spcannon.cxx:237-347): C <- alpha*A*B + beta*C on a
 * kary-ary ndim-cube.  Rank digits base kary: digit 2j = A-direction coordinate, digit 2j+1 = B-direction
 * coordinate (:59-62).  A is canonicalised to m x k, B to B^T = n x k (:262-267); slices are contiguous ranges of
 * those arrays.  uni_stagger (:33-84), bdr_shift (:87-162), uni_shift (:165-234) are simulated for all ranks in
 * lockstep: every MPI_Put between two fences becomes a copy into the target's fresh buffer, then A = buf_A,
 * B = buf_B (:76-77,159-160,231-232).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int P, kary, ndim, n, m, k, bidir;
  double alpha;
  double **A, **B, **bufA, **bufB; /* per rank, m*k and n*k doubles */
  double* const* C;
} spc_t;

static int digit(int rank, int kary, int pos) {
  for (int i = 0; i < pos; ++i) rank /= kary;
  return rank % kary;
}
static int ipow(int a, int e) {
  int r = 1;
  while (e-- > 0) r *= a;
  return r;
}
static void spc_swap(spc_t* s) {
  for (int r = 0; r < s->P; ++r) { /* memcpy(A, buf_A), memcpy(B, buf_B) */
    memcpy(s->A[r], s->bufA[r], sizeof(double) * (size_t)s->m * (size_t)s->k);
    memcpy(s->B[r], s->bufB[r], sizeof(double) * (size_t)s->n * (size_t)s->k);
  }
}

static void spc_stagger(spc_t* s, int level) { /* :33-84 */
  const int half = s->ndim / 2;
  const int64_t bA = 2 * (int64_t)s->m * s->k / s->ndim, bB = 2 * (int64_t)s->k * s->n / s->ndim;
  for (int r = 0; r < s->P; ++r)
    for (int j = 0; j < half; ++j) {
      const int i = (j + level) % half;
      const int tA = digit(r, s->kary, 2 * j), tB = digit(r, s->kary, 2 * j + 1);
      const int dstA = r + (wrap(tA - tB, s->kary) - tA) * ipow(s->kary, 2 * j);
      const int dstB = r + (wrap(tB - tA, s->kary) - tB) * ipow(s->kary, 2 * j + 1);
      memcpy(s->bufA[dstA] + i * bA, s->A[r] + i * bA, sizeof(double) * (size_t)bA);
      memcpy(s->bufB[dstB] + i * bB, s->B[r] + i * bB, sizeof(double) * (size_t)bB);
    }
  spc_swap(s);
  if (level < half - 1) spc_stagger(s, level + 1);
}

static void spc_shift(spc_t* s, int level, double beta) { /* bdr_shift :87-162 / uni_shift :165-234 */
  const int half = s->ndim / 2;
  double dbeta = beta;
  for (int ka = 0; ka < s->kary; ++ka) {
    if (level < half - 1) {
      spc_shift(s, level + 1, dbeta);
    } else {
      for (int r = 0; r < s->P; ++r) /* DGEMM('N','T',m,n,k,alpha,A,m,B,n,dbeta,C,m) :117,195 */
        oracle_dgemm('N', 'T', s->m, s->n, s->k, s->alpha, s->A[r], s->m, s->B[r], s->n, dbeta, s->C[r], s->m);
    }
    dbeta = 1.0;
    for (int r = 0; r < s->P; ++r)
      for (int j = 0; j < half; ++j) {
        const int i = (j + level) % half;
        const int tA = digit(r, s->kary, 2 * j), tB = digit(r, s->kary, 2 * j + 1);
        const int sA = ipow(s->kary, 2 * j), sB = ipow(s->kary, 2 * j + 1);
        const int upA = r + (wrap(tA + 1, s->kary) - tA) * sA, dnA = r + (wrap(tA - 1, s->kary) - tA) * sA;
        const int upB = r + (wrap(tB + 1, s->kary) - tB) * sB, dnB = r + (wrap(tB - 1, s->kary) - tB) * sB;
        if (s->bidir) { /* halves 2i and 2i+1 travel in opposite directions (:139-152) */
          const int64_t bA = (int64_t)s->m * s->k / s->ndim, bB = (int64_t)s->k * s->n / s->ndim;
          memcpy(s->bufA[upA] + 2 * i * bA, s->A[r] + 2 * i * bA, sizeof(double) * (size_t)bA);
          memcpy(s->bufA[dnA] + (2 * i + 1) * bA, s->A[r] + (2 * i + 1) * bA, sizeof(double) * (size_t)bA);
          memcpy(s->bufB[upB] + 2 * i * bB, s->B[r] + 2 * i * bB, sizeof(double) * (size_t)bB);
          memcpy(s->bufB[dnB] + (2 * i + 1) * bB, s->B[r] + (2 * i + 1) * bB, sizeof(double) * (size_t)bB);
        } else { /* whole slice i travels +1 (:217-224) */
          const int64_t bA = 2 * (int64_t)s->m * s->k / s->ndim, bB = 2 * (int64_t)s->k * s->n / s->ndim;
          memcpy(s->bufA[upA] + i * bA, s->A[r] + i * bA, sizeof(double) * (size_t)bA);
          memcpy(s->bufB[upB] + i * bB, s->B[r] + i * bB, sizeof(double) * (size_t)bB);
        }
      }
    spc_swap(s);
  }
}

int oracle_spcannon(int bidir, int kary, int ndim, int n, int m, int k, char transp_A, double alpha, double* const* A,
                    char transp_B, double beta, double* const* B, double* const* C) {
  if (ndim < 2 || ndim % 2 != 0 || kary < 1 || k % ndim != 0) return -1; /* :252 assert(k%ndim == 0) */
  spc_t s;
  s.P = ipow(kary, ndim);
  s.kary = kary;
  s.ndim = ndim;
  s.n = n;
  s.m = m;
  s.k = k;
  s.bidir = bidir;
  s.alpha = alpha;
  s.C = C;
  s.A = (double**)malloc(sizeof(double*) * (size_t)s.P);
  s.B = (double**)malloc(sizeof(double*) * (size_t)s.P);
  s.bufA = (double**)malloc(sizeof(double*) * (size_t)s.P);
  s.bufB = (double**)malloc(sizeof(double*) * (size_t)s.P);
  for (int r = 0; r < s.P; ++r) {
    s.A[r] = dalloc((size_t)m * (size_t)k);
    s.B[r] = dalloc((size_t)n * (size_t)k);
    s.bufA[r] = dalloc((size_t)m * (size_t)k);
    s.bufB[r] = dalloc((size_t)n * (size_t)k);
    /* canonicalise (:262-267): A -> m x k, B -> B^T (n x k) */
    if (is_t(transp_A)) oracle_transpose(k, m, A[r], k, s.A[r], m); else memcpy(s.A[r], A[r], sizeof(double) * (size_t)m * (size_t)k);
    if (!is_t(transp_B)) oracle_transpose(k, n, B[r], k, s.B[r], n); else memcpy(s.B[r], B[r], sizeof(double) * (size_t)n * (size_t)k);
  }
  spc_stagger(&s, 0);
  spc_shift(&s, 0, beta);
  for (int r = 0; r < s.P; ++r) {
    free(s.A[r]);
    free(s.B[r]);
    free(s.bufA[r]);
    free(s.bufB[r]);
  }
  free(s.A);
  free(s.B);
  free(s.bufA);
  free(s.bufB);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * upd_A GEMM pair (alg/QR/qr_2d/qr_2d.cxx:259,265,271,275), W_is_T case, one process column of nprow ranks:
 *   W_r = Y_r^T A_r ; W = sum_r W_r (MPI_Allreduce over ccol) ; W <- T^-1 W (cdtrsm L,L,N,N) ; A_r -= Y_r W
 * ---------------------------------------------------------------------------------------------------------- */
int oracle_upd_A(int nprow, const int64_t* mb, int64_t kb, int64_t b, double* const* Y, const int64_t* lda_Y,
                 double* const* A, const int64_t* lda_A, const double* T) {
  if (nprow <= 0 || kb < 0 || b <= 0) return -1;
  double* W = dalloc((size_t)(b * kb));
  double* Wr = dalloc((size_t)(b * kb));
  for (int r = 0; r < nprow; ++r) {
    if (mb[r] > 0 && kb > 0) {
      oracle_dgemm('T', 'N', b, kb, mb[r], 1.0, Y[r], lda_Y[r], A[r], lda_A[r], 0.0, Wr, b); /* :259 */
      for (int64_t e = 0; e < b * kb; ++e) W[e] += Wr[e];                                     /* :265 */
    }
  }
  for (int64_t j = 0; j < kb; ++j) /* :271 forward substitution with lower-triangular, non-unit T (ld = b) */
    for (int64_t i = 0; i < b; ++i) {
      double x = W[i + j * b];
      for (int64_t p = 0; p < i; ++p) x -= T[i + p * b] * W[p + j * b];
      W[i + j * b] = x / T[i + i * b];
    }
  for (int r = 0; r < nprow; ++r)
    if (mb[r] > 0 && kb > 0)
      oracle_dgemm('N', 'N', mb[r], kb, b, -1.0, Y[r], lda_Y[r], W, b, 1.0, A[r], lda_A[r]); /* :275 */
  free(W);
  free(Wr);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * update_A (alg/QR/qr_2d/qr_2d.cxx:124-177) + upd_A (:224-282) + compute_invT_from_Y (:22-60), all ranks simulated.
 * ---------------------------------------------------------------------------------------------------------- */
void oracle_update_A_extents(int nprow, int npcol, int rrow, int rcol, int myrow, int mycol, int64_t m, int64_t k,
                             int64_t b, int64_t* mb, int64_t* kb) {
  int64_t x = (m / b) / nprow; /* :140-147 */
  if ((myrow + nprow - rrow) % nprow < (m / b) % nprow) x++;
  *mb = x * b;
  x = (k / b) / npcol;
  if ((mycol + npcol - rcol - 1) % npcol < (k / b) % npcol) x++;
  *kb = x * b;
}

int oracle_update_A(int nprow, int npcol, int rrow, int rcol, int64_t m, int64_t k, int64_t b, double* const* Y,
                    double* const* A, const double* W, int mode, int64_t* mb_out, int64_t* kb_out) {
  if (nprow <= 0 || npcol <= 0 || b <= 0 || m % b || k % b) return -1;
  /* Ybuf of every grid row: the root column's panel, on the root row with zeroed upper triangle and unit diagonal
   * (copy_lower + the explicit 1.0, :157-163), then MPI_Bcast along the row (:168) */
  double** Ybuf = (double**)malloc(sizeof(double*) * (size_t)nprow);
  int64_t* mbs = (int64_t*)malloc(sizeof(int64_t) * (size_t)nprow);
  for (int pr = 0; pr < nprow; ++pr) {
    int64_t mb, kb;
    oracle_update_A_extents(nprow, npcol, rrow, rcol, pr, rcol, m, k, b, &mb, &kb);
    mbs[pr] = mb;
    Ybuf[pr] = dalloc((size_t)(mb * b));
    const double* src = Y[pr + rcol * nprow];
    for (int64_t j = 0; j < b; ++j)
      for (int64_t r = 0; r < mb; ++r) {
        double v = src[r + j * mb];
        if (pr == rrow) {
          if (r < j) v = 0.0;
          if (r == j) v = 1.0;
        }
        Ybuf[pr][r + j * mb] = v;
      }
  }
  /* T: mode 1 -> W; mode 0 -> lower triangle of sum_rows Ybuf^T Ybuf with halved diagonal (:36-50), rest zero (:235);
   * mode 2 -> W is the b x b UPPER-triangular factor the panel QR left on the root rank (hh_recon_qr, what QR_2D :325 hands
   * in with W_is_T == false): the root solves W^T X = -Y1 (comp_bcast_T_from_W :193-195, compute_invT_from_W,
   * alg/QR/hh_recon/hh_recon.cxx:26-31: cdtrsm('L','U','T','N', alpha = -1)), Y1 = the top b x b of its Ybuf, and every rank
   * receives the lower triangle of X with its diagonal (pack_lower / MPI_Bcast / unpack_lower into zeros, :198-205) */
  double* T = dalloc((size_t)(b * b));
  if (mode == 1) {
    memcpy(T, W, sizeof(double) * (size_t)(b * b));
  } else if (mode == 2) {
    if (mbs[rrow] < b) return -1;
    const double* Y1 = Ybuf[rrow];
    const int64_t ldy = mbs[rrow];
    double* X = dalloc((size_t)(b * b));
    for (int64_t j = 0; j < b; ++j)
      for (int64_t i = 0; i < b; ++i) { /* forward substitution with the lower-triangular W^T: (W^T)(i,p) = W(p,i) */
        double x = -Y1[i + j * ldy];
        for (int64_t p = 0; p < i; ++p) x -= W[p + i * b] * X[p + j * b];
        X[i + j * b] = x / W[i + i * b];
      }
    for (int64_t j = 0; j < b; ++j)
      for (int64_t i = j; i < b; ++i) T[i + j * b] = X[i + j * b];
    free(X);
  } else {
    double* S = dalloc((size_t)(b * b));
    for (int pr = 0; pr < nprow; ++pr)
      if (mbs[pr] > 0) oracle_dgemm('T', 'N', b, b, mbs[pr], 1.0, Ybuf[pr], mbs[pr], Ybuf[pr], mbs[pr], 1.0, S, b);
    for (int64_t j = 0; j < b; ++j)
      for (int64_t i = j; i < b; ++i) T[i + j * b] = (i == j) ? S[i + j * b] / 2.0 : S[i + j * b];
    free(S);
  }
  /* per grid column: W = sum_rows Ybuf^T A (:259,265), W <- T^-1 W (:271), A -= Ybuf W (:275) */
  for (int pc = 0; pc < npcol; ++pc) {
    int64_t mb0, kb;
    oracle_update_A_extents(nprow, npcol, rrow, rcol, 0, pc, m, k, b, &mb0, &kb);
    if (kb_out) kb_out[pc] = kb;
    if (kb == 0) continue;
    double* Wm = dalloc((size_t)(b * kb));
    double* Wr = dalloc((size_t)(b * kb));
    for (int pr = 0; pr < nprow; ++pr)
      if (mbs[pr] > 0) {
        oracle_dgemm('T', 'N', b, kb, mbs[pr], 1.0, Ybuf[pr], mbs[pr], A[pr + pc * nprow], mbs[pr], 0.0, Wr, b);
        for (int64_t e = 0; e < b * kb; ++e) Wm[e] += Wr[e];
      }
    for (int64_t j = 0; j < kb; ++j)
      for (int64_t i = 0; i < b; ++i) {
        double x = Wm[i + j * b];
        for (int64_t p = 0; p < i; ++p) x -= T[i + p * b] * Wm[p + j * b];
        Wm[i + j * b] = x / T[i + i * b];
      }
    for (int pr = 0; pr < nprow; ++pr)
      if (mbs[pr] > 0)
        oracle_dgemm('N', 'N', mbs[pr], kb, b, -1.0, Ybuf[pr], mbs[pr], Wm, b, 1.0, A[pr + pc * nprow], mbs[pr]);
    free(Wm);
    free(Wr);
  }
  for (int pr = 0; pr < nprow; ++pr) {
    if (mb_out) mb_out[pr] = mbs[pr];
    free(Ybuf[pr]);
  }
  free(Ybuf);
  free(mbs);
  free(T);
  return 0;
}

/* update_Yamamoto_A / upd_Yamamoto_A, qr_y2d.cxx:68-169: per grid column W = sum_rows Qm^T A (:140,146), W2 = -T W (:156),
 * A -= Qm W2 (:160); the panel and T are those of the root column (MPI_Bcast :109-112) */
int oracle_update_Yamamoto_A(int nprow, int npcol, int rrow, int rcol, int64_t m, int64_t k, int64_t b, double* const* Qm,
                             double* const* A, const double* T) {
  if (nprow <= 0 || npcol <= 0 || b <= 0 || m % b || k % b) return -1;
  for (int pc = 0; pc < npcol; ++pc) {
    int64_t mb0, kb;
    oracle_update_A_extents(nprow, npcol, rrow, rcol, 0, pc, m, k, b, &mb0, &kb);
    if (kb == 0) continue;
    double* W = dalloc((size_t)(b * kb));
    double* Wr = dalloc((size_t)(b * kb));
    double* W2 = dalloc((size_t)(b * kb));
    for (int pr = 0; pr < nprow; ++pr) {
      int64_t mb, kb2;
      oracle_update_A_extents(nprow, npcol, rrow, rcol, pr, pc, m, k, b, &mb, &kb2);
      if (mb == 0) continue;
      oracle_dgemm('T', 'N', b, kb, mb, 1.0, Qm[pr + rcol * nprow], mb, A[pr + pc * nprow], mb, 0.0, Wr, b);
      for (int64_t e = 0; e < b * kb; ++e) W[e] += Wr[e];
    }
    oracle_dgemm('N', 'N', b, kb, b, -1.0, T, b, W, b, 0.0, W2, b);
    for (int pr = 0; pr < nprow; ++pr) {
      int64_t mb, kb2;
      oracle_update_A_extents(nprow, npcol, rrow, rcol, pr, pc, m, k, b, &mb, &kb2);
      if (mb == 0) continue;
      oracle_dgemm('N', 'N', mb, kb, b, -1.0, Qm[pr + rcol * nprow], mb, W2, b, 1.0, A[pr + pc * nprow], mb);
    }
    free(W);
    free(Wr);
    free(W2);
  }
  return 0;
}

/* ===== LU accelerator seam: host-fallback semantics of alg/LU/lu_offload.cxx ======================================== */
void oracle_off_init(oracle_off_t* o) { memset(o, 0, sizeof(*o)); }

void oracle_off_destroy(oracle_off_t* o) {
  for (int i = 0; i < 3; i++) free(o->mat[i]);
  memset(o, 0, sizeof(*o));
}

/* alloc_A / alloc_L / alloc_U, non-accelerator branch (lu_offload.cxx:488,513,524): posix_memalign of `size` doubles */
int oracle_off_alloc(oracle_off_t* o, int mat, int64_t size) {
  if (mat < 0 || mat > 2 || size < 0) return -1;
  free(o->mat[mat]);
  o->mat[mat] = (double*)malloc(sizeof(double) * (size_t)(size > 0 ? size : 1));
  o->size[mat] = size;
  return o->mat[mat] ? 0 : -1;
}

/* get_mat_handle (lu_offload.cxx:159-175): the host array itself */
double* oracle_off_handle(oracle_off_t* o, int mat) { return (mat < 0 || mat > 2) ? NULL : o->mat[mat]; }

/* offload_gemm_A, non-accelerator branch (lu_offload.cxx:238-249): cdgemm on handle + offset */
int oracle_off_gemm(oracle_off_t* o, char tA, char tB, int64_t m, int64_t n, int64_t k, double alpha, int64_t offset_A,
                    int mat_A, int64_t lda_A, int64_t offset_B, int mat_B, int64_t lda_B, double beta, int64_t offset_C,
                    int mat_C, int64_t lda_C) {
  double *a = oracle_off_handle(o, mat_A), *b = oracle_off_handle(o, mat_B), *c = oracle_off_handle(o, mat_C);
  if (!a || !b || !c) return -1;
  oracle_dgemm(tA, tB, m, n, k, alpha, a + offset_A, lda_A, b + offset_B, lda_B, beta, c + offset_C, lda_C);
  return 0;
}

/* upload_lda_cpy, non-accelerator branch (lu_offload.cxx:381-382): lda_cpy(host A -> handle + offset_B) */
int oracle_off_upload(oracle_off_t* o, int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A,
                      int64_t offset_B, int mat_B) {
  double* b = oracle_off_handle(o, mat_B);
  if (!b) return -1;
  oracle_lda_cpy(nrow, ncol, lda_A, lda_B, A, b + offset_B);
  return 0;
}

/* download_lda_cpy, non-accelerator branch (lu_offload.cxx:353-354): lda_cpy(handle + offset_A -> host B) */
int oracle_off_download(oracle_off_t* o, int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, int64_t offset_A,
                        double* B, int mat_A) {
  double* a = oracle_off_handle(o, mat_A);
  if (!a) return -1;
  oracle_lda_cpy(nrow, ncol, lda_A, lda_B, a + offset_A, B);
  return 0;
}

/* offload_sparse_rw, non-accelerator branch (lu_offload.cxx:459-475): row i is the strided vector
 * handle[offsets[i] + j*lda_B], j < ncol, and the contiguous host vector A + i*lda_A; rows are processed in order */
int oracle_off_sparse_rw(oracle_off_t* o, int64_t nrow, int64_t ncol, int64_t lda_B, double* A, int64_t lda_A,
                         const int* offsets, int mat_B, char rw) {
  if (ncol == 0 || nrow == 0) return 0;
  double* b = oracle_off_handle(o, mat_B);
  if (!b) return -1;
  if (rw != 'r' && rw != 'w' && rw != 's') return -1;
  for (int64_t i = 0; i < nrow; i++) {
    double* row = b + offsets[i];
    double* host = A + i * lda_A;
    for (int64_t j = 0; j < ncol; j++) {
      double* e = row + j * lda_B;
      if (rw == 'r') {
        host[j] = *e;
      } else if (rw == 'w') {
        *e = host[j];
      } else {
        double t = *e;
        *e = host[j];
        host[j] = t;
      }
    }
  }
  return 0;
}

/* splitmix64 finaliser -> [-0.5, 0.5) */
double oracle_off_value(uint64_t seed, uint64_t idx) {
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + idx;
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

/* ===== block-cyclic <-> blocked redistribution ========================================================================= */
/* test/QR/test_qr_2d.cxx:87-94: global = myrow*b + (row%b) + (row/b)*b*nprow, with the root rotation of
 * qr_2d.cxx:140-147 (the rank holding global block 0 is `root`) */
int64_t oracle_cyclic_global_index(int64_t loc, int64_t nb, int me, int np, int root) {
  const int mc = ((me - root) % np + np) % np;
  return (int64_t)mc * nb + loc % nb + (loc / nb) * nb * np;
}

int oracle_redistribute(int to_cyclic, int64_t m, int64_t n, int64_t nb, int nprow, int npcol, int rrow, int rcol,
                        double* const* in, double* const* out) {
  if (nprow < 1 || npcol < 1 || nb < 1 || m % (nb * nprow) || n % (nb * npcol)) return -1;
  const int64_t rows = m / nprow, cols = n / npcol;
  for (int pc = 0; pc < npcol; pc++)
    for (int pr = 0; pr < nprow; pr++) {
      /* walk the CYCLIC local piece of (pr, pc); find where each element lives in the blocked layout */
      const double* cyc_in = in[pr + pc * nprow];
      double* cyc_out = out[pr + pc * nprow];
      for (int64_t c = 0; c < cols; c++) {
        const int64_t gc = oracle_cyclic_global_index(c, nb, pc, npcol, rcol);
        const int bc = (int)(gc / cols);
        for (int64_t r = 0; r < rows; r++) {
          const int64_t gr = oracle_cyclic_global_index(r, nb, pr, nprow, rrow);
          const int br = (int)(gr / rows);
          const int64_t blk_idx = (gr % rows) + (gc % cols) * rows; /* inside blocked owner (br, bc) */
          if (to_cyclic) cyc_out[r + c * rows] = in[br + bc * nprow][blk_idx];
          else out[br + bc * nprow][blk_idx] = cyc_in[r + c * rows];
        }
      }
    }
  return 0;
}
