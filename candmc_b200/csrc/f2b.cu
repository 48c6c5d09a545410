// candmc_b200 — trailing update of the symmetric full -> band reduction.
// Reference: one level of sym_full2band, alg/SE/full_to_band.cxx:57-239, everything after the panel QR (:96):
//   invT from Y                    compute_invT_from_Y, alg/QR/qr_2d/qr_2d.cxx:22-60 (root column, MPI_Bcast along the row)
//   W  = Y^T A                     cdgemm('T','N') :122, MPI_Reduce over the grid column onto the diagonal :126-130
//   Z  = Y^T W^T                   cdgemm('T','T') :157, MPI_Allreduce over the diagonal communicator :161
//   U  = Y invT^-1                 cdtrsm('R','L','N','N') :172
//   V' = W - Z U^T / 2             cdgemm('N','T', alpha = -.5, beta = 1) :186
//   V' down the columns, U along the rows   MPI_Bcast x 2 :205-207
//   UV' and its mirror image       cdgemm('N','N') :213, MPI_Sendrecv with the transposed grid partner :222-224
//   A -= UV' + (VU')               two cdaxpy per column :239-241
// The two large GEMMs (2 * 2 * mb * kb * b flop) run on the FP64 DMMA kernel; the rank-2b symmetric update is one HBM-bound
// kernel that reads the partner's block through a shared-memory transpose (32 B algorithmic per element of the trailing
// block: A read + written, UV' and VU' read once).  Communication goes through the processor view's NCCL communicators.
#include "../../include/candmc_b200.h"
#include "comm.h"
#include "common.cuh"
#include "runtime.h"
#include "staging.h"

namespace candmc {
namespace {

// T = lower triangle of S with halved diagonal, zero above (qr_2d.cxx:36-50: pack_lower, T_buf[diag] / 2, unpack_lower)
__global__ void __launch_bounds__(256)
f2b_tril_halve_kernel(const double* __restrict__ S, double* __restrict__ T, int b) {
  const int total = b * b;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int j = e / b, i = e - j * b;
    T[e] = (i > j) ? S[e] : (i == j ? 0.5 * S[e] : 0.0);
  }
}

// U L = Y for lower-triangular L (b x b, ld b): one thread per row of U, columns from the last to the first;
// consecutive threads own consecutive rows, so every access to Y and U is coalesced and L is a broadcast.
__global__ void __launch_bounds__(256)
f2b_trsm_rlnn_kernel(int64_t rows, int b, const double* __restrict__ L, const double* __restrict__ Y, int64_t ldy,
                     double* __restrict__ U, int64_t ldu) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= rows) return;
  for (int j = b - 1; j >= 0; --j) {
    double acc = Y[r + j * ldy];
    for (int i = j + 1; i < b; ++i) acc -= U[r + i * ldu] * __ldg(L + i + static_cast<int64_t>(j) * b);
    U[r + j * ldu] = acc / __ldg(L + j + static_cast<int64_t>(j) * b);
  }
}

// A(r, c) -= P(r, c) + Q(c, r): P is mb x kb (ld mb), Q is kb x mb (ld kb).  64 x 64 tiles; Q's tile is read along its own
// columns (coalesced) and turned through padded shared memory, as in transpose_f64.
constexpr int FT = 64;
__global__ void __launch_bounds__(256)
f2b_rank2_update_kernel(int64_t mb, int64_t kb, double* __restrict__ A, int64_t lda, const double* __restrict__ P,
                        const double* __restrict__ Q, int64_t tiles_r, int64_t tiles_c) {
  __shared__ double tile[FT][FT + 1];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  const int64_t ntiles = tiles_r * tiles_c;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t r0 = (t % tiles_r) * FT, c0 = (t / tiles_r) * FT;
    // Q tile: rows c0.. of Q (contiguous), columns r0..
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t qc = c0 + tx, qr = r0 + ty + 4 * i;   // Q(qc, qr) = Q[qc + qr * kb]
      tile[ty + 4 * i][tx] = (qc < kb && qr < mb) ? __ldg(Q + qc + qr * kb) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t r = r0 + tx, c = c0 + ty + 4 * i;
      if (r < mb && c < kb) A[r + c * lda] -= __ldg(P + r + c * mb) + tile[tx][ty + 4 * i];
    }
    __syncthreads();
  }
}

int launch_grid(int64_t work, int per_cta) {
  int64_t g = (work + per_cta - 1) / per_cta;
  const int64_t cap = static_cast<int64_t>(runtime().num_sms) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace candmc

using namespace candmc;

extern "C" {

int candmc_sym_full2band_extents(int64_t n, int64_t b, int64_t b_sub, int np, int myrow, int mycol, int rrow, int rcol,
                                 int64_t* loc_row_offset, int64_t* loc_col_offset, int64_t* mb, int64_t* kb) {
  CANDMC_CHECK(np > 0 && b_sub > 0 && b >= b_sub && b % b_sub == 0 && n >= b, "sym_full2band: b must be a multiple of b_sub");
  CANDMC_CHECK(myrow >= 0 && myrow < np && mycol >= 0 && mycol < np && rrow >= 0 && rrow < np && rcol >= 0 && rcol < np,
               "sym_full2band: coordinates outside the %d x %d grid", np, np);
  const int64_t s = b / b_sub, t = (n - b) / b_sub;
  // full_to_band.cxx:57-79, same expressions (C remainder semantics included)
  int64_t ro = b_sub * (b / (b_sub * np));
  if ((myrow + np - rrow) % np < s % np) ro += b_sub;
  int64_t co = b_sub * (b / (b_sub * np));
  if ((mycol + np - rcol) % np < s % np) co += b_sub;
  int64_t m = t / np;
  if ((myrow + np - rrow - (s % np)) % np < t % np) m++;
  int64_t k = t / np;
  if ((mycol + np - rcol - (s % np)) % np < t % np) k++;
  if (loc_row_offset) *loc_row_offset = ro;
  if (loc_col_offset) *loc_col_offset = co;
  if (mb) *mb = m * b_sub;
  if (kb) *kb = k * b_sub;
  return OK;
}

int candmc_sym_full2band_update(double* A, int64_t lda_A, int64_t n, int64_t b, int64_t b_sub, const candmc_pview_t* pv,
                                candmc_comm_t* cdiag, const double* Y, int64_t lda_Y, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(pv != nullptr && pv->crow != nullptr && pv->ccol != nullptr && pv->cworld != nullptr,
               "sym_full2band: null processor view");
  candmc_comm *crow = pv->crow, *ccol = pv->ccol, *world = pv->cworld;
  const int np = ccol->size, myrow = ccol->rank, mycol = crow->rank;
  CANDMC_CHECK(crow->size == np && world->size == np * np, "sym_full2band: needs a square processor grid");  // :49
  CANDMC_CHECK(b_sub > 0 && b >= b_sub && b % b_sub == 0, "sym_full2band: b must be a multiple of b_sub");     // :52
  if (n <= b) return OK;                                                                                      // :35
  // the reference's offsets and the mb == kb it asserts on the diagonal (:139) hold only for whole rounds of blocks;
  // outside that its own run fails (heap corruption on 2 x 2 with b / b_sub = 3, observed) — refuse instead
  CANDMC_CHECK((b / b_sub) % np == 0 && ((n - b) / b_sub) % np == 0 && (n - b) % b_sub == 0,
               "sym_full2band: b / b_sub and (n - b) / b_sub must be multiples of the grid dimension %d", np);
  CANDMC_CHECK(b < (1 << 15), "sym_full2band: band width too large");
  const bool diag = (myrow == mycol);
  CANDMC_CHECK(!diag || np == 1 || (cdiag != nullptr && cdiag->size == np), "sym_full2band: diagonal ranks need the diagonal communicator");
  int64_t ro, co, mb, kb;
  CANDMC_TRY(candmc_sym_full2band_extents(n, b, b_sub, np, myrow, mycol, pv->rrow, pv->rcol, &ro, &co, &mb, &kb));
  CANDMC_CHECK(is_device_ptr(A) && is_device_ptr(Y), "sym_full2band: operands must be device pointers");
  CANDMC_CHECK(mb == 0 || (Y != nullptr && lda_Y >= mb && A != nullptr && lda_A >= ro + mb), "sym_full2band: bad leading dimension");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* At = A + ro + co * lda_A;  // trailing block, :122

  void* wsv = nullptr;
  const int64_t bb = b * b, bbp = bb + (bb & 1);  // every piece starts on a 16-byte boundary (TMA path of the GEMM)
  CANDMC_TRY(workspace_get(sizeof(double) * (3 * bbp + b * kb + mb * b + 2 * mb * kb + 16), &wsv));
  double* S = static_cast<double*>(wsv);
  double* T = S + bbp;
  double* Z = T + bbp;
  double* W = Z + bbp;                   // b x kb, becomes V'
  double* U = W + b * kb + (b * kb & 1); // mb x b
  double* UVT = U + mb * b + (mb * b & 1);
  double* VUT = UVT + mb * kb + (mb * kb & 1);

  // ---- invT (replicated), qr_2d.cxx:22-60 with the rotated root row — the root column forms it, the row broadcast ships it
  if (mycol == pv->rcol) {
    if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, b, mb, 1.0, Y, lda_Y, Y, lda_Y, 0.0, S, b, st));
    else CANDMC_TRY(fill_f64(S, bb, 0.0, st));
    if (np > 1) CANDMC_TRY(comm_allreduce(ccol, S, S, bb, st));
    f2b_tril_halve_kernel<<<launch_grid(bb, 256), 256, 0, st>>>(S, T, (int)b);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
  }
  CANDMC_TRY(comm_bcast(crow, T, T, bb, pv->rcol, st));

  // ---- W = Y^T A summed over the grid column (:122-131; the all-reduce leaves it on the diagonal rank too)
  if (kb > 0) {
    if (mb > 0) CANDMC_TRY(gemm_f64('T', 'N', b, kb, mb, 1.0, Y, lda_Y, At, lda_A, 0.0, W, b, st));
    else CANDMC_TRY(fill_f64(W, b * kb, 0.0, st));
    if (np > 1) CANDMC_TRY(comm_allreduce(ccol, W, W, b * kb, st));
  }

  // ---- on the diagonal: Z, U, V' (:137-190)
  if (diag) {
    const int64_t lb = mb;
    CANDMC_CHECK(mb == kb, "sym_full2band: diagonal rank with mb != kb");
    if (lb > 0) CANDMC_TRY(gemm_f64('T', 'T', b, b, lb, 1.0, Y, lda_Y, W, b, 0.0, Z, b, st));
    else CANDMC_TRY(fill_f64(Z, bb, 0.0, st));
    if (np > 1) CANDMC_TRY(comm_allreduce(cdiag, Z, Z, bb, st));
    if (lb > 0) {
      f2b_trsm_rlnn_kernel<<<(int)((lb + 255) / 256), 256, 0, st>>>(lb, (int)b, T, Y, lda_Y, U, lb);
      CANDMC_CUDA(cudaGetLastError());
      runtime().launches++;
      CANDMC_TRY(gemm_f64('N', 'T', b, lb, b, -0.5, Z, b, U, lb, 1.0, W, b, st));
    }
  }
  // ---- V' down the grid column from the diagonal rank, U along the grid row from the diagonal rank (:205-207)
  if (kb > 0) CANDMC_TRY(comm_bcast(ccol, W, W, b * kb, mycol, st));
  if (mb > 0) CANDMC_TRY(comm_bcast(crow, U, U, mb * b, myrow, st));

  // ---- UV', its mirror image from the transposed partner, and the rank-2b update (:209-243)
  if (mb > 0 && kb > 0) {
    CANDMC_TRY(gemm_f64('N', 'N', mb, kb, b, 1.0, U, mb, W, b, 0.0, UVT, mb, st));
    const double* mirror = UVT;
    if (!diag) {
      const int partner = mycol + myrow * np;  // rank = row + col * np, so this is (row = mycol, col = myrow), :55
      CANDMC_TRY(comm_sendrecv(world, UVT, mb * kb, partner, VUT, kb * mb, partner, st));
      mirror = VUT;
    }
    const int64_t tr = (mb + FT - 1) / FT, tc = (kb + FT - 1) / FT;
    f2b_rank2_update_kernel<<<launch_grid(tr * tc, 1), 256, 0, st>>>(mb, kb, At, lda_A, UVT, mirror, tr, tc);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
  }
  return OK;
}

}  // extern "C"
