// candmc_b200 — the pack / replication operations of the reference's distributed-matrix wrapper (alg/SE/dmatrix.cxx) on
// device memory.  A candmc_dmat_t is DMatrix without its ScaLAPACK descriptor: global extents,
// block size, leading dimension, local data pointer and the pview (row / column / world communicators + rotating roots).
//   get_mynrow / get_myncol   dmatrix.cxx:194-203      block-cyclic local extents for the current roots (host arithmetic)
//   slice                     dmatrix.cxx:367-394      sub-matrix by reference: rotated roots + moved pointer (host arithmetic)
//   get_contig                dmatrix.cxx:470-484      lda_cpy into a packed block
//   replicate_vertical/_horizontal  :268-304           MPI_Allgather of the packed blocks over the column / row communicator
//   reduce_scatter_horizontal :310-355                 data += my chunk of the sum over the row communicator (the reference's
//                                                      hand-rolled inverted butterfly + final fix-up exchange is one
//                                                      ncclReduceScatter here; the summation order differs, rounding only)
//   transpose_data            dmatrix.cxx:252-263      packed blocks swapped with the transposed grid partner
//   foldcols / foldrows       dmatrix.cxx:527-584      local block-row regrouping: nrow x ncol <-> nrow/f x ncol*f
// All data movement is HBM- or NVLink-bound: 16 B algorithmic per element for the local kernels.
#include "../../include/candmc_b200.h"
#include "comm.h"
#include "common.cuh"
#include "runtime.h"
#include "staging.h"

namespace candmc {
namespace {

// dmatrix.cxx:194-203
int64_t local_extent(int64_t n, int64_t b, int np, int rank, int root) {
  const int64_t nb = n / b;
  return (nb / np + ((nb % np) > (rank + np - root) % np ? 1 : 0)) * b;
}

int check_dmat(const candmc_dmat_t* A, const char* what, int64_t* mr, int64_t* mc) {
  CANDMC_CHECK(A != nullptr && A->pv.crow != nullptr && A->pv.ccol != nullptr, "%s: null matrix or processor view", what);
  // extents need not be multiples of b: like the reference (dmatrix.cxx:194-203) only whole blocks are distributed
  CANDMC_CHECK(A->b > 0 && A->nrow >= 0 && A->ncol >= 0, "%s: negative extent or block size", what);
  const int nprow = A->pv.ccol->size, npcol = A->pv.crow->size;
  CANDMC_CHECK(A->pv.rrow >= 0 && A->pv.rrow < nprow && A->pv.rcol >= 0 && A->pv.rcol < npcol, "%s: root outside the grid", what);
  *mr = local_extent(A->nrow, A->b, nprow, A->pv.ccol->rank, A->pv.rrow);
  *mc = local_extent(A->ncol, A->b, npcol, A->pv.crow->rank, A->pv.rcol);
  CANDMC_CHECK(*mr == 0 || *mc == 0 || (A->data != nullptr && A->lda >= *mr), "%s: lda %lld < %lld local rows", what,
               (long long)A->lda, (long long)*mr);
  CANDMC_CHECK(is_device_ptr(A->data), "%s: data must be a device pointer", what);
  return OK;
}

// every rank of `c` must hold the same number of blocks along this axis (MPI_Allgather / the butterfly assume it)
int check_even(int64_t n, int64_t b, candmc_comm* c, const char* what) {
  CANDMC_CHECK((n / b) % c->size == 0, "%s: %lld blocks do not divide evenly over %d ranks", what, (long long)(n / b), c->size);
  return OK;
}

// packed copy of the local piece: the operand itself when it is already packed
int packed(const candmc_dmat_t* A, int64_t mr, int64_t mc, double* scratch, const double** out, cudaStream_t st) {
  if (A->lda == mr || mc <= 1) {
    *out = A->data;
    return OK;
  }
  CANDMC_TRY(lda_copy_f64(mr, mc, A->lda, mr, A->data, scratch, st));
  *out = scratch;
  return OK;
}

int replicate(const candmc_dmat_t* A, candmc_comm* c, int64_t n_axis, double* rep, cudaStream_t st, const char* what) {
  int64_t mr, mc;
  CANDMC_TRY(check_dmat(A, what, &mr, &mc));
  CANDMC_TRY(check_even(n_axis, A->b, c, what));
  CANDMC_CHECK(rep != nullptr && is_device_ptr(rep), "%s: the output must be a device pointer", what);
  const int64_t mine = mr * mc;
  if (mine == 0) return OK;
  void* ws = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * mine, &ws));
  const double* src = nullptr;
  CANDMC_TRY(packed(A, mr, mc, static_cast<double*>(ws), &src, st));
  if (c->size == 1) {
    CANDMC_CUDA(cudaMemcpyAsync(rep, src, sizeof(double) * mine, cudaMemcpyDeviceToDevice, st));
    return OK;
  }
  CANDMC_NCCL(ncclAllGather(src, rep, (size_t)mine, ncclDouble, c->nccl, st));
  return OK;
}

// Source element of output element (rr, cc) of a fold; shared by the kernel and the CPU-side test hook.
template <bool FOLDCOLS>
__host__ __device__ __forceinline__ int64_t fold_src_index(unsigned rr, int64_t cc, int64_t ocol, int64_t mc, unsigned b,
                                                           unsigned f, int64_t lda) {
  const unsigned blk = rr / b, w = rr - blk * b;
  if (FOLDCOLS) {
    const int64_t i_col = cc / mc, c_in = cc - i_col * mc;  // column group of the folded matrix, column inside it
    return (static_cast<int64_t>(blk) * f + i_col) * b + w + c_in * lda;
  }
  const unsigned j = blk / f, i = blk - j * f;
  return (static_cast<int64_t>(i) * ocol + cc) * lda + static_cast<int64_t>(j) * b + w;
}

// FOLDCOLS: out is (mr/f) x (mc*f), ld = mr/f:  out[(i*mc + c)*(mr/f) + j*b + w] = in[(j*f + i)*b + w + c*lda]
// else    : out is (mr*f) x (mc/f), ld = mr*f:  out[c*(mr*f) + (j*f + i)*b + w]  = in[(i*(mc/f) + c)*lda + j*b + w]
// grid.y strides the output columns, grid.x * block the rows of one output column (32-bit arithmetic per element).
template <bool FOLDCOLS>
__global__ void __launch_bounds__(256)
fold_kernel(const double* __restrict__ in, int64_t lda, double* __restrict__ out, unsigned orow, int64_t ocol, int64_t mc,
            unsigned b, unsigned f) {
  const unsigned xstride = gridDim.x * blockDim.x;
  for (int64_t cc = blockIdx.y; cc < ocol; cc += gridDim.y) {
    double* ocolp = out + cc * static_cast<int64_t>(orow);
    for (unsigned rr = blockIdx.x * blockDim.x + threadIdx.x; rr < orow; rr += xstride)
      ocolp[rr] = __ldg(in + fold_src_index<FOLDCOLS>(rr, cc, ocol, mc, b, f, lda));
  }
}

int fold(const candmc_dmat_t* A, int64_t f, double* out, bool cols, cudaStream_t st) {
  const char* what = cols ? "foldcols" : "foldrows";
  int64_t mr, mc;
  CANDMC_TRY(check_dmat(A, what, &mr, &mc));
  CANDMC_CHECK(f >= 1 && f < (1LL << 31), "%s: factor must be positive", what);
  CANDMC_CHECK(cols ? mr % (A->b * f) == 0 : mc % f == 0, "%s: local extent not divisible by the factor", what);
  CANDMC_CHECK(out != nullptr && is_device_ptr(out) && out != A->data, "%s: needs a distinct device output", what);
  if (mr == 0 || mc == 0) return OK;
  const int64_t orow = cols ? mr / f : mr * f, ocol = cols ? mc * f : mc / f;
  CANDMC_CHECK(orow < (1LL << 31) && A->b < (1LL << 31), "%s: more than 2^31-1 local rows", what);
  const int64_t cap = static_cast<int64_t>(runtime().num_sms) * 8;
  int64_t gx = (orow + 255) / 256;
  if (gx > cap) gx = cap;
  int64_t gy = cap / gx;
  if (gy < 1) gy = 1;
  if (gy > ocol) gy = ocol;
  if (gy > 65535) gy = 65535;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  if (cols) fold_kernel<true><<<grid, 256, 0, st>>>(A->data, A->lda, out, (unsigned)orow, ocol, mc, (unsigned)A->b, (unsigned)f);
  else fold_kernel<false><<<grid, 256, 0, st>>>(A->data, A->lda, out, (unsigned)orow, ocol, mc, (unsigned)A->b, (unsigned)f);
  CANDMC_CUDA(cudaGetLastError());
  ++runtime().launches;
  return OK;
}

}  // namespace
}  // namespace candmc

using namespace candmc;

extern "C" {

int candmc_dmat_local_extents(const candmc_dmat_t* A, int64_t* mynrow, int64_t* myncol) {
  CANDMC_CHECK(A != nullptr && A->pv.crow != nullptr && A->pv.ccol != nullptr && A->b > 0 && mynrow && myncol,
               "get_mynrow: null argument");
  *mynrow = local_extent(A->nrow, A->b, A->pv.ccol->size, A->pv.ccol->rank, A->pv.rrow);
  *myncol = local_extent(A->ncol, A->b, A->pv.crow->size, A->pv.crow->rank, A->pv.rcol);
  return OK;
}

int candmc_dmat_slice(const candmc_dmat_t* A, int64_t firstrow, int64_t numrows, int64_t firstcol, int64_t numcols,
                      candmc_dmat_t* out) {
  CANDMC_CHECK(A != nullptr && out != nullptr && A->pv.crow != nullptr && A->pv.ccol != nullptr && A->b > 0, "slice: null argument");
  CANDMC_CHECK(firstrow >= 0 && firstcol >= 0 && firstrow % A->b == 0 && firstcol % A->b == 0,
               "slice: the corner must sit on a block boundary");  // LIBT_ASSERT, dmatrix.cxx:380-381
  CANDMC_CHECK(numrows >= 0 && numcols >= 0 && firstrow + numrows <= A->nrow && firstcol + numcols <= A->ncol,
               "slice: outside the matrix");
  const int nprow = A->pv.ccol->size, npcol = A->pv.crow->size;
  candmc_dmat_t rest = *A;  // everything below / right of the corner, with the roots rotated to the corner's owner
  rest.pv.rrow = (int)((A->pv.rrow + firstrow / A->b) % nprow);
  rest.pv.rcol = (int)((A->pv.rcol + firstcol / A->b) % npcol);
  rest.nrow = A->nrow - firstrow;
  rest.ncol = A->ncol - firstcol;
  const int64_t mr0 = local_extent(A->nrow, A->b, nprow, A->pv.ccol->rank, A->pv.rrow);
  const int64_t mc0 = local_extent(A->ncol, A->b, npcol, A->pv.crow->rank, A->pv.rcol);
  const int64_t mr1 = local_extent(rest.nrow, A->b, nprow, A->pv.ccol->rank, rest.pv.rrow);
  const int64_t mc1 = local_extent(rest.ncol, A->b, npcol, A->pv.crow->rank, rest.pv.rcol);
  rest.data = A->data + (mr0 - mr1) + (mc0 - mc1) * A->lda;
  rest.nrow = numrows;
  rest.ncol = numcols;
  *out = rest;
  return OK;
}

int candmc_dmat_get_contig(const candmc_dmat_t* A, double* out, void* stream) {
  CANDMC_TRY(runtime_require());
  int64_t mr, mc;
  CANDMC_TRY(check_dmat(A, "get_contig", &mr, &mc));
  CANDMC_CHECK(out != nullptr && is_device_ptr(out), "get_contig: the output must be a device pointer");
  if (mr == 0 || mc == 0) return OK;
  return lda_copy_f64(mr, mc, A->lda, mr, A->data, out, static_cast<cudaStream_t>(stream));
}

int candmc_dmat_replicate_vertical(const candmc_dmat_t* A, double* rep, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(A != nullptr && A->pv.ccol != nullptr, "replicate_vertical: null argument");
  return replicate(A, A->pv.ccol, A->nrow, rep, static_cast<cudaStream_t>(stream), "replicate_vertical");
}

int candmc_dmat_replicate_horizontal(const candmc_dmat_t* A, double* rep, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(A != nullptr && A->pv.crow != nullptr, "replicate_horizontal: null argument");
  return replicate(A, A->pv.crow, A->ncol, rep, static_cast<cudaStream_t>(stream), "replicate_horizontal");
}

int candmc_dmat_reduce_scatter_horizontal(const candmc_dmat_t* A, double* cntrb, void* stream) {
  CANDMC_TRY(runtime_require());
  int64_t mr, mc;
  CANDMC_TRY(check_dmat(A, "reduce_scatter_horizontal", &mr, &mc));
  candmc_comm* crow = A->pv.crow;
  CANDMC_TRY(check_even(A->ncol, A->b, crow, "reduce_scatter_horizontal"));
  CANDMC_CHECK(cntrb != nullptr && is_device_ptr(cntrb), "reduce_scatter_horizontal: the contribution must be a device pointer");
  const int64_t mine = mr * mc;
  if (mine == 0) return OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // my chunk of the summed contributions lands in place (cntrb is scratch on return, as in the reference)
  double* chunk = cntrb + (int64_t)crow->rank * mine;
  if (crow->size > 1) CANDMC_NCCL(ncclReduceScatter(cntrb, chunk, (size_t)mine, ncclDouble, ncclSum, crow->nccl, st));
  return lda_axpby_f64(mr, mc, mr, A->lda, chunk, A->data, 1.0, 1.0, st);  // cdaxpy onto the local data, dmatrix.cxx:353
}

int candmc_dmat_transpose_data(const candmc_dmat_t* A, double* out, void* stream) {
  CANDMC_TRY(runtime_require());
  int64_t mr, mc;
  CANDMC_TRY(check_dmat(A, "transpose_data", &mr, &mc));
  CANDMC_CHECK(A->pv.cworld != nullptr, "transpose_data: needs the world communicator");
  candmc_comm *crow = A->pv.crow, *ccol = A->pv.ccol, *world = A->pv.cworld;
  CANDMC_CHECK(crow->size == ccol->size && world->size == crow->size * ccol->size, "transpose_data: needs a square grid");
  CANDMC_TRY(check_even(A->nrow, A->b, ccol, "transpose_data"));
  CANDMC_TRY(check_even(A->ncol, A->b, crow, "transpose_data"));
  CANDMC_CHECK(out != nullptr && is_device_ptr(out) && out != A->data, "transpose_data: needs a distinct device output");
  const int64_t mine = mr * mc;
  if (mine == 0) return OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* ws = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * mine, &ws));
  const double* src = nullptr;
  CANDMC_TRY(packed(A, mr, mc, static_cast<double*>(ws), &src, st));
  const int partner = crow->rank + ccol->rank * crow->size;  // dmatrix.cxx:257-259
  return comm_sendrecv(world, src, mine, partner, out, mine, partner, st);
}

// CPU-testable index map of the fold kernels (pure host code)
int candmc_debug_fold_src_index(int foldcols, int64_t mr, int64_t mc, int64_t b, int64_t f, int64_t lda, int64_t rr, int64_t cc,
                                int64_t* src) {
  CANDMC_CHECK(src != nullptr && b > 0 && f > 0 && mr >= 0 && mc >= 0, "fold index: bad arguments");
  const int64_t orow = foldcols ? mr / f : mr * f, ocol = foldcols ? mc * f : mc / f;
  CANDMC_CHECK(rr >= 0 && rr < orow && cc >= 0 && cc < ocol, "fold index: out of range");
  *src = foldcols ? fold_src_index<true>((unsigned)rr, cc, ocol, mc, (unsigned)b, (unsigned)f, lda)
                  : fold_src_index<false>((unsigned)rr, cc, ocol, mc, (unsigned)b, (unsigned)f, lda);
  return OK;
}

int candmc_dmat_foldcols(const candmc_dmat_t* A, int64_t factor, double* out, void* stream) {
  CANDMC_TRY(runtime_require());
  return fold(A, factor, out, true, static_cast<cudaStream_t>(stream));
}

int candmc_dmat_foldrows(const candmc_dmat_t* A, int64_t factor, double* out, void* stream) {
  CANDMC_TRY(runtime_require());
  return fold(A, factor, out, false, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
