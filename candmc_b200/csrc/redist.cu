// candmc_b200 — block-cyclic <-> blocked redistribution on the processor grid (BASELINE.json
// north_star: "cyclic/blocked layout redistribution and packing are coalesced, vectorised gather kernels").
//
// It lets the blocked multiplies (SUMMA / 2.5D / Cannon) consume and produce the ScaLAPACK-style block-cyclic matrices the
// reference's QR / SE drivers hold (test/QR/test_qr_2d.cxx:87-94, alg/SE/dmatrix.cxx:194-203).  Index plan: redist.h.
//   cyclic -> blocked:  rows over `ccol`:  pack contiguous row ranges -> all-to-all -> scatter block rows (stride nprow)
//                       cols over `crow`:  contiguous column ranges go out straight from the intermediate matrix
//                                          -> all-to-all -> scatter block columns (stride npcol) into the destination
//   blocked -> cyclic:  the same two exchanges run backwards (gather kernels first, contiguous copies last).
// Kernels are pure HBM streaming: 16 B algorithmic per element (8 read + 8 written), double2 accesses when nb, the local
// extents and the leading dimensions are even; coalesced along the column-major rows on both sides.
// Exchanges are grouped ncclSend/ncclRecv on the full-width axis communicators (the self segment is a device copy).
#include "redist.h"

#include "../../include/candmc_b200.h"
#include "comm.h"
#include "common.cuh"
#include "runtime.h"

namespace candmc {
namespace {

constexpr int RD_THREADS = 256;

// X is the local rows x cols matrix (leading dimension ldx) in BLOCKED order along the permuted axis; SEG is the
// segmented exchange buffer.  GATHER: SEG <- X (blocked -> cyclic, before the exchange); otherwise X <- SEG.
// VEC = 2 moves double2 (requires even nb / rows / ldx and 16 B aligned bases).
// grid.y strides the columns, grid.x * block the rows of a column; everything per element is 32-bit arithmetic (two
// divisions by nb and P when block rows are permuted, none when block columns are: their terms are per column).
template <bool ROWS_AXIS, bool GATHER, int VEC>
__global__ void __launch_bounds__(RD_THREADS)
permute_blocks_kernel(AxisPlan pl, double* __restrict__ X, int64_t ldx, double* __restrict__ SEG, unsigned rows,
                      int64_t cols) {
  const unsigned rv = rows / VEC;
  const unsigned nb = (unsigned)pl.nb;
  const unsigned xstride = gridDim.x * blockDim.x;
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    int64_t col_base = 0;  // block columns: position of (row 0, column c) inside the segments
    if (!ROWS_AXIS) col_base = strided_segment_index(pl, false, c / nb, (int)(c % nb), 0, rows);
    double* xcol = X + c * ldx;
    for (unsigned r2 = blockIdx.x * blockDim.x + threadIdx.x; r2 < rv; r2 += xstride) {
      const unsigned r = r2 * VEC;
      int64_t idx;
      if (ROWS_AXIS) {
        const unsigned blk = r / nb, w = r - blk * nb;
        int p;
        unsigned t;
        strided_owner(pl, blk, &p, &t);
        const int64_t seg_rows = (int64_t)pl.scnt[p] * nb;
        idx = pl.soff[p] * nb * cols + c * seg_rows + (int64_t)t * nb + w;
      } else {
        idx = col_base + r;
      }
      double* xp = xcol + r;
      double* sp = SEG + idx;
      if (VEC == 2) {
        if (GATHER) *reinterpret_cast<double2*>(sp) = *reinterpret_cast<const double2*>(xp);
        else *reinterpret_cast<double2*>(xp) = *reinterpret_cast<const double2*>(sp);
      } else {
        if (GATHER) *sp = *xp;
        else *xp = *sp;
      }
    }
  }
}

template <bool ROWS_AXIS, bool GATHER>
int launch_permute(const AxisPlan& pl, double* X, int64_t ldx, double* SEG, int64_t rows, int64_t cols,
                   cudaStream_t st) {
  if (rows == 0 || cols == 0) return OK;
  CANDMC_CHECK(rows < (1LL << 31), "redistribute: more than 2^31-1 local rows");
  const bool vec = pl.nb % 2 == 0 && rows % 2 == 0 && ldx % 2 == 0 && reinterpret_cast<uintptr_t>(X) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(SEG) % 16 == 0;
  const int64_t rv = vec ? rows / 2 : rows;
  // about 8 CTAs per SM in total: as many along a column as it has work for, the rest across columns
  const int64_t cap = static_cast<int64_t>(runtime().num_sms) * 8;
  int64_t gx = (rv + RD_THREADS - 1) / RD_THREADS;
  if (gx > cap) gx = cap;
  int64_t gy = cap / gx;
  if (gy < 1) gy = 1;
  if (gy > cols) gy = cols;
  if (gy > 65535) gy = 65535;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  if (vec) permute_blocks_kernel<ROWS_AXIS, GATHER, 2><<<grid, RD_THREADS, 0, st>>>(pl, X, ldx, SEG, (unsigned)rows, cols);
  else permute_blocks_kernel<ROWS_AXIS, GATHER, 1><<<grid, RD_THREADS, 0, st>>>(pl, X, ldx, SEG, (unsigned)rows, cols);
  CANDMC_CUDA(cudaGetLastError());
  ++runtime().launches;
  return OK;
}

// Grouped exchange: segment p of `send` (soff[p], scnt[p] doubles) goes to rank p, segment p of `recv` comes from rank p.
int alltoallv(candmc_comm* c, const double* send, const int64_t* soff, const int64_t* scnt, double* recv,
              const int64_t* roff, const int64_t* rcnt, cudaStream_t st) {
  const int me = c->rank;
  CANDMC_CHECK(scnt[me] == rcnt[me], "redistribute: self segment sizes differ");
  if (scnt[me] > 0)
    CANDMC_CUDA(cudaMemcpyAsync(recv + roff[me], send + soff[me], sizeof(double) * scnt[me], cudaMemcpyDeviceToDevice, st));
  if (c->size == 1) return OK;
  CANDMC_NCCL(ncclGroupStart());
  for (int p = 0; p < c->size; ++p) {
    if (p == me) continue;
    if (scnt[p] > 0) CANDMC_NCCL(ncclSend(send + soff[p], (size_t)scnt[p], ncclDouble, p, c->nccl, st));
    if (rcnt[p] > 0) CANDMC_NCCL(ncclRecv(recv + roff[p], (size_t)rcnt[p], ncclDouble, p, c->nccl, st));
  }
  CANDMC_NCCL(ncclGroupEnd());
  return OK;
}

// One axis of the redistribution.  in: rows x cols (ld_in), out: rows x cols (ld_out); `scratch` holds 2*rows*cols doubles.
// to_blocked: `in` is cyclic along this axis and `out` blocked; otherwise the reverse.
int axis_exchange(bool rows_axis, bool to_blocked, const AxisPlan& pl, candmc_comm* comm, const double* in, int64_t ld_in,
                  double* out, int64_t ld_out, int64_t rows, int64_t cols, double* scratch, cudaStream_t st) {
  const int P = pl.P;
  if (P == 1) return lda_copy_f64(rows, cols, ld_in, ld_out, in, out, st);
  const int64_t other = rows_axis ? cols : rows;  // extent of the axis that is not permuted
  const int64_t per_block = (int64_t)pl.nb * other;
  double* sbuf = scratch;
  double* rbuf = scratch + rows * cols;
  int64_t c_off[REDIST_MAX_P] = {0}, c_cnt[REDIST_MAX_P] = {0}, s_off[REDIST_MAX_P] = {0}, s_cnt[REDIST_MAX_P] = {0};
  for (int p = 0; p < P; ++p) {
    c_off[p] = pl.coff[p] * per_block;
    c_cnt[p] = pl.ccnt[p] * per_block;
    s_off[p] = pl.soff[p] * per_block;
    s_cnt[p] = pl.scnt[p] * per_block;
  }
  if (to_blocked) {
    // contiguous ranges of the cyclic input -> send segments
    const double* send = sbuf;
    if (!rows_axis && ld_in == rows) {
      send = in;  // whole column ranges of a packed matrix are already contiguous: segment p starts at column lo[p]*nb
      for (int p = 0; p < P; ++p) c_off[p] = (int64_t)pl.lo[p] * per_block;
    } else {
      for (int p = 0; p < P; ++p) {
        if (pl.ccnt[p] == 0) continue;
        const int64_t ext = (int64_t)pl.ccnt[p] * pl.nb, at = (int64_t)pl.lo[p] * pl.nb;
        if (rows_axis) CANDMC_TRY(lda_copy_f64(ext, cols, ld_in, ext, in + at, sbuf + c_off[p], st));
        else CANDMC_TRY(lda_copy_f64(rows, ext, ld_in, rows, in + at * ld_in, sbuf + c_off[p], st));
      }
    }
    CANDMC_TRY(alltoallv(comm, send, c_off, c_cnt, rbuf, s_off, s_cnt, st));
    if (rows_axis) return launch_permute<true, false>(pl, out, ld_out, rbuf, rows, cols, st);
    return launch_permute<false, false>(pl, out, ld_out, rbuf, rows, cols, st);
  }
  // blocked input -> strided gather into send segments -> exchange -> contiguous ranges of the cyclic output
  if (rows_axis) CANDMC_TRY((launch_permute<true, true>(pl, const_cast<double*>(in), ld_in, sbuf, rows, cols, st)));
  else CANDMC_TRY((launch_permute<false, true>(pl, const_cast<double*>(in), ld_in, sbuf, rows, cols, st)));
  if (!rows_axis && ld_out == rows) {
    for (int p = 0; p < P; ++p) c_off[p] = (int64_t)pl.lo[p] * per_block;
    return alltoallv(comm, sbuf, s_off, s_cnt, out, c_off, c_cnt, st);
  }
  CANDMC_TRY(alltoallv(comm, sbuf, s_off, s_cnt, rbuf, c_off, c_cnt, st));
  for (int p = 0; p < P; ++p) {
    if (pl.ccnt[p] == 0) continue;
    const int64_t ext = (int64_t)pl.ccnt[p] * pl.nb, at = (int64_t)pl.lo[p] * pl.nb;
    if (rows_axis) CANDMC_TRY(lda_copy_f64(ext, cols, ext, ld_out, rbuf + c_off[p], out + at, st));
    else CANDMC_TRY(lda_copy_f64(rows, ext, rows, ld_out, rbuf + c_off[p], out + at * ld_out, st));
  }
  return OK;
}

}  // namespace
}  // namespace candmc

using namespace candmc;

extern "C" {

int candmc_redistribute(int to_cyclic, int64_t m, int64_t n, int64_t nb, const double* src, int64_t ld_src, double* dst,
                        int64_t ld_dst, const candmc_pview_t* pv, void* stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(pv != nullptr && pv->crow != nullptr && pv->ccol != nullptr, "redistribute: null processor view");
  candmc_comm* crow = pv->crow;  // ranks of my grid row: rank = my column, size = npcol
  candmc_comm* ccol = pv->ccol;  // ranks of my grid column: rank = my row, size = nprow
  const int nprow = ccol->size, npcol = crow->size;
  CANDMC_CHECK(nprow <= REDIST_MAX_P && npcol <= REDIST_MAX_P, "redistribute: grid axis larger than %d", REDIST_MAX_P);
  CANDMC_CHECK(m >= 0 && n >= 0 && nb >= 1 && nb < (1 << 30), "redistribute: bad extents m=%lld n=%lld nb=%lld", (long long)m,
               (long long)n, (long long)nb);
  CANDMC_CHECK(m % (nb * nprow) == 0 && n % (nb * npcol) == 0,
               "redistribute: m (%lld) and n (%lld) must be multiples of nb*nprow (%lld) and nb*npcol (%lld)", (long long)m,
               (long long)n, (long long)(nb * nprow), (long long)(nb * npcol));
  CANDMC_CHECK(pv->rrow >= 0 && pv->rrow < nprow && pv->rcol >= 0 && pv->rcol < npcol, "redistribute: root outside the grid");
  const int64_t rows = m / nprow, cols = n / npcol;
  if (rows == 0 || cols == 0) return OK;
  CANDMC_CHECK(src != nullptr && dst != nullptr && src != dst, "redistribute: needs distinct source and destination");
  CANDMC_CHECK(ld_src >= rows && ld_dst >= rows, "redistribute: leading dimension smaller than the %lld local rows",
               (long long)rows);
  AxisPlan prow, pcol;
  CANDMC_CHECK(axis_plan(nprow, ccol->rank, pv->rrow, rows / nb, (int)nb, &prow), "redistribute: bad row plan");
  CANDMC_CHECK(axis_plan(npcol, crow->rank, pv->rcol, cols / nb, (int)nb, &pcol), "redistribute: bad column plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* ws = nullptr;
  CANDMC_TRY(workspace_get(sizeof(double) * 3 * rows * cols, &ws));
  double* mid = static_cast<double*>(ws);  // rows x cols, ld = rows: blocked along one axis, cyclic along the other
  double* scratch = mid + rows * cols;
  if (!to_cyclic) {
    CANDMC_TRY(axis_exchange(true, true, prow, ccol, src, ld_src, mid, rows, rows, cols, scratch, st));
    CANDMC_TRY(axis_exchange(false, true, pcol, crow, mid, rows, dst, ld_dst, rows, cols, scratch, st));
  } else {
    CANDMC_TRY(axis_exchange(false, false, pcol, crow, src, ld_src, mid, rows, rows, cols, scratch, st));
    CANDMC_TRY(axis_exchange(true, false, prow, ccol, mid, rows, dst, ld_dst, rows, cols, scratch, st));
  }
  return OK;
}

// Test hook: runs ONE permute kernel for the plan of rank `me` of `P` (any P, no communicator needed), so that a single GPU
// can play every rank of a grid axis in turn (tests/redist_worker.py).  gather != 0: SEG <- X, else X <- SEG.
int candmc_debug_redist_permute(int P, int me, int root, int64_t K, int nb, int rows_axis, int gather, double* X,
                                int64_t ldx, double* SEG, int64_t rows, int64_t cols, void* stream) {
  CANDMC_TRY(runtime_require());
  AxisPlan pl;
  CANDMC_CHECK(axis_plan(P, me, root, K, nb, &pl), "redist plan: bad arguments");
  CANDMC_CHECK((rows_axis ? rows : cols) == K * nb && ldx >= rows, "redist permute: extents do not match the plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows_axis) return gather ? launch_permute<true, true>(pl, X, ldx, SEG, rows, cols, st)
                               : launch_permute<true, false>(pl, X, ldx, SEG, rows, cols, st);
  return gather ? launch_permute<false, true>(pl, X, ldx, SEG, rows, cols, st)
                : launch_permute<false, false>(pl, X, ldx, SEG, rows, cols, st);
}

// ---- pure host helpers for the CPU tests (no GPU needed) ------------------------------------------------------------
int candmc_redist_axis_plan(int P, int me, int root, int64_t K, int nb, int* lo, int* ccnt, int* first, int* scnt) {
  AxisPlan pl;
  CANDMC_CHECK(axis_plan(P, me, root, K, nb, &pl), "redist plan: bad arguments");
  for (int p = 0; p < P; ++p) {
    lo[p] = pl.lo[p];
    ccnt[p] = pl.ccnt[p];
    first[p] = pl.first[p];
    scnt[p] = pl.scnt[p];
  }
  return OK;
}

int candmc_redist_strided_index(int P, int me, int root, int64_t K, int nb, int rows_axis, int64_t blk, int w, int64_t o,
                                int64_t other, int64_t* idx, int* peer) {
  AxisPlan pl;
  CANDMC_CHECK(axis_plan(P, me, root, K, nb, &pl), "redist plan: bad arguments");
  CANDMC_CHECK(blk >= 0 && blk < K && w >= 0 && w < nb && o >= 0 && o < other, "redist index: out of range");
  *idx = strided_segment_index(pl, rows_axis != 0, blk, w, o, other);
  *peer = (int)(((int64_t)me * K + blk + root) % P);
  return OK;
}

}  // extern "C"
