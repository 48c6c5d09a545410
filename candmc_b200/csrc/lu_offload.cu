// candmc_b200 — accelerator seam of the reference's 2.5D LU.
//
// The reference keeps three "offloaded" matrices on an accelerator (enum OFF_MAT {OFF_A, OFF_L, OFF_U},
// alg/LU/lu_offload.h:19) and drives them with a handful of calls: allocate / free, strided upload / download
// (upload_lda_cpy / download_lda_cpy, lu_offload.cxx:338-392), row gather / scatter / swap over PCI
// (offload_sparse_rw, lu_offload.cxx:424-476), a GEMM between sub-blocks addressed by (matrix, element offset)
// (offload_gemm_A, lu_offload.cxx:216-251) that may run asynchronously, and wait_gemm (lu_offload.cxx:130-134).
// Its accelerator was a Xeon Phi; without one the calls fall back to host memcpy + dgemm_.  Here the three matrices
// live in HBM, the GEMM is gemm_f64 (TMA + DMMA, gemm_f64.cu) on its own stream, and transfers run on a second
// stream so that the host-side panel work and PCIe traffic overlap the trailing-matrix update.
//
// Ordering: calls are issued by one host thread.  GEMMs execute in issue order on `gemm_stream`.  Transfers execute
// in issue order on `xfer_stream`.  Between the two streams a small scoreboard keeps program order wherever two
// operations touch overlapping elements (a transfer that writes a block an in-flight GEMM reads or writes, a
// transfer that reads a block an in-flight GEMM writes); a GEMM waits for every transfer issued before it.  The
// result is what a fully serial execution produces — the reference's host fallback — while non-conflicting
// transfers proceed under a running GEMM.  candmc_off_set_overlap(0) puts everything on one stream.
//
// Host-visible completion: downloads and sparse reads return with the data in the caller's buffer; uploads return
// once the caller's buffer may be reused (pageable sources are staged by the driver before the call returns).
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include <vector>

#include "../../include/candmc_b200.h"
#include "common.cuh"
#include "runtime.h"
#include "staging.h"

namespace candmc {
namespace {

constexpr int NMAT = 3;

struct OffMat {
  double* dev = nullptr;     // HBM
  int64_t size = 0;          // doubles
  double* mirror = nullptr;  // pinned host mirror handed out by candmc_off_host_mirror (lazily allocated)
  bool mirror_dirty = false; // the host may have written the mirror since the last flush
};

// A strided block [rows x cols] at element offset `off`, leading dimension `ld`, of offloaded matrix `mat`.
struct Block {
  int mat;
  int64_t off, ld, rows, cols;
  int64_t lo() const { return off; }
  int64_t hi() const { return (rows == 0 || cols == 0) ? off : off + (cols - 1) * ld + rows; }  // one past the last element
};

struct InflightGemm {
  Block a, b, c;
  cudaEvent_t done;
};

struct OffState {
  OffMat mat[NMAT];
  cudaStream_t gemm_stream = nullptr, xfer_stream = nullptr;
  bool overlap = true;
  std::vector<InflightGemm> inflight;
  std::vector<cudaEvent_t> free_events;
  cudaEvent_t xfer_tail = nullptr;  // re-recorded before every GEMM
  // grow-only scratch for sparse row traffic: device rows, pinned host rows, device + pinned offsets
  double* d_rows = nullptr;
  double* h_rows = nullptr;
  int64_t rows_cap = 0;
  int64_t* d_offs = nullptr;
  int64_t* h_offs = nullptr;
  int64_t offs_cap = 0;
  int64_t transfer_hint = 0;
  int64_t n_gemm = 0, n_upload = 0, n_download = 0, n_sparse = 0, n_waits = 0;
  bool ready = false;
};

OffState g_off;

// Conservative test: do two blocks of the same matrix share an element?  Exact when both use the same leading
// dimension and neither wraps a column; otherwise falls back to the enclosing element intervals.
bool blocks_overlap(const Block& x, const Block& y) {
  if (x.mat != y.mat) return false;
  if (x.rows == 0 || x.cols == 0 || y.rows == 0 || y.cols == 0) return false;
  if (x.hi() <= y.lo() || y.hi() <= x.lo()) return false;
  if (x.ld == y.ld && x.ld > 0) {
    const int64_t xr = x.off % x.ld, xc = x.off / x.ld, yr = y.off % y.ld, yc = y.off / y.ld;
    if (xr + x.rows <= x.ld && yr + y.rows <= y.ld) {
      const bool rows_hit = xr < yr + y.rows && yr < xr + x.rows;
      const bool cols_hit = xc < yc + y.cols && yc < xc + x.cols;
      return rows_hit && cols_hit;
    }
  }
  return true;
}

int off_init() {
  if (g_off.ready) return OK;
  if (!runtime().initialized) {
    // one process per GPU: pick the device like the launchers do (LOCAL_RANK), modulo what is visible
    int dev = -1;
    const char* e = getenv("CANDMC_OFF_DEVICE");
    if (e == nullptr) e = getenv("LOCAL_RANK");
    if (e != nullptr) {
      int count = 0;
      if (cudaGetDeviceCount(&count) == cudaSuccess && count > 0) dev = atoi(e) % count;
      else cudaGetLastError();
    }
    CANDMC_TRY(runtime_init(dev));
  }
  CANDMC_CUDA(cudaStreamCreateWithFlags(&g_off.gemm_stream, cudaStreamNonBlocking));
  CANDMC_CUDA(cudaStreamCreateWithFlags(&g_off.xfer_stream, cudaStreamNonBlocking));
  CANDMC_CUDA(cudaEventCreateWithFlags(&g_off.xfer_tail, cudaEventDisableTiming));
  g_off.ready = true;
  return OK;
}

cudaStream_t xfer_stream() { return g_off.overlap ? g_off.xfer_stream : g_off.gemm_stream; }

int event_get(cudaEvent_t* ev) {
  if (!g_off.free_events.empty()) {
    *ev = g_off.free_events.back();
    g_off.free_events.pop_back();
    return OK;
  }
  CANDMC_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  return OK;
}

// Drops GEMMs that have finished from the scoreboard.
int retire_finished() {
  size_t keep = 0;
  for (size_t i = 0; i < g_off.inflight.size(); ++i) {
    cudaError_t q = cudaEventQuery(g_off.inflight[i].done);
    if (q == cudaSuccess) {
      g_off.free_events.push_back(g_off.inflight[i].done);
    } else if (q == cudaErrorNotReady) {
      g_off.inflight[keep++] = g_off.inflight[i];
    } else {
      set_last_error("lu_offload: cudaEventQuery -> %s", cudaGetErrorString(q));
      return ERR_CUDA;
    }
  }
  g_off.inflight.resize(keep);
  return OK;
}

// Makes the transfer stream wait for the in-flight GEMMs whose operands conflict with an access to `blk`.
int order_transfer_after_gemms(const Block& blk, bool transfer_writes) {
  if (!g_off.overlap) return OK;  // single stream: already ordered
  CANDMC_TRY(retire_finished());
  for (const InflightGemm& g : g_off.inflight) {
    const bool conflict = blocks_overlap(blk, g.c) ||
                          (transfer_writes && (blocks_overlap(blk, g.a) || blocks_overlap(blk, g.b)));
    if (conflict) {
      CANDMC_CUDA(cudaStreamWaitEvent(g_off.xfer_stream, g.done, 0));
      ++g_off.n_waits;
    }
  }
  return OK;
}

int check_mat(int mat, const char* what) {
  CANDMC_CHECK(mat >= 0 && mat < NMAT, "%s: unknown offloaded matrix %d", what, mat);
  CANDMC_CHECK(g_off.mat[mat].dev != nullptr, "%s: offloaded matrix %d is not allocated", what, mat);
  return OK;
}

int check_block(const Block& b, const char* what) {
  CANDMC_TRY(check_mat(b.mat, what));
  CANDMC_CHECK(b.rows >= 0 && b.cols >= 0 && b.off >= 0, "%s: negative extent or offset", what);
  if (b.rows == 0 || b.cols == 0) return OK;
  CANDMC_CHECK(b.ld >= b.rows || b.cols == 1, "%s: leading dimension %lld smaller than %lld rows", what,
               (long long)b.ld, (long long)b.rows);
  CANDMC_CHECK(b.hi() <= g_off.mat[b.mat].size, "%s: block [%lld,%lld) exceeds the %lld doubles of matrix %d", what,
               (long long)b.lo(), (long long)b.hi(), (long long)g_off.mat[b.mat].size, b.mat);
  return OK;
}

// If the host obtained the mirror of `mat` (get_mat_handle) it may have written it: push it to HBM before anything
// else touches the matrix.  The reference's only use is the initial memcpy of A (lu_25d_pvt.cxx:1600-1602).
int flush_mirror(int mat) {
  OffMat& m = g_off.mat[mat];
  if (!m.mirror_dirty) return OK;
  // the upload is a transfer that writes the whole matrix
  Block whole{mat, 0, m.size, m.size, 1};
  CANDMC_TRY(order_transfer_after_gemms(whole, true));
  CANDMC_CUDA(cudaMemcpyAsync(m.dev, m.mirror, sizeof(double) * m.size, cudaMemcpyHostToDevice, xfer_stream()));
  m.mirror_dirty = false;
  return OK;
}

int rows_scratch(int64_t elems) {
  if (elems <= g_off.rows_cap) return OK;
  CANDMC_CUDA(cudaStreamSynchronize(xfer_stream()));
  if (g_off.d_rows) CANDMC_CUDA(cudaFree(g_off.d_rows));
  if (g_off.h_rows) CANDMC_CUDA(cudaFreeHost(g_off.h_rows));
  g_off.d_rows = nullptr;
  g_off.h_rows = nullptr;
  g_off.rows_cap = 0;
  const int64_t cap = std::max<int64_t>(elems, std::max<int64_t>(g_off.transfer_hint, 1 << 16));
  CANDMC_CUDA(cudaMalloc(&g_off.d_rows, sizeof(double) * cap));
  CANDMC_CUDA(cudaMallocHost(&g_off.h_rows, sizeof(double) * cap));
  g_off.rows_cap = cap;
  return OK;
}

int offs_scratch(int64_t n) {
  if (n <= g_off.offs_cap) return OK;
  CANDMC_CUDA(cudaStreamSynchronize(xfer_stream()));
  if (g_off.d_offs) CANDMC_CUDA(cudaFree(g_off.d_offs));
  if (g_off.h_offs) CANDMC_CUDA(cudaFreeHost(g_off.h_offs));
  g_off.d_offs = nullptr;
  g_off.h_offs = nullptr;
  g_off.offs_cap = 0;
  const int64_t cap = std::max<int64_t>(n, 4096);
  CANDMC_CUDA(cudaMalloc(&g_off.d_offs, sizeof(int64_t) * cap));
  CANDMC_CUDA(cudaMallocHost(&g_off.h_offs, sizeof(int64_t) * cap));
  g_off.offs_cap = cap;
  return OK;
}

// ---- sparse row kernels ---------------------------------------------------------------------------------------------
// Row i of the packed buffer (ncol contiguous doubles) <-> elements M[offs[i] + j*ld], j < ncol, of a column-major
// matrix: the packed side is coalesced, the matrix side touches one 32 B sector per element (rows are scattered by
// pivoting, so there is nothing better to coalesce on).  MODE 0 = gather (read), 1 = scatter (write), 2 = swap.
template <int MODE>
__global__ void __launch_bounds__(256)
sparse_rows_kernel(double* __restrict__ M, int64_t ld, const int64_t* __restrict__ offs, int64_t nrow, unsigned ncol,
                   double* __restrict__ rows_in_out, int64_t row0) {
  // grid.y strides the rows, grid.x * block the elements of a row
  const unsigned xstride = gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.y; i < nrow; i += gridDim.y) {
    double* mrow = M + offs[row0 + i];
    double* prow = rows_in_out + (row0 + i) * static_cast<int64_t>(ncol);
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < ncol; j += xstride) {
      double* mp = mrow + static_cast<int64_t>(j) * ld;
      if (MODE == 0) {
        prow[j] = *mp;
      } else if (MODE == 1) {
        *mp = prow[j];
      } else {
        const double old = *mp;
        *mp = prow[j];
        prow[j] = old;
      }
    }
  }
}

int launch_sparse(int mode, double* M, int64_t ld, int64_t nrow, int64_t ncol, int64_t row0, cudaStream_t s) {
  if (nrow == 0 || ncol == 0) return OK;
  CANDMC_CHECK(ncol < (1LL << 31), "offload_sparse_rw: more than 2^31-1 columns");
  const int64_t cap = static_cast<int64_t>(runtime().num_sms) * 8;
  int64_t gx = (ncol + 255) / 256;
  if (gx > cap) gx = cap;
  int64_t gy = cap / gx;
  if (gy < 1) gy = 1;
  if (gy > nrow) gy = nrow;
  if (gy > 65535) gy = 65535;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  const unsigned nc = (unsigned)ncol;
  if (mode == 0) sparse_rows_kernel<0><<<grid, 256, 0, s>>>(M, ld, g_off.d_offs, nrow, nc, g_off.d_rows, row0);
  else if (mode == 1) sparse_rows_kernel<1><<<grid, 256, 0, s>>>(M, ld, g_off.d_offs, nrow, nc, g_off.d_rows, row0);
  else sparse_rows_kernel<2><<<grid, 256, 0, s>>>(M, ld, g_off.d_offs, nrow, nc, g_off.d_rows, row0);
  CANDMC_CUDA(cudaGetLastError());
  ++runtime().launches;
  return OK;
}

int free_mat(int mat) {
  CANDMC_CHECK(mat >= 0 && mat < NMAT, "free_offload: unknown offloaded matrix %d", mat);
  OffMat& m = g_off.mat[mat];
  if (m.dev == nullptr && m.mirror == nullptr) return OK;
  // nothing may still be using the buffer
  if (g_off.ready) {
    CANDMC_CUDA(cudaStreamSynchronize(g_off.gemm_stream));
    CANDMC_CUDA(cudaStreamSynchronize(g_off.xfer_stream));
    CANDMC_TRY(retire_finished());
  }
  if (m.dev) CANDMC_CUDA(cudaFree(m.dev));
  if (m.mirror) CANDMC_CUDA(cudaFreeHost(m.mirror));
  m = OffMat();
  return OK;
}

}  // namespace
}  // namespace candmc

using namespace candmc;

extern "C" {

int candmc_off_set_device(int rank) {
  CANDMC_CHECK(rank >= 0, "set_mic_rank: negative rank");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    set_last_error("set_mic_rank: no CUDA device visible (the LU offload path has no CPU fallback)");
    return ERR_NODEVICE;
  }
  if (runtime().initialized) {
    // the reference calls this at the top of every factorisation (lu_25d_pvt.cxx:1569): repeating the binding is fine,
    // changing it is not (the offloaded matrices live on the bound device)
    CANDMC_CHECK(runtime().device == rank % count, "set_mic_rank: the library is already bound to device %d",
                 runtime().device);
    return OK;
  }
  return runtime_init(rank % count);
}

int candmc_off_set_overlap(int enable) {
  if (g_off.ready) {
    CANDMC_CUDA(cudaStreamSynchronize(g_off.gemm_stream));
    CANDMC_CUDA(cudaStreamSynchronize(g_off.xfer_stream));
    CANDMC_TRY(retire_finished());
  }
  g_off.overlap = enable != 0;
  return OK;
}

int candmc_off_alloc(int mat, int64_t size, const double* host_init) {
  CANDMC_TRY(off_init());
  CANDMC_CHECK(mat >= 0 && mat < NMAT, "alloc_offload: unknown offloaded matrix %d", mat);
  CANDMC_CHECK(size >= 0, "alloc_offload: negative size");
  CANDMC_TRY(free_mat(mat));
  OffMat& m = g_off.mat[mat];
  cudaError_t e = cudaMalloc(&m.dev, sizeof(double) * std::max<int64_t>(size, 1));
  if (e != cudaSuccess) {
    m.dev = nullptr;
    set_last_error("alloc_offload: cudaMalloc(%lld doubles) failed: %s", (long long)size, cudaGetErrorString(e));
    return ERR_NOMEM;
  }
  m.size = size;
  if (host_init != nullptr && size > 0) {
    CANDMC_CUDA(cudaMemcpyAsync(m.dev, host_init, sizeof(double) * size, cudaMemcpyHostToDevice, xfer_stream()));
    ++g_off.n_upload;
  }
  return OK;
}

int candmc_off_free(int mat) { return free_mat(mat); }

int candmc_off_alloc_transfer(int64_t size) {
  CANDMC_TRY(off_init());
  CANDMC_CHECK(size >= 0, "alloc_transfer: negative size");
  g_off.transfer_hint = size;
  return rows_scratch(size);
}

int candmc_off_free_transfer(void) {
  if (!g_off.ready) return OK;
  CANDMC_CUDA(cudaStreamSynchronize(g_off.gemm_stream));
  CANDMC_CUDA(cudaStreamSynchronize(g_off.xfer_stream));
  if (g_off.d_rows) CANDMC_CUDA(cudaFree(g_off.d_rows));
  if (g_off.h_rows) CANDMC_CUDA(cudaFreeHost(g_off.h_rows));
  if (g_off.d_offs) CANDMC_CUDA(cudaFree(g_off.d_offs));
  if (g_off.h_offs) CANDMC_CUDA(cudaFreeHost(g_off.h_offs));
  g_off.d_rows = g_off.h_rows = nullptr;
  g_off.d_offs = g_off.h_offs = nullptr;
  g_off.rows_cap = g_off.offs_cap = 0;
  g_off.transfer_hint = 0;
  return OK;
}

int candmc_off_device_ptr(int mat, double** out, int64_t* size) {
  CANDMC_TRY(check_mat(mat, "get_mat_handle"));
  CANDMC_CHECK(out != nullptr, "get_mat_handle: null output");
  CANDMC_TRY(flush_mirror(mat));
  *out = g_off.mat[mat].dev;
  if (size) *size = g_off.mat[mat].size;
  return OK;
}

int candmc_off_size(int mat, int64_t* size) {
  CANDMC_TRY(check_mat(mat, "candmc_off_size"));
  CANDMC_CHECK(size != nullptr, "candmc_off_size: null output");
  *size = g_off.mat[mat].size;
  return OK;
}

int candmc_off_host_mirror(int mat, double** out) {
  CANDMC_TRY(check_mat(mat, "get_mat_handle"));
  CANDMC_CHECK(out != nullptr, "get_mat_handle: null output");
  OffMat& m = g_off.mat[mat];
  if (m.mirror == nullptr) {
    cudaError_t e = cudaMallocHost(&m.mirror, sizeof(double) * std::max<int64_t>(m.size, 1));
    if (e != cudaSuccess) {
      m.mirror = nullptr;
      set_last_error("get_mat_handle: cudaMallocHost(%lld doubles) failed: %s", (long long)m.size,
                     cudaGetErrorString(e));
      return ERR_NOMEM;
    }
  }
  // hand the caller the CURRENT contents (all queued work on the matrix finishes first); if the mirror already holds
  // writes that have not been pushed yet, it IS the current contents
  if (!m.mirror_dirty) {
    CANDMC_CUDA(cudaStreamSynchronize(g_off.gemm_stream));
    CANDMC_CUDA(cudaMemcpyAsync(m.mirror, m.dev, sizeof(double) * m.size, cudaMemcpyDeviceToHost, xfer_stream()));
    CANDMC_CUDA(cudaStreamSynchronize(xfer_stream()));
  }
  // ... and assume it writes them: the mirror goes back to HBM before the next operation on this matrix
  m.mirror_dirty = true;
  *out = m.mirror;
  return OK;
}

int candmc_off_gemm(char tA, char tB, int64_t m, int64_t n, int64_t k, double alpha, int64_t offset_A, int mat_A,
                    int64_t lda_A, int64_t offset_B, int mat_B, int64_t lda_B, double beta, int64_t offset_C,
                    int mat_C, int64_t lda_C) {
  CANDMC_TRY(off_init());
  const bool ta = (tA == 'T' || tA == 't'), tb = (tB == 'T' || tB == 't');
  CANDMC_CHECK(ta || tA == 'N' || tA == 'n', "offload_gemm_A: bad transpose flag '%c' for A", tA);
  CANDMC_CHECK(tb || tB == 'N' || tB == 'n', "offload_gemm_A: bad transpose flag '%c' for B", tB);
  CANDMC_CHECK(m >= 0 && n >= 0 && k >= 0, "offload_gemm_A: negative dimension");
  Block a{mat_A, offset_A, lda_A, ta ? k : m, ta ? m : k};
  Block b{mat_B, offset_B, lda_B, tb ? n : k, tb ? k : n};
  Block c{mat_C, offset_C, lda_C, m, n};
  CANDMC_TRY(check_block(a, "offload_gemm_A(A)"));
  CANDMC_TRY(check_block(b, "offload_gemm_A(B)"));
  CANDMC_TRY(check_block(c, "offload_gemm_A(C)"));
  CANDMC_CHECK(!blocks_overlap(c, a) && !blocks_overlap(c, b), "offload_gemm_A: C overlaps an input block");
  CANDMC_TRY(flush_mirror(mat_A));
  CANDMC_TRY(flush_mirror(mat_B));
  CANDMC_TRY(flush_mirror(mat_C));
  if (m == 0 || n == 0) return OK;
  if (g_off.overlap) {
    // every transfer issued so far happens before this GEMM
    CANDMC_CUDA(cudaEventRecord(g_off.xfer_tail, g_off.xfer_stream));
    CANDMC_CUDA(cudaStreamWaitEvent(g_off.gemm_stream, g_off.xfer_tail, 0));
  }
  CANDMC_TRY(gemm_f64(ta ? 'T' : 'N', tb ? 'T' : 'N', m, n, k, alpha, g_off.mat[mat_A].dev + offset_A, lda_A,
                      g_off.mat[mat_B].dev + offset_B, lda_B, beta, g_off.mat[mat_C].dev + offset_C, lda_C,
                      g_off.gemm_stream));
  ++g_off.n_gemm;
  if (g_off.overlap) {
    InflightGemm g{a, b, c, nullptr};
    CANDMC_TRY(event_get(&g.done));
    CANDMC_CUDA(cudaEventRecord(g.done, g_off.gemm_stream));
    g_off.inflight.push_back(g);
    if (g_off.inflight.size() > 64) CANDMC_TRY(retire_finished());
  }
  return OK;
}

int candmc_off_wait_gemm(void) {
  if (!g_off.ready) return OK;
  CANDMC_CUDA(cudaStreamSynchronize(g_off.gemm_stream));
  return retire_finished();
}

int candmc_off_upload(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, int64_t offset_B,
                      int mat_B) {
  CANDMC_TRY(off_init());
  Block dst{mat_B, offset_B, lda_B, nrow, ncol};
  CANDMC_TRY(check_block(dst, "upload_lda_cpy"));
  if (nrow == 0 || ncol == 0) return OK;
  CANDMC_CHECK(A != nullptr, "upload_lda_cpy: null source");
  CANDMC_CHECK(lda_A >= nrow || ncol == 1, "upload_lda_cpy: source leading dimension smaller than nrow");
  CANDMC_TRY(flush_mirror(mat_B));
  CANDMC_TRY(order_transfer_after_gemms(dst, true));
  double* d = g_off.mat[mat_B].dev + offset_B;
  cudaStream_t s = xfer_stream();
  if (is_device_ptr(A)) {
    CANDMC_TRY(lda_copy_f64(nrow, ncol, lda_A, lda_B, A, d, s));
  } else if ((lda_A == nrow && lda_B == nrow) || ncol == 1) {
    CANDMC_CUDA(cudaMemcpyAsync(d, A, sizeof(double) * nrow * ncol, cudaMemcpyHostToDevice, s));
  } else {
    CANDMC_CUDA(cudaMemcpy2DAsync(d, lda_B * 8, A, lda_A * 8, nrow * 8, ncol, cudaMemcpyHostToDevice, s));
  }
  ++g_off.n_upload;
  return OK;
}

int candmc_off_download(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, int64_t offset_A, double* B,
                        int mat_A) {
  CANDMC_TRY(off_init());
  Block src{mat_A, offset_A, lda_A, nrow, ncol};
  CANDMC_TRY(check_block(src, "download_lda_cpy"));
  if (nrow == 0 || ncol == 0) return OK;
  CANDMC_CHECK(B != nullptr, "download_lda_cpy: null destination");
  CANDMC_CHECK(lda_B >= nrow || ncol == 1, "download_lda_cpy: destination leading dimension smaller than nrow");
  CANDMC_TRY(flush_mirror(mat_A));
  CANDMC_TRY(order_transfer_after_gemms(src, false));
  const double* d = g_off.mat[mat_A].dev + offset_A;
  cudaStream_t s = xfer_stream();
  if (is_device_ptr(B)) {
    CANDMC_TRY(lda_copy_f64(nrow, ncol, lda_A, lda_B, d, B, s));
  } else if ((lda_A == nrow && lda_B == nrow) || ncol == 1) {
    CANDMC_CUDA(cudaMemcpyAsync(B, d, sizeof(double) * nrow * ncol, cudaMemcpyDeviceToHost, s));
  } else {
    CANDMC_CUDA(cudaMemcpy2DAsync(B, lda_B * 8, d, lda_A * 8, nrow * 8, ncol, cudaMemcpyDeviceToHost, s));
  }
  CANDMC_CUDA(cudaStreamSynchronize(s));
  ++g_off.n_download;
  return OK;
}

int candmc_off_sparse_rw(int64_t nrow, int64_t ncol, int64_t lda_B, double* A, int64_t lda_A, const int* offsets,
                         int mat_B, char rw) {
  if (ncol == 0 || nrow == 0) return OK;  // lu_offload.cxx:432
  CANDMC_TRY(off_init());
  CANDMC_CHECK(rw == 'r' || rw == 'w' || rw == 's', "offload_sparse_rw: mode '%c' is not r, w or s", rw);
  CANDMC_TRY(check_mat(mat_B, "offload_sparse_rw"));
  CANDMC_CHECK(nrow > 0 && ncol > 0, "offload_sparse_rw: negative extent");
  CANDMC_CHECK(A != nullptr && offsets != nullptr, "offload_sparse_rw: null buffer");
  CANDMC_CHECK(lda_A >= ncol || nrow == 1, "offload_sparse_rw: host row stride smaller than ncol");
  CANDMC_CHECK(lda_B >= 1 || ncol == 1, "offload_sparse_rw: bad leading dimension");
  CANDMC_CHECK(!is_device_ptr(A), "offload_sparse_rw: the row buffer must be host memory");
  OffMat& m = g_off.mat[mat_B];
  CANDMC_TRY(flush_mirror(mat_B));
  CANDMC_TRY(offs_scratch(nrow));
  CANDMC_TRY(rows_scratch(nrow * ncol));
  int64_t lo = INT64_MAX, hi = 0;
  for (int64_t i = 0; i < nrow; ++i) {
    const int64_t o = offsets[i];
    CANDMC_CHECK(o >= 0 && o + (ncol - 1) * lda_B < m.size, "offload_sparse_rw: row %lld (offset %lld) leaves the matrix",
                 (long long)i, (long long)o);
    g_off.h_offs[i] = o;
    lo = std::min(lo, o);
    hi = std::max(hi, o + (ncol - 1) * lda_B + 1);
  }
  // rows that alias each other must be processed one after the other, as the reference's loop does (only matters for
  // writes and swaps; pivoting never produces them, so this is the slow, exact path)
  bool distinct = true;
  if (rw != 'r' && nrow > 1) {
    // rows o1, o2 share an element iff o1 == o2 (mod ld) and they start fewer than ncol columns apart
    const int64_t ld = ncol == 1 ? 1 : lda_B;
    std::vector<std::pair<int64_t, int64_t>> key(nrow);  // (offset mod ld, offset)
    for (int64_t i = 0; i < nrow; ++i) key[i] = std::make_pair(g_off.h_offs[i] % ld, g_off.h_offs[i]);
    std::sort(key.begin(), key.end());
    for (int64_t i = 1; i < nrow && distinct; ++i)
      distinct = key[i].first != key[i - 1].first || (key[i].second - key[i - 1].second) / ld >= ncol;
  }
  // scoreboard: treat the touched rows as one enclosing interval of the matrix
  Block span{mat_B, lo, hi - lo, hi - lo, 1};
  CANDMC_TRY(order_transfer_after_gemms(span, rw != 'r'));
  cudaStream_t s = xfer_stream();
  CANDMC_CUDA(cudaMemcpyAsync(g_off.d_offs, g_off.h_offs, sizeof(int64_t) * nrow, cudaMemcpyHostToDevice, s));
  if (rw != 'r') {
    // pack the host rows (stride lda_A) into the pinned buffer, then one contiguous H2D
    for (int64_t i = 0; i < nrow; ++i) memcpy(g_off.h_rows + i * ncol, A + i * lda_A, sizeof(double) * ncol);
    CANDMC_CUDA(cudaMemcpyAsync(g_off.d_rows, g_off.h_rows, sizeof(double) * nrow * ncol, cudaMemcpyHostToDevice, s));
  }
  const int mode = rw == 'r' ? 0 : (rw == 'w' ? 1 : 2);
  if (distinct) {
    CANDMC_TRY(launch_sparse(mode, m.dev, lda_B, nrow, ncol, 0, s));
  } else {
    for (int64_t i = 0; i < nrow; ++i) CANDMC_TRY(launch_sparse(mode, m.dev, lda_B, 1, ncol, i, s));
  }
  if (rw != 'w') {
    CANDMC_CUDA(cudaMemcpyAsync(g_off.h_rows, g_off.d_rows, sizeof(double) * nrow * ncol, cudaMemcpyDeviceToHost, s));
  }
  // the pinned buffers are reused by the next call, and reads must be visible to the caller: finish here
  CANDMC_CUDA(cudaStreamSynchronize(s));
  if (rw != 'w') {
    for (int64_t i = 0; i < nrow; ++i) memcpy(A + i * lda_A, g_off.h_rows + i * ncol, sizeof(double) * ncol);
  }
  ++g_off.n_sparse;
  return OK;
}

int candmc_off_sync(void) {
  if (!g_off.ready) return OK;
  CANDMC_CUDA(cudaStreamSynchronize(g_off.gemm_stream));
  CANDMC_CUDA(cudaStreamSynchronize(g_off.xfer_stream));
  return retire_finished();
}

int candmc_off_stats(int64_t* out6) {
  CANDMC_CHECK(out6 != nullptr, "candmc_off_stats: null output");
  out6[0] = g_off.n_gemm;
  out6[1] = g_off.n_upload;
  out6[2] = g_off.n_download;
  out6[3] = g_off.n_sparse;
  out6[4] = g_off.n_waits;
  out6[5] = static_cast<int64_t>(g_off.inflight.size());
  return OK;
}

// CPU-testable piece of the scoreboard (tests/test_host_logic.py): do two strided blocks of one matrix overlap?
int candmc_off_blocks_overlap(int64_t off_x, int64_t ld_x, int64_t rows_x, int64_t cols_x, int64_t off_y,
                              int64_t ld_y, int64_t rows_y, int64_t cols_y) {
  Block x{0, off_x, ld_x, rows_x, cols_x}, y{0, off_y, ld_y, rows_y, cols_y};
  return blocks_overlap(x, y) ? 1 : 0;
}

}  // extern "C"
