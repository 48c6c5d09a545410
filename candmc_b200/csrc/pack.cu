// candmc_b200 — HBM-bound packing / layout kernels.
//   lda_copy_f64   <- reference `lda_cpy`        (alg/shared/util.h:459-471)  strided sub-matrix copy
//   lda_axpby_f64  <- reference scaled `lda_cpy` (alg/shared/util.h:484-501)  B = b*B + a*A
//   transpose_f64  <- reference `TRANSPOSE`/`naive_transp` (alg/MM/splitdim_cannon/spcannon_internal.h:66-72)
//   drand48_fill_f64: the reference unit test's per-element generator (test/MM/topo_pdgemm_unit.cxx:250-256) on device
//   frob_diff_f64  : ||X-Y||_F^2 and ||Y||_F^2 for the rel-Frobenius parity bar
// These are pure streaming kernels: 16 B per thread per access when alignment allows, grid = k * #SMs,
// algorithmic traffic 16 B/element (copy, transpose) or 24 B/element (axpby).
#include <algorithm>

#include "common.cuh"
#include "runtime.h"

namespace candmc {

namespace {

constexpr int PACK_THREADS = 256;
constexpr int PACK_CTAS_PER_SM = 8;

// Strided sub-matrix copy / axpby in tiles of PACK_ITEMS elements of T (double2 when alignment allows, else double): a tile is
// 2^tr_log2 consecutive rows x (PACK_ITEMS >> tr_log2) columns, so a thread finds its element with shifts and masks — the only
// 64-bit division is one per tile, not one per element — and issues all PACK_UNROLL of its loads before the first store.
// Tall columns give tiles of 1024 rows x 1 column (16 KiB contiguous per tile), short ones pack several columns per tile.
constexpr int PACK_UNROLL = 4;
constexpr int PACK_ITEMS = PACK_THREADS * PACK_UNROLL;

template <bool AXPBY, class T>
__device__ __forceinline__ T axpby_elem(T va, T vb, double a, double b);
template <>
__device__ __forceinline__ double axpby_elem<true, double>(double va, double vb, double a, double b) { return vb * b + va * a; }
template <>
__device__ __forceinline__ double2 axpby_elem<true, double2>(double2 va, double2 vb, double a, double b) {
  return make_double2(vb.x * b + va.x * a, vb.y * b + va.y * a);
}
template <>
__device__ __forceinline__ double axpby_elem<false, double>(double va, double, double, double) { return va; }
template <>
__device__ __forceinline__ double2 axpby_elem<false, double2>(double2 va, double2, double, double) { return va; }

template <bool AXPBY, class T>
__global__ void __launch_bounds__(PACK_THREADS)
lda_tile_kernel(int64_t nrow /* in units of T */, int64_t ncol, int64_t lda, int64_t ldb, const T* __restrict__ A,
                T* __restrict__ B, double a, double b, int tr_log2, int64_t tiles_r, int64_t ntiles) {
  const int tr_mask = (1 << tr_log2) - 1;
  const int tc = PACK_ITEMS >> tr_log2;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t tile_c = tile / tiles_r, tile_r = tile - tile_c * tiles_r;
    const int64_t row0 = tile_r << tr_log2, col0 = tile_c * tc;
    T va[PACK_UNROLL], vb[PACK_UNROLL];
    bool live[PACK_UNROLL];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
      const int i = threadIdx.x + u * PACK_THREADS;
      const int64_t r = row0 + (i & tr_mask), c = col0 + (i >> tr_log2);
      live[u] = r < nrow && c < ncol;
      if (live[u]) {
        va[u] = __ldg(A + c * lda + r);
        if (AXPBY) vb[u] = B[c * ldb + r];
      }
    }
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
      const int i = threadIdx.x + u * PACK_THREADS;
      const int64_t r = row0 + (i & tr_mask), c = col0 + (i >> tr_log2);
      if (live[u]) B[c * ldb + r] = axpby_elem<AXPBY, T>(va[u], vb[u], a, b);
    }
  }
}

int pack_grid(int64_t work_items) {
  int64_t g = (work_items + PACK_THREADS - 1) / PACK_THREADS;
  const int64_t cap = static_cast<int64_t>(runtime().num_sms) * PACK_CTAS_PER_SM;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

template <bool AXPBY>
int lda_dispatch(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B, double a,
                 double b, cudaStream_t stream, int max_ctas = 0) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(nrow >= 0 && ncol >= 0, "lda_cpy: negative extent");
  CANDMC_CHECK(lda_A >= nrow && lda_B >= nrow, "lda_cpy: leading dimension smaller than nrow");
  if (nrow == 0 || ncol == 0) return OK;
  // contiguous case == one long column (the reference's single-memcpy branch, util.h:463-464)
  if (lda_A == nrow && lda_B == nrow) {
    nrow = nrow * ncol;
    ncol = 1;
    lda_A = lda_B = nrow;
  }
  const bool vec = (nrow % 2 == 0) && (lda_A % 2 == 0) && (lda_B % 2 == 0) &&
                   (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 16 == 0);
  const int64_t nr = vec ? nrow / 2 : nrow;   // rows in units of the element type the kernel moves
  int tr_log2 = 0;
  while ((1 << tr_log2) < PACK_ITEMS && (static_cast<int64_t>(1) << tr_log2) < nr) ++tr_log2;
  const int64_t tiles_r = (nr + (1 << tr_log2) - 1) >> tr_log2;
  const int tc = PACK_ITEMS >> tr_log2;
  const int64_t ntiles = tiles_r * ((ncol + tc - 1) / tc);
  int64_t grid = ntiles;
  const int64_t cap = max_ctas > 0 ? max_ctas : static_cast<int64_t>(runtime().num_sms) * PACK_CTAS_PER_SM;
  if (grid > cap) grid = cap;
  if (vec) {
    lda_tile_kernel<AXPBY, double2><<<static_cast<int>(grid), PACK_THREADS, 0, stream>>>(
        nr, ncol, lda_A / 2, lda_B / 2, reinterpret_cast<const double2*>(A), reinterpret_cast<double2*>(B), a, b, tr_log2,
        tiles_r, ntiles);
  } else {
    lda_tile_kernel<AXPBY, double><<<static_cast<int>(grid), PACK_THREADS, 0, stream>>>(nr, ncol, lda_A, lda_B, A, B, a, b,
                                                                                         tr_log2, tiles_r, ntiles);
  }
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

// B(cols x rows, ldb) = A(rows x cols, lda)^T, 64x64 tiles through padded shared memory; both the global
// read and the global write are row-contiguous (coalesced 512 B per warp-row pair).
constexpr int TT = 64;
__global__ void __launch_bounds__(256, 6)
transpose_kernel(int64_t rows, int64_t cols, const double* __restrict__ A, int64_t lda, double* __restrict__ B,
                 int64_t ldb, int64_t tiles_r, int64_t tiles_c) {
  __shared__ double tile[TT][TT + 1];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  const int64_t ntiles = tiles_r * tiles_c;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t r0 = (t % tiles_r) * TT, c0 = (t / tiles_r) * TT;
    double v[16];
    // all 16 loads of a thread are issued before the first use (memory-level parallelism: 128 B in flight per thread)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t r = r0 + tx, c = c0 + ty + 4 * i;
      v[i] = (r < rows && c < cols) ? __ldg(A + r + c * lda) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[ty + 4 * i][tx] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tile[tx][ty + 4 * i];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t c = c0 + tx, r = r0 + ty + 4 * i;
      if (r < rows && c < cols) B[c + r * ldb] = v[i];
    }
    __syncthreads();
  }
}

// The same through the TMA unit (the default whenever both matrices are 16-byte aligned with even leading dimensions): a
// persistent CTA per SM streams 64 x 64 tiles — four 16-row boxes per tile land by `cp.async.bulk.tensor` under the 128-byte
// swizzle (TP_NIN tiles in flight per SM, completion on mbarriers), the threads turn a tile around inside shared memory
// (one conflict-free LDS.128 per pair of rows: the swizzle spreads eight consecutive columns over the eight 16-byte bank
// groups; the stores along the output's contiguous dimension are 256 B per warp), and the turned tile leaves by one TMA
// store (TP_NOUT tiles in flight).  No thread computes a global address; ragged edges are clipped / zero-filled by the TMA
// unit.  Algorithmic traffic 16 B per element.
constexpr int TP = 64;
constexpr int TP_BYTES = TP * TP * 8;
constexpr int TP_NIN = 3, TP_NOUT = 3;
constexpr int TP_SMEM = (TP_NIN + TP_NOUT) * TP_BYTES + 1024 /*align slack*/ + 64 /*barriers*/;

__global__ void __launch_bounds__(256, 1)
transpose_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, int tiles_r,
                     int tiles_c) {
  extern __shared__ uint8_t tp_smem_raw[];
  uint8_t* smem = tp_smem_raw + ((1024u - (smem_u32(tp_smem_raw) & 1023u)) & 1023u);
  uint8_t* in = smem;                            // TP_NIN tiles: box q (rows 16q .. 16q+15) at q * 8 KiB, one 128 B line per column
  uint8_t* out = smem + TP_NIN * TP_BYTES;       // TP_NOUT tiles: out[r][c], c contiguous
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (TP_NIN + TP_NOUT) * TP_BYTES);
  const int64_t ntiles = static_cast<int64_t>(tiles_r) * tiles_c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto issue_load = [&](int64_t t, int s) {
    const int r0 = static_cast<int>(t % tiles_r) * TP, c0 = static_cast<int>(t / tiles_r) * TP;
    mbar_expect_tx(&full[s], TP_BYTES);
#pragma unroll
    for (int q = 0; q < 4; ++q) tma_load_2d(in + s * TP_BYTES + q * 8192, &tmIn, &full[s], r0 + 16 * q, c0);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < TP_NIN; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < TP_NIN; ++s) {
      const int64_t t = blockIdx.x + static_cast<int64_t>(s) * gridDim.x;
      if (t < ntiles) issue_load(t, s);
    }
  }
  const int c = (warp & 1) * 32 + lane;   // my column of the input tile = my position along the output's contiguous dimension
  int i = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
    const int s = i % TP_NIN, o = i % TP_NOUT;
    mbar_wait(&full[s], (i / TP_NIN) & 1);
    const uint32_t ibase = smem_u32(in + s * TP_BYTES) + c * 128;
    double* ob = reinterpret_cast<double*>(out + o * TP_BYTES) + c;
    // out[o] was last read by the store of tile i - TP_NOUT: complete, thread 0 waited for it before the barrier of tile i - 1
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = (warp >> 1) * 8 + u, q = p >> 3, j = p & 7;   // rows 16q + 2j, 16q + 2j + 1
      const double2 v = lds_f64x2(ibase + q * 8192 + ((j ^ (c & 7)) << 4));
      ob[(16 * q + 2 * j) * TP] = v.x;
      ob[(16 * q + 2 * j + 1) * TP] = v.y;
    }
    fence_proxy_async();   // my writes of out[o] become visible to the TMA store
    if (threadIdx.x == 0) tma_store_wait_read<TP_NOUT - 2>();   // the stores up to tile i - 2 have read their buffers
    __syncthreads();
    if (threadIdx.x == 0) {
      const int r0 = static_cast<int>(t % tiles_r) * TP, c0 = static_cast<int>(t / tiles_r) * TP;
      tma_store_2d(&tmOut, out + o * TP_BYTES, c0, r0);
      tma_store_commit();
      const int64_t tn = t + static_cast<int64_t>(TP_NIN) * gridDim.x;   // in[s] has been read by every thread: refill it
      if (tn < ntiles) issue_load(tn, s);
    }
  }
  if (threadIdx.x == 0) tma_store_wait_all<0>();
}

__global__ void fill_kernel(double* __restrict__ X, int64_t count, double v) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < count; e += stride) X[e] = v;
}

// glibc srand48/drand48: X0 = (seed mod 2^32) << 16 | 0x330E ; X <- (0x5DEECE66D * X + 0xB) mod 2^48 ; value X / 2^48
__device__ __forceinline__ double drand48_nth(uint64_t seed, int nth /*0 = first draw*/) {
  const uint64_t mask = (1ULL << 48) - 1;
  uint64_t x = ((seed & 0xffffffffULL) << 16) | 0x330EULL;
  for (int i = 0; i <= nth; ++i) x = (0x5DEECE66DULL * x + 0xBULL) & mask;
  return static_cast<double>(x) * (1.0 / 281474976710656.0);
}

__global__ void drand48_fill_kernel(double* __restrict__ X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0,
                                    int64_t col0, int64_t n_global, int which) {
  const int64_t total = nrow * ncol;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total; e += stride) {
    const int64_t c = e / nrow, r = e - c * nrow;
    const uint64_t seed = static_cast<uint64_t>((col0 + c) * n_global + (row0 + r));
    X[r + c * ld] = drand48_nth(seed, which);
  }
}

__global__ void __launch_bounds__(256)
frob_diff_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y, int64_t ldy, int64_t nrow,
                 int64_t ncol, double* __restrict__ out2) {
  double d2 = 0.0, y2 = 0.0;
  const int64_t total = nrow * ncol;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total; e += stride) {
    const int64_t c = e / nrow, r = e - c * nrow;
    const double x = X[r + c * ldx], y = Y[r + c * ldy];
    d2 += (x - y) * (x - y);
    y2 += y * y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    y2 += __shfl_xor_sync(0xffffffffu, y2, o);
  }
  __shared__ double s[2][8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    s[0][w] = d2;
    s[1][w] = y2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) {
      a += s[0][i];
      b += s[1][i];
    }
    atomicAdd(out2, a);
    atomicAdd(out2 + 1, b);
  }
}

}  // namespace

int lda_copy_f64(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                 cudaStream_t stream) {
  return lda_dispatch<false>(nrow, ncol, lda_A, lda_B, A, B, 1.0, 0.0, stream);
}

int lda_copy_f64_capped(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                        cudaStream_t stream, int max_ctas) {
  return lda_dispatch<false>(nrow, ncol, lda_A, lda_B, A, B, 1.0, 0.0, stream, max_ctas);
}

int lda_axpby_f64(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B, double a,
                  double b, cudaStream_t stream) {
  return lda_dispatch<true>(nrow, ncol, lda_A, lda_B, A, B, a, b, stream);
}

int transpose_f64(int64_t rows, int64_t cols, const double* A, int64_t lda, double* B, int64_t ldb,
                  cudaStream_t stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(rows >= 0 && cols >= 0, "transpose: negative extent");
  CANDMC_CHECK(lda >= rows && ldb >= cols, "transpose: leading dimension too small");
  if (rows == 0 || cols == 0) return OK;
  const bool tma_ok = runtime().transpose_tma && reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0 &&
                      lda % 2 == 0 && ldb % 2 == 0 && rows < (1LL << 31) - TP && cols < (1LL << 31) - TP;
  if (tma_ok) {
    CUtensorMap tmIn, tmOut;
    CANDMC_TRY(encode_tmap_f64(&tmIn, A, rows, cols, lda, 16, TP));
    CANDMC_TRY(encode_tmap_f64_linear(&tmOut, B, cols, rows, ldb, TP, TP));
    static bool configured = false;
    if (!configured) {
      CANDMC_CUDA(cudaFuncSetAttribute(transpose_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
      configured = true;
    }
    const int64_t ttr = (rows + TP - 1) / TP, ttc = (cols + TP - 1) / TP;
    const int64_t g = std::min<int64_t>(ttr * ttc, runtime().num_sms);
    transpose_tma_kernel<<<static_cast<int>(g), 256, TP_SMEM, stream>>>(tmIn, tmOut, static_cast<int>(ttr), static_cast<int>(ttc));
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
    return OK;
  }
  const int64_t tr = (rows + TT - 1) / TT, tc = (cols + TT - 1) / TT;
  int64_t grid = tr * tc;
  const int64_t cap = static_cast<int64_t>(runtime().num_sms) * 6;
  if (grid > cap) grid = cap;
  transpose_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(rows, cols, A, lda, B, ldb, tr, tc);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

int fill_f64(double* X, int64_t count, double value, cudaStream_t stream) {
  CANDMC_TRY(runtime_require());
  if (count <= 0) return OK;
  fill_kernel<<<pack_grid(count), PACK_THREADS, 0, stream>>>(X, count, value);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

int drand48_fill_f64(double* X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0, int64_t col0, int64_t n_global,
                     int which, cudaStream_t stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(which == 0 || which == 1, "drand48_fill: which must be 0 (A, first draw) or 1 (B, second draw)");
  if (nrow <= 0 || ncol <= 0) return OK;
  drand48_fill_kernel<<<pack_grid(nrow * ncol), PACK_THREADS, 0, stream>>>(X, nrow, ncol, ld, row0, col0, n_global,
                                                                            which);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

int frob_diff_f64(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t nrow, int64_t ncol,
                  double* out2, cudaStream_t stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(double), stream));
  if (nrow <= 0 || ncol <= 0) return OK;
  frob_diff_kernel<<<pack_grid(nrow * ncol), 256, 0, stream>>>(X, ldx, Y, ldy, nrow, ncol, out2);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

}  // namespace candmc
