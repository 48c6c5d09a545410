// candmc_b200 — HBM-bound packing / layout kernels.
//   lda_copy_f64   <- reference `lda_cpy`        (alg/shared/util.h:459-471)  strided sub-matrix copy
//   lda_axpby_f64  <- reference scaled `lda_cpy` (alg/shared/util.h:484-501)  B = b*B + a*A
//   transpose_f64  <- reference `TRANSPOSE`/`naive_transp` (alg/MM/splitdim_cannon/spcannon_internal.h:66-72)
//   drand48_fill_f64: the reference unit test's per-element generator (test/MM/topo_pdgemm_unit.cxx:250-256) on device
//   frob_diff_f64  : ||X-Y||_F^2 and ||Y||_F^2 for the rel-Frobenius parity bar
// These are pure streaming kernels: 16 B per thread per access when alignment allows, grid = k * #SMs,
// algorithmic traffic 16 B/element (copy, transpose) or 24 B/element (axpby).
#include "common.cuh"
#include "runtime.h"

namespace candmc {

namespace {

constexpr int PACK_THREADS = 256;
constexpr int PACK_CTAS_PER_SM = 8;

// One "row of work" = a column of the sub-matrix; vector path moves double2.
template <bool AXPBY>
__global__ void __launch_bounds__(PACK_THREADS)
lda_vec2_kernel(int64_t nrow2 /* nrow/2 */, int64_t ncol, int64_t lda2, int64_t ldb2, const double2* __restrict__ A,
                double2* __restrict__ B, double a, double b) {
  const int64_t total = nrow2 * ncol;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total; e += stride) {
    const int64_t c = e / nrow2, r = e - c * nrow2;
    const double2 va = __ldg(A + c * lda2 + r);
    double2* bp = B + c * ldb2 + r;
    if (AXPBY) {
      const double2 vb = *bp;
      *bp = make_double2(vb.x * b + va.x * a, vb.y * b + va.y * a);
    } else {
      *bp = va;
    }
  }
}

template <bool AXPBY>
__global__ void __launch_bounds__(PACK_THREADS)
lda_scalar_kernel(int64_t nrow, int64_t ncol, int64_t lda, int64_t ldb, const double* __restrict__ A,
                  double* __restrict__ B, double a, double b) {
  const int64_t total = nrow * ncol;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total; e += stride) {
    const int64_t c = e / nrow, r = e - c * nrow;
    const double va = __ldg(A + c * lda + r);
    double* bp = B + c * ldb + r;
    *bp = AXPBY ? (*bp * b + va * a) : va;
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
                 double b, cudaStream_t stream) {
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
  if (vec) {
    lda_vec2_kernel<AXPBY><<<pack_grid(nrow / 2 * ncol), PACK_THREADS, 0, stream>>>(
        nrow / 2, ncol, lda_A / 2, lda_B / 2, reinterpret_cast<const double2*>(A), reinterpret_cast<double2*>(B), a,
        b);
  } else {
    lda_scalar_kernel<AXPBY><<<pack_grid(nrow * ncol), PACK_THREADS, 0, stream>>>(nrow, ncol, lda_A, lda_B, A, B, a,
                                                                                   b);
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
