// candmc_b200 — FP32 GEMM for sm_100a on the 5th-generation tensor cores:  C = alpha * op(A) * op(B) + beta * C
// (column-major, Fortran sgemm semantics).  The optional single-precision companion of gemm_f64.cu (BASELINE north star:
// "FP64 (and optional FP32) GEMM kernel using DMMA/tcgen05 tiles"); the reference itself has no single-precision path
// (cdgemm, alg/shared/lapack.h:10-16, is double only), so the argument list mirrors cdgemm's with float data.
//
// Design:
//   * tcgen05.mma kind::tf32, M = 128, N = 128, K = 8 per instruction, issued by ONE thread; the 128 x 128 FP32 accumulator
//     lives in TMEM (128 lanes x 128 columns), double-buffered (256 columns) so the epilogue of tile t drains while the MMAs
//     of tile t+1 run.
//   * FP32 accuracy from TF32 multipliers by operand splitting ("3xTF32"): x = hi + lo with hi = x truncated to TF32's 10
//     mantissa bits (exactly what the tensor core reads of a 32-bit word) and lo = x - hi (exact in FP32), then
//     a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b, all three accumulated in the same FP32 TMEM accumulator.  Dropped: lo_a*lo_b
//     (<= 2^-20 |a b|) and the TF32 truncation of lo (<= 2^-20 |x|): per-product relative error <= 3 * 2^-20 before
//     accumulation, against 2^-11 for a single TF32 product.  candmc_set_f32_mode(1) switches to one TF32 product (fast mode).
//   * Operands are K-major in shared memory: a TMA box of 32 floats (one 128-byte line) x 128 rows per operand and k-block,
//     128-byte swizzle; the UMMA shared-memory descriptor addresses that layout directly (8-row groups 1024 B apart), and the
//     four K = 8 steps of a k-block advance the descriptor's start address by 32 B inside the swizzle span.
//     In column-major terms op(A) is K-major when transa = 'T' and op(B) when transb = 'N'; the other two cases are
//     M/N-major, which for 32-bit operands needs the 32-byte-base swizzle atom — not built: the host transposes such an operand
//     into a K-major scratch copy first (pack_kmajor_f32_kernel, HBM-bound, 8 bytes per element against 2*N or 2*M flops).
//   * The split is done in shared memory, layout-agnostically: four warps turn each landed stage {A, B} into {hi_A, hi_B} in
//     place and {lo_A, lo_B} in a second buffer at the same offsets (float4 per thread, conflict-free), then fence the writes
//     towards the async proxy and arrive on the stage's `split` barrier, which is what the MMA thread waits for.
//   * Warp roles (384 threads): warp 0 TMA producer (one lane), warp 1 TMEM owner + MMA issuer (one lane), warps 4-7 operand
//     split, warps 8-11 epilogue (tcgen05.ld 32 lanes x 32 columns per warp, alpha/beta, coalesced column-major stores).
//     Pipelines: full -> split -> (MMA) -> empty per smem stage; tmem_full / tmem_empty per accumulator buffer.
//   * Persistent CTAs, static round-robin over tiles rasterised in groups of 8 tile rows (L2 reuse of the B panel).
//   * Any m, n, k: TMA zero-fills out-of-range rows and k (zeros split into zeros), stores are predicated.
// Algorithmic work per launch: 2*M*N*K flop (the three TF32 products per FP32 product are overhead, not credit).
#include "common.cuh"
#include "runtime.h"
#include "staging.h"

namespace candmc {

namespace {

constexpr int FM = 128;                 // CTA tile rows = UMMA M
constexpr int FN = 128;                 // CTA tile cols = UMMA N
constexpr int FK = 32;                  // k per stage: one 128-byte swizzle span of floats
constexpr int F_OPER_BYTES = FM * FK * 4;          // 16 KiB per operand per stage
constexpr int F_PAIR_BYTES = 2 * F_OPER_BYTES;     // {A, B}: 32 KiB
constexpr int F_THREADS = 384;
constexpr int F_SPLIT_WARP0 = 4, F_EPI_WARP0 = 8;  // epilogue warps 8..11: warp % 4 selects the TMEM lane quarter
constexpr int F_TMEM_COLS = 2 * FN;                // two accumulator buffers
constexpr int F_RASTER = 8;

template <bool SPLIT3>
struct FCfg {
  static constexpr int NSTAGE = SPLIT3 ? 3 : 6;
  static constexpr int STAGE_BYTES = SPLIT3 ? 2 * F_PAIR_BYTES : F_PAIR_BYTES;   // + {lo_A, lo_B}
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers + TMEM slot*/;
};

__device__ __forceinline__ void f32_tile_coord(int t, int tilesM, int tilesN, int* tm, int* tn) {
  const int group = F_RASTER * tilesN;
  const int g = t / group, r = t - g * group;
  const int rows = min(F_RASTER, tilesM - g * F_RASTER);
  *tm = g * F_RASTER + r % rows;
  *tn = r / rows;
}

template <bool SPLIT3>
__global__ void __launch_bounds__(F_THREADS, 1)
gemm_f32_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                     int64_t ldc, int M, int N, int K, float alpha, float beta, int tilesM, int tilesN) {
  using Cfg = FCfg<SPLIT3>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024 B: the swizzle atom
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE_BYTES);
  uint64_t* split_bar = full_bar + NSTAGE;
  uint64_t* empty_bar = split_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KT = (K + FK - 1) / FK;
  const int ntiles = tilesM * tilesN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, F_TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int tm, tn;
        f32_tile_coord(t, tilesM, tilesN, &tm, &tn);
        for (int kt = 0; kt < KT; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], F_PAIR_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kt * FK, tm * FM);
          tma_load_2d(sa + F_OPER_BYTES, &tmB, &full_bar[stage], kt * FK, tn * FN);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(FM, FN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[buf], bphase ^ 1);   // the epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + buf * FN;
        for (int kt = 0; kt < KT; ++kt) {
          mbar_wait(SPLIT3 ? &split_bar[stage] : &full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + F_OPER_BYTES;
#pragma unroll
          for (int ks = 0; ks < FK / 8; ++ks) {
            const uint64_t dah = umma_desc_kmajor_sw128(sa + ks * 32), dbh = umma_desc_kmajor_sw128(sb + ks * 32);
            const uint32_t first = (kt | ks) ? 1u : 0u;
            if (SPLIT3) {
              const uint64_t dal = umma_desc_kmajor_sw128(sa + F_PAIR_BYTES + ks * 32);
              const uint64_t dbl = umma_desc_kmajor_sw128(sb + F_PAIR_BYTES + ks * 32);
              umma_tf32(acc, dal, dbh, idesc, first);   // small terms first
              umma_tf32(acc, dah, dbl, idesc, 1u);
              umma_tf32(acc, dah, dbh, idesc, 1u);
            } else {
              umma_tf32(acc, dah, dbh, idesc, first);
            }
          }
          umma_commit(&empty_bar[stage]);                    // the stage is free once these MMAs have read it
          if (kt + 1 == KT) umma_commit(&tfull_bar[buf]);    // ... and after the last one the accumulator is complete
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (SPLIT3 && warp >= F_SPLIT_WARP0 && warp < F_SPLIT_WARP0 + 4) {
    // ===================== operand split: {A, B} -> {hi_A, hi_B} in place, {lo_A, lo_B} behind =====================
    const int tid = threadIdx.x - F_SPLIT_WARP0 * 32;   // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int kt = 0; kt < KT; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll 4
        for (int i = 0; i < F_PAIR_BYTES / 16 / 128; ++i) {
          const uint32_t a = base + (i * 128 + tid) * 16;
          const float4 x = lds_f32x4(a);
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
          sts_f32x4(a, h);
          sts_f32x4(a + F_PAIR_BYTES, l);
        }
        fence_proxy_async();            // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(&split_bar[stage]);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= F_EPI_WARP0) {
    // ===================== epilogue: TMEM -> registers -> C =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      int tm, tn;
      f32_tile_coord(t, tilesM, tilesN, &tm, &tn);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[buf], bphase);
      tcgen05_fence_after();
      const int row = tm * FM + q * 32 + lane;
      float* crow = C + row;
#pragma unroll 1
      for (int c0 = 0; c0 < FN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * FN + c0, v);
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = tn * FN + c0 + j;
            if (col < N) {
              float* p = crow + static_cast<int64_t>(col) * ldc;   // lanes of a warp: 32 consecutive rows of one column
              const float acc = alpha * __uint_as_float(v[j]);
              *p = (beta == 0.0f) ? acc : acc + beta * *p;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }

  // ---- teardown: every role has finished its tiles; the allocating warp returns the TMEM columns ----
  __syncwarp();   // the single-lane roles' idle lanes wait here for their working lane before the block-wide barrier
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, F_TMEM_COLS);
}

// dst(k x r, ld_dst) K-major copy of an operand: TRANS = false: dst = src (k x r, ld_src) re-pitched; TRANS = true: dst = src^T
// with src r x k.  64 x 64 tiles through padded shared memory, 8 bytes per element of HBM traffic.
template <bool TRANS>
__global__ void __launch_bounds__(256)
pack_kmajor_f32_kernel(int k, int r, const float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst) {
  __shared__ float tile[64][65];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  const int tilesK = (k + 63) / 64, tilesR = (r + 63) / 64;
  for (int t = blockIdx.x; t < tilesK * tilesR; t += gridDim.x) {
    const int k0 = (t % tilesK) * 64, r0 = (t / tilesK) * 64;
    if (!TRANS) {
      for (int i = ty; i < 64; i += 4) {
        const int kk = k0 + tx, rr = r0 + i;
        if (kk < k && rr < r) dst[kk + static_cast<int64_t>(rr) * ld_dst] = src[kk + static_cast<int64_t>(rr) * ld_src];
      }
    } else {
      for (int i = ty; i < 64; i += 4) {   // src is r x k: read with r contiguous
        const int rr = r0 + tx, kk = k0 + i;
        tile[i][tx] = (rr < r && kk < k) ? src[rr + static_cast<int64_t>(kk) * ld_src] : 0.0f;
      }
      __syncthreads();
      for (int i = ty; i < 64; i += 4) {   // write with k contiguous
        const int kk = k0 + tx, rr = r0 + i;
        if (kk < k && rr < r) dst[kk + static_cast<int64_t>(rr) * ld_dst] = tile[tx][i];
      }
      __syncthreads();
    }
  }
}

__global__ void scale_c_f32_kernel(int M, int N, float beta, float* __restrict__ C, int64_t ldc) {
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float* p = C + (e % M) + (e / M) * ldc;
    *p = (beta == 0.0f) ? 0.0f : beta * *p;
  }
}

bool f_trans(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }
bool f_notrans(char t) { return t == 'N' || t == 'n'; }

int g_f32_mode = 3;   // TF32 products per FP32 product: 3 (FP32-level accuracy, default) or 1 (TF32 accuracy)

template <bool SPLIT3>
int launch_f32(const CUtensorMap& tmA, const CUtensorMap& tmB, float* C, int64_t ldc, int M, int N, int K, float alpha,
               float beta, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_f32_umma_kernel<SPLIT3>;
  if (!configured) {
    CANDMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FCfg<SPLIT3>::SMEM_BYTES));
    configured = true;
  }
  const int tilesM = (M + FM - 1) / FM, tilesN = (N + FN - 1) / FN;
  const int64_t tiles = static_cast<int64_t>(tilesM) * tilesN;
  CANDMC_CHECK(tiles < (1LL << 31), "sgemm: too many tiles");
  const int sms = runtime().num_sms;
  const int avail = sms - runtime().gemm_reserve_sms > 0 ? sms - runtime().gemm_reserve_sms : 1;
  const int grid = static_cast<int>(tiles < avail ? tiles : avail);
  kern<<<grid, F_THREADS, FCfg<SPLIT3>::SMEM_BYTES, stream>>>(tmA, tmB, C, ldc, M, N, K, alpha, beta, tilesM, tilesN);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

}  // namespace

int gemm_f32(char transa, char transb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
             const float* B, int64_t ldb, float beta, float* C, int64_t ldc, cudaStream_t stream) {
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(f_trans(transa) || f_notrans(transa), "sgemm: bad transa '%c'", transa);
  CANDMC_CHECK(f_trans(transb) || f_notrans(transb), "sgemm: bad transb '%c'", transb);
  const bool tA = f_trans(transa), tB = f_trans(transb);
  CANDMC_CHECK(m >= 0 && n >= 0 && k >= 0, "sgemm: negative dimension m=%lld n=%lld k=%lld", (long long)m, (long long)n,
               (long long)k);
  CANDMC_CHECK(m < (1LL << 31) && n < (1LL << 31) && k < (1LL << 31), "sgemm: dimension exceeds 2^31-1");
  const int64_t rowsA = tA ? k : m, rowsB = tB ? n : k;
  CANDMC_CHECK(lda >= (rowsA > 1 ? rowsA : 1), "sgemm: lda=%lld < %lld", (long long)lda, (long long)rowsA);
  CANDMC_CHECK(ldb >= (rowsB > 1 ? rowsB : 1), "sgemm: ldb=%lld < %lld", (long long)ldb, (long long)rowsB);
  CANDMC_CHECK(ldc >= (m > 1 ? m : 1), "sgemm: ldc=%lld < %lld", (long long)ldc, (long long)m);
  if (m == 0 || n == 0) return OK;
  const int M = (int)m, N = (int)n, K = (int)k;
  if (k == 0 || alpha == 0.0f) {
    if (beta == 1.0f) return OK;
    int grid = (int)((m * n + 255) / 256);
    const int cap = runtime().num_sms * 8;
    if (grid > cap) grid = cap;
    scale_c_f32_kernel<<<grid, 256, 0, stream>>>(M, N, beta, C, ldc);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
    return OK;
  }
  // K-major operands for the tensor core: op(A) as (k x m, k contiguous), op(B) as (k x n, k contiguous).  'T' for A and 'N'
  // for B are that already when TMA can read them in place (16-byte aligned base and pitch); everything else goes through
  // one pack kernel into the workspace.
  auto tma_ready = [](const float* p, int64_t ld) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % 4 == 0; };
  const bool packA = !tA || !tma_ready(A, lda), packB = tB || !tma_ready(B, ldb);
  const int64_t kp = (k + 3) / 4 * 4;   // pitch of the packed copies
  const float* Ak = A;
  const float* Bk = B;
  int64_t ldak = lda, ldbk = ldb;
  if (packA || packB) {
    void* wsv = nullptr;
    CANDMC_TRY(workspace_get(sizeof(float) * kp * ((packA ? m : 0) + (packB ? n : 0)) + 64, &wsv));
    float* ws = static_cast<float*>(wsv);
    const int cap = runtime().num_sms * 8;
    if (packA) {
      const int64_t tiles = ((k + 63) / 64) * ((m + 63) / 64);
      const int grid = (int)(tiles < cap ? tiles : cap);
      if (tA) pack_kmajor_f32_kernel<false><<<grid, 256, 0, stream>>>(K, M, A, lda, ws, kp);
      else pack_kmajor_f32_kernel<true><<<grid, 256, 0, stream>>>(K, M, A, lda, ws, kp);
      CANDMC_CUDA(cudaGetLastError());
      runtime().launches++;
      Ak = ws;
      ldak = kp;
      ws += kp * m;
    }
    if (packB) {
      const int64_t tiles = ((k + 63) / 64) * ((n + 63) / 64);
      const int grid = (int)(tiles < cap ? tiles : cap);
      if (!tB) pack_kmajor_f32_kernel<false><<<grid, 256, 0, stream>>>(K, N, B, ldb, ws, kp);
      else pack_kmajor_f32_kernel<true><<<grid, 256, 0, stream>>>(K, N, B, ldb, ws, kp);
      CANDMC_CUDA(cudaGetLastError());
      runtime().launches++;
      Bk = ws;
      ldbk = kp;
    }
  }
  CUtensorMap tmA, tmB;
  CANDMC_TRY(encode_tmap_f32(&tmA, Ak, k, m, ldak, FK, FM));
  CANDMC_TRY(encode_tmap_f32(&tmB, Bk, k, n, ldbk, FK, FN));
  if (g_f32_mode == 3) return launch_f32<true>(tmA, tmB, C, ldc, M, N, K, alpha, beta, stream);
  return launch_f32<false>(tmA, tmB, C, ldc, M, N, K, alpha, beta, stream);
}

}  // namespace candmc

extern "C" {

int candmc_set_f32_mode(int tf32_products) {
  CANDMC_CHECK(tf32_products == 1 || tf32_products == 3, "candmc_set_f32_mode: 3 (FP32-level accuracy) or 1 (TF32 accuracy)");
  candmc::g_f32_mode = tf32_products;
  return candmc::OK;
}

int candmc_sgemm(char transa, char transb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* stream) {
  CANDMC_TRY(candmc::runtime_require());
  CANDMC_CHECK(m <= 0 || n <= 0 || (candmc::is_device_ptr(A) && candmc::is_device_ptr(B) && candmc::is_device_ptr(C)),
               "sgemm: operands must be device pointers");
  return candmc::gemm_f32(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
