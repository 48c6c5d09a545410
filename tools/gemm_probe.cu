// gemm_probe — native correctness + speed probe for libcandmc_b200.so's local kernels, through the C ABI.
// cuBLAS is linked HERE ONLY, as the cross-check the north-star allows; it is never part of the library.
//   gemm_probe check            : shape/transpose/alpha/beta sweep vs a host long-double reference and cuBLAS
//   gemm_probe speed [n ...]    : TFLOP/s of candmc_dgemm vs cublasDgemm, device-timed (CUDA events)
//   gemm_probe pack             : GB/s of lda_cpy / scaled lda_cpy / transpose
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "candmc_b200.h"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)
#define CC(x)                                                                  \
  do {                                                                         \
    int r = (x);                                                               \
    if (r != 0) {                                                              \
      printf("candmc error %d: %s (%s:%d)\n", r, candmc_last_error(), __FILE__, __LINE__); \
      exit(3);                                                                 \
    }                                                                          \
  } while (0)

static double urand() { return drand48(); }

static void host_gemm(char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B,
                      int ldb, double beta, double* C, int ldc) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      long double s = 0;
      for (int p = 0; p < k; ++p) {
        double a = (ta == 'N') ? A[i + (size_t)p * lda] : A[p + (size_t)i * lda];
        double b = (tb == 'N') ? B[p + (size_t)j * ldb] : B[j + (size_t)p * ldb];
        s += (long double)a * b;
      }
      double c0 = (beta == 0.0) ? 0.0 : beta * C[i + (size_t)j * ldc];
      C[i + (size_t)j * ldc] = (double)(alpha * s) + c0;
    }
}

static int check_one(char ta, char tb, int m, int n, int k, double alpha, double beta, int pad, int misalign,
                     bool force_generic) {
  const int rowsA = (ta == 'N') ? m : k, colsA = (ta == 'N') ? k : m;
  const int rowsB = (tb == 'N') ? k : n, colsB = (tb == 'N') ? n : k;
  const int lda = (rowsA > 0 ? rowsA : 1) + pad, ldb = (rowsB > 0 ? rowsB : 1) + pad, ldc = m + pad;
  std::vector<double> A((size_t)lda * colsA + 2), B((size_t)ldb * colsB + 2), C((size_t)ldc * n + 2), R;
  for (auto& x : A) x = urand() - 0.5;
  for (auto& x : B) x = urand() - 0.5;
  for (auto& x : C) x = (beta == 0.0) ? NAN : urand() - 0.5;  // beta==0 must not read C
  R = C;
  double *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 8 + 16));
  CK(cudaMalloc(&dB, B.size() * 8 + 16));
  CK(cudaMalloc(&dC, C.size() * 8 + 16));
  double* pA = dA + misalign;
  double* pB = dB + misalign;
  CK(cudaMemcpy(pA, A.data(), A.size() * 8 - 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pB, B.data(), B.size() * 8 - 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
  candmc_debug_force_generic_gemm(force_generic);
  CC(candmc_dgemm(ta, tb, m, n, k, alpha, pA, lda, pB, ldb, beta, dC, ldc, 0));
  candmc_debug_force_generic_gemm(0);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
  host_gemm(ta, tb, m, n, k, alpha, A.data(), lda, B.data(), ldb, beta, R.data(), ldc);
  double d2 = 0, r2 = 0;
  int bad_pad = 0;
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) {
      double d = C[i + (size_t)j * ldc] - R[i + (size_t)j * ldc];
      if (!(d == d)) d = 1e300;
      d2 += d * d;
      r2 += R[i + (size_t)j * ldc] * R[i + (size_t)j * ldc];
    }
    for (int i = m; i < ldc; ++i) {  // padding rows of C must be untouched
      double a = C[i + (size_t)j * ldc], b = R[i + (size_t)j * ldc];
      if (memcmp(&a, &b, 8) != 0) bad_pad++;
    }
  }
  const double rel = (r2 > 0) ? sqrt(d2 / r2) : sqrt(d2);
  const double tol = 10.0 * (k > 0 ? k : 1) * 2.220446049250313e-16;
  const int ok = (rel <= tol) && bad_pad == 0;
  if (!ok)
    printf("FAIL %c%c m=%d n=%d k=%d alpha=%g beta=%g pad=%d mis=%d generic=%d rel=%.3e tol=%.3e bad_pad=%d\n", ta, tb,
           m, n, k, alpha, beta, pad, misalign, (int)force_generic, rel, tol, bad_pad);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dC);
  return ok;
}

static int run_check() {
  int total = 0, pass = 0;
  const char tr[2] = {'N', 'T'};
  const int shapes[][3] = {{128, 128, 16},  {128, 128, 128}, {256, 384, 64}, {1, 1, 1},      {7, 5, 3},
                           {130, 70, 33},   {64, 200, 100},  {257, 129, 17}, {100, 300, 250}, {512, 96, 40},
                           {31, 1000, 8},   {1000, 31, 9},   {384, 256, 512}, {128, 128, 0},  {333, 222, 111},
                           {128, 128, 4096}, {250, 256, 3000}};
  for (auto& s : shapes)
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int v = 0; v < 3; ++v) {
          const double alpha = (v == 0) ? 1.0 : (v == 1 ? 1.2 : -1.0);
          const double beta = (v == 0) ? 0.0 : (v == 1 ? 0.8 : 1.0);
          for (int pad = 0; pad <= 2; pad += 2) {
            total++;
            pass += check_one(tr[a], tr[b], s[0], s[1], s[2], alpha, beta, pad, 0, false);
          }
          // unaligned operands -> generic path must be chosen automatically
          total++;
          pass += check_one(tr[a], tr[b], s[0], s[1], s[2], alpha, beta, 1, 1, false);
          total++;
          pass += check_one(tr[a], tr[b], s[0], s[1], s[2], alpha, beta, 0, 0, true);
        }
  printf("{\"probe\":\"check\",\"cases\":%d,\"passed\":%d}\n", total, pass);
  return pass == total ? 0 : 1;
}

static float time_loop(int iters, cudaStream_t st, void (*fn)(void*), void* ctx) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) fn(ctx);
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / iters;
}

struct GemmCtx {
  char ta, tb;
  int m, n, k;
  double *A, *B, *C;
  int lda, ldb, ldc;
  cublasHandle_t h;
  double beta;
};
static void run_ours(void* p) {
  GemmCtx* c = (GemmCtx*)p;
  CC(candmc_dgemm(c->ta, c->tb, c->m, c->n, c->k, 1.0, c->A, c->lda, c->B, c->ldb, c->beta, c->C, c->ldc, 0));
}
static void run_cublas(void* p) {
  GemmCtx* c = (GemmCtx*)p;
  const double one = 1.0;
  cublasDgemm(c->h, c->ta == 'N' ? CUBLAS_OP_N : CUBLAS_OP_T, c->tb == 'N' ? CUBLAS_OP_N : CUBLAS_OP_T, c->m, c->n,
              c->k, &one, c->A, c->lda, c->B, c->ldb, &c->beta, c->C, c->ldc);
}

static void speed_one(cublasHandle_t h, char ta, char tb, int m, int n, int k, int iters, double beta = 0.0,
                      const char* tag = "speed") {
  GemmCtx c;
  c.ta = ta; c.tb = tb; c.m = m; c.n = n; c.k = k; c.h = h; c.beta = beta;
  const int rowsA = (ta == 'N') ? m : k, colsA = (ta == 'N') ? k : m;
  const int rowsB = (tb == 'N') ? k : n, colsB = (tb == 'N') ? n : k;
  c.lda = rowsA; c.ldb = rowsB; c.ldc = m;
  double* ref;
  CK(cudaMalloc(&c.A, (size_t)rowsA * colsA * 8));
  CK(cudaMalloc(&c.B, (size_t)rowsB * colsB * 8));
  CK(cudaMalloc(&c.C, (size_t)m * n * 8));
  CK(cudaMalloc(&ref, (size_t)m * n * 8));
  CC(candmc_fill_drand48(c.A, rowsA, colsA, c.lda, 0, 0, rowsA, 0, 0));
  CC(candmc_fill_drand48(c.B, rowsB, colsB, c.ldb, 0, 0, rowsB, 1, 0));
  // correctness vs cuBLAS at full size (beta != 0: both start from the same C)
  if (beta != 0.0) {
    CC(candmc_fill_drand48(c.C, m, n, c.ldc, 0, 0, m, 0, 0));
    CK(cudaMemcpy(ref, c.C, (size_t)m * n * 8, cudaMemcpyDeviceToDevice));
  }
  run_ours(&c);
  double* keep = c.C;
  c.C = ref;
  run_cublas(&c);
  c.C = keep;
  double fr[2];
  CC(candmc_frob_diff(c.C, c.ldc, ref, m, m, n, fr, 0));
  const double rel = sqrt(fr[0] / fr[1]);
  for (int w = 0; w < 2; ++w) run_ours(&c);
  const float ms_ours = time_loop(iters, 0, run_ours, &c);
  for (int w = 0; w < 2; ++w) run_cublas(&c);
  const float ms_cublas = time_loop(iters, 0, run_cublas, &c);
  const double fl = 2.0 * m * n * (double)k;
  printf("{\"probe\":\"%s\",\"trans\":\"%c%c\",\"m\":%d,\"n\":%d,\"k\":%d,\"beta\":%g,\"ms\":%.4f,\"tflops\":%.3f,"
         "\"cublas_ms\":%.4f,\"cublas_tflops\":%.3f,\"rel_frob_vs_cublas\":%.3e,\"tol_10_k_eps\":%.3e}\n",
         tag, ta, tb, m, n, k, beta, ms_ours, fl / ms_ours * 1e-9, ms_cublas, fl / ms_cublas * 1e-9, rel,
         10.0 * k * 2.220446049250313e-16);
  fflush(stdout);
  cudaFree(c.A); cudaFree(c.B); cudaFree(c.C); cudaFree(ref);
}

struct PackCtx {
  int kind;
  int64_t rows, cols, lda, ldb;
  double *A, *B;
};
static void run_pack(void* p) {
  PackCtx* c = (PackCtx*)p;
  if (c->kind == 0) CC(candmc_lda_cpy(c->rows, c->cols, c->lda, c->ldb, c->A, c->B, 0));
  if (c->kind == 1) CC(candmc_lda_cpy_scaled(c->rows, c->cols, c->lda, c->ldb, c->A, c->B, 0.5, 0.25, 0));
  if (c->kind == 2) CC(candmc_transpose(c->rows, c->cols, c->A, c->lda, c->B, c->ldb, 0));
}

static void pack_speed(int64_t n) {
  // n = 8192: 512 MiB per matrix (config 2's block), larger than L2
  PackCtx c;
  c.rows = n; c.cols = n; c.lda = 2 * n; c.ldb = n;
  CK(cudaMalloc(&c.A, (size_t)c.lda * n * 8));
  CK(cudaMalloc(&c.B, (size_t)n * n * 8));
  CC(candmc_fill_drand48(c.A, c.lda, n, c.lda, 0, 0, c.lda, 0, 0));
  const char* names[3] = {"lda_cpy", "lda_cpy_scaled", "transpose"};
  const double bytes_per_elem[3] = {16, 24, 16};
  for (int kind = 0; kind < 3; ++kind) {
    c.kind = kind;
    for (int w = 0; w < 3; ++w) run_pack(&c);
    const float ms = time_loop(10, 0, run_pack, &c);
    printf("{\"probe\":\"pack\",\"kernel\":\"%s\",\"rows\":%lld,\"cols\":%lld,\"ms\":%.4f,\"GBps\":%.1f}\n", names[kind],
           (long long)n, (long long)n, ms, bytes_per_elem[kind] * n * n / ms * 1e-6);
  }
  cudaFree(c.A); cudaFree(c.B);
}

int main(int argc, char** argv) {
  CC(candmc_init(0));
  const char* mode = argc > 1 ? argv[1] : "check";
  if (!strcmp(mode, "check")) return run_check();
  if (!strcmp(mode, "pack")) {
    pack_speed(argc > 2 ? atoll(argv[2]) : 8192);
    return 0;
  }
  if (!strcmp(mode, "one")) {  // gemm_probe one TA TB m n k beta [reserve_sms] [prefetch_c] [iters]: one shape, for ncu captures
    if (argc < 8) {
      printf("usage: gemm_probe one TA TB m n k beta [reserve_sms] [prefetch_c] [iters]\n");
      return 1;
    }
    cublasHandle_t h;
    cublasCreate(&h);
    if (argc > 8) candmc_debug_gemm_reserve_sms(atoi(argv[8]));
    if (argc > 9) candmc_debug_prefetch_c(atoi(argv[9]));
    speed_one(h, argv[2][0], argv[3][0], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), argc > 10 ? atoi(argv[10]) : 1, atof(argv[7]), "one");
    cublasDestroy(h);
    return 0;
  }
  if (!strcmp(mode, "sched")) {  // A/B: static round-robin vs dynamic (atomic counter) tile schedule
    cublasHandle_t h;
    cublasCreate(&h);
    for (int rep = 0; rep < 2; ++rep)
      for (int st = 1; st >= 0; --st) {
        candmc_debug_static_schedule(st);
        printf("{\"probe\":\"sched\",\"static\":%d}\n", st);
        speed_one(h, 'N', 'N', 4096, 4096, 4096, 5);
        speed_one(h, 'N', 'N', 8192, 8192, 8192, 5);
      }
    candmc_debug_static_schedule(0);
    cublasDestroy(h);
    return 0;
  }
  if (!strcmp(mode, "chunk")) {
    // the launches a SUMMA / 2.5D sweep is made of: one k-chunk of a panel accumulated onto C (beta = 1), with and without the
    // SMs the sweep leaves to NCCL, with and without the L2 prefetch of the C tile; plus the CAQR update's second GEMM
    cublasHandle_t h;
    cublasCreate(&h);
    const int shapes[][3] = {{16384, 16384, 2048}, {8192, 8192, 1024}, {12288, 12288, 1536}, {65536, 8192, 512}, {16384, 16384, 14336}};
    for (int pf = 0; pf <= 1; ++pf)
      for (int rs = 0; rs <= 2; rs += 2) {
        candmc_debug_prefetch_c(pf);
        candmc_debug_gemm_reserve_sms(rs);
        char tag[64];
        snprintf(tag, sizeof tag, "chunk_pf%d_reserve%d", pf, rs);
        for (auto& s : shapes) {
          if (s[2] == 14336 && (rs == 0 || pf == 0)) continue;
          speed_one(h, 'N', 'N', s[0], s[1], s[2], s[2] > 4096 ? 2 : 5, 1.0, tag);
          if (pf == 0) speed_one(h, 'N', 'N', s[0], s[1], s[2], s[2] > 4096 ? 2 : 5, 0.0, tag);
        }
      }
    candmc_debug_prefetch_c(0);
    candmc_debug_gemm_reserve_sms(0);
    cublasDestroy(h);
    return 0;
  }
  if (!strcmp(mode, "tile")) {
    // A/B of the two CTA shapes (128 x 128 tiles, one CTA per SM / 128 x 64 tiles, two CTAs per SM) at the shapes the sweeps and
    // the CAQR update launch, plus square ones down to where tiles stop filling waves
    cublasHandle_t h;
    cublasCreate(&h);
    struct Sh { char ta, tb; int m, n, k; double beta; int reserve; };
    const Sh shapes[] = {{'N', 'N', 16384, 16384, 2048, 1.0, 2}, {'N', 'N', 16384, 16384, 2048, 0.0, 2}, {'N', 'N', 8192, 8192, 1024, 1.0, 2},
                         {'N', 'N', 8192, 8192, 7168, 1.0, 2},   {'N', 'N', 12288, 12288, 1536, 1.0, 0}, {'N', 'T', 12288, 12288, 12288, 1.0, 0},
                         {'N', 'N', 65536, 8192, 512, 1.0, 0},   {'N', 'N', 65536, 8192, 512, 0.0, 0},   {'N', 'N', 32768, 4096, 512, 1.0, 0},
                         {'N', 'N', 1024, 1024, 1024, 0.0, 0},   {'N', 'N', 2048, 2048, 2048, 0.0, 0},   {'N', 'N', 4096, 4096, 4096, 0.0, 0},
                         {'N', 'N', 8192, 8192, 8192, 0.0, 0},   {'T', 'N', 8192, 8192, 8192, 0.0, 0},   {'N', 'N', 16384, 16384, 16384, 0.0, 0}};
    for (auto& sh : shapes)
      for (int tile : {128, 64}) {
        candmc_debug_gemm_tile(tile);
        candmc_debug_gemm_reserve_sms(sh.reserve);
        char tag[64];
        snprintf(tag, sizeof tag, "tile%d_reserve%d", tile, sh.reserve);
        speed_one(h, sh.ta, sh.tb, sh.m, sh.n, sh.k, sh.k > 8192 ? 2 : 5, sh.beta, tag);
      }
    candmc_debug_gemm_tile(0);
    candmc_debug_gemm_reserve_sms(0);
    cublasDestroy(h);
    return 0;
  }
  if (!strcmp(mode, "speed")) {
    cublasHandle_t h;
    cublasCreate(&h);
    std::vector<int> ns;
    for (int i = 2; i < argc; ++i) ns.push_back(atoi(argv[i]));
    if (ns.empty()) ns = {1024, 2048, 4096, 8192, 16384};
    for (int n : ns) speed_one(h, 'N', 'N', n, n, n, n >= 16384 ? 3 : 5);
    speed_one(h, 'N', 'T', 8192, 8192, 8192, 5);
    speed_one(h, 'T', 'N', 8192, 8192, 8192, 5);
    speed_one(h, 'T', 'T', 8192, 8192, 8192, 5);
    // config 5 shapes (CAQR trailing update): W = Y^T A (512 x 8192 x 65536) and A -= Y W (65536 x 8192 x 512)
    speed_one(h, 'T', 'N', 512, 8192, 65536, 5);
    speed_one(h, 'N', 'N', 65536, 8192, 512, 5);
    cublasDestroy(h);
    return 0;
  }
  printf("unknown mode %s\n", mode);
  return 1;
}
