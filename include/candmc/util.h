/* candmc/util.h — the pieces of alg/shared/util.h the CANMM path and its drivers use: lda_cpy (both overloads,
 * util.h:459-501), ABORT (:135-137), WRAP (:140-142), ALIGN_BYTES (:144-146), MIN/MAX and the no-op profiling /
 * debug macros the drivers mention.  lda_cpy<double> runs on the GPU (device or host pointers, through the C ABI);
 * other element types are plain host copies like the reference.
 */
#ifndef CANDMC_UTIL_H
#define CANDMC_UTIL_H

#include <assert.h>
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../candmc_b200.h"

typedef int64_t long_int;

#ifndef ABORT
#define ABORT                  \
  do {                         \
    candmc_shim_abort(__FILE__, __LINE__); \
  } while (0)
#endif
#ifndef WRAP
#define WRAP(a, b) ((a + b) % b)
#endif
#ifndef ALIGN_BYTES
#define ALIGN_BYTES 16
#endif
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef LIBT_ASSERT
#define LIBT_ASSERT(...) assert(__VA_ARGS__)
#endif

/* profiling / debug macro families of util.h:172-384 — timing here is CUDA events + ncu (bench.py), so these are no-ops */
#ifndef TAU_FSTART
#define TAU_FSTART(ARG)
#define TAU_FSTOP(ARG)
#define TAU_PROFILE_TIMER(ARG1, ARG2, ARG3, ARG4)
#define TAU_PROFILE_INIT(argc, argv)
#define TAU_PROFILE_SET_NODE(ARG)
#define TAU_PROFILE_START(ARG)
#define TAU_PROFILE_STOP(ARG)
#define TAU_PROFILE_SET_CONTEXT(ARG)
#endif
#ifndef DEBUG_PRINTF
#define DEBUG_PRINTF(...) do {} while (0)
#endif
#ifndef RANK_PRINTF
#define RANK_PRINTF(...) do {} while (0)
#endif
#ifndef DPRINTF
#define DPRINTF(...) do {} while (0)
#endif
#ifndef VPRINTF
#define VPRINTF(...) do {} while (0)
#endif

extern "C" void candmc_shim_abort(const char* file, int line);
extern "C" void candmc_shim_check(int status, const char* what);

void print_matrix(double const* M, int n, int m);
void print_matrix(double const* M, int n, int m, int lda);

/* lda_cpy, util.h:459-471 */
template <typename dtype>
inline void lda_cpy(const int nrow, const int ncol, const int lda_A, const int lda_B, const dtype* A, dtype* B) {
  if (lda_A == nrow && lda_B == nrow) {
    memcpy(B, A, (size_t)nrow * ncol * sizeof(dtype));
  } else {
    for (int i = 0; i < ncol; i++) memcpy(B + (size_t)lda_B * i, A + (size_t)lda_A * i, nrow * sizeof(dtype));
  }
}
template <>
inline void lda_cpy<double>(const int nrow, const int ncol, const int lda_A, const int lda_B, const double* A,
                            double* B) {
  candmc_shim_check(candmc_lda_cpy(nrow, ncol, lda_A, lda_B, A, B, 0), "lda_cpy");
}

/* scaled lda_cpy, util.h:484-501: B = B*b + A*a */
template <typename dtype>
inline void lda_cpy(const int nrow, const int ncol, const int lda_A, const int lda_B, const dtype* A, dtype* B,
                    const dtype a, const dtype b) {
  for (int i = 0; i < ncol; i++)
    for (int j = 0; j < nrow; j++) B[(size_t)lda_B * i + j] = B[(size_t)lda_B * i + j] * b + A[(size_t)lda_A * i + j] * a;
}
template <>
inline void lda_cpy<double>(const int nrow, const int ncol, const int lda_A, const int lda_B, const double* A,
                            double* B, const double a, const double b) {
  candmc_shim_check(candmc_lda_cpy_scaled(nrow, ncol, lda_A, lda_B, A, B, a, b, 0), "lda_cpy(scaled)");
}

#endif
