/* candmc/lu_offload.h — the accelerator seam of the 2.5D LU with the reference's names and argument order
 * (alg/LU/lu_offload.h:19-101), implemented on a B200 by libcandmc_lu_offload.so (candmc_b200/csrc/lu_offload_shim.cxx
 * over the C ABI candmc_off_* of include/candmc_b200.h).
 *
 * Use: compile the reference's alg/LU sources with -DOFFLOAD -DOFFLOAD_FAT_GEMM as they are (they include their own
 * copy of this interface) and link libcandmc_lu_offload.so INSTEAD of alg/LU/lu_offload.cxx.  The offloaded matrices
 * then live in HBM, offload_gemm_A runs the TMA + DMMA GEMM asynchronously and wait_gemm joins it.
 *
 * Differences from the reference's host fallback, all at the edges:
 *   - get_mat_handle(m) returns a pinned host mirror with the matrix's current contents; data written through it is
 *     uploaded before the next operation on m (that covers the reference's only use, the initial memcpy of the local
 *     matrix, lu_25d_pvt.cxx:1600-1602).  It is not a live alias of device memory.
 *   - alloc_A(size, ptr) uploads ptr when it is not NULL (the reference does so on its accelerator build only).
 *   - a failed precondition or CUDA error prints a message and ends the process (ABORT semantics, util.h:127-138).
 */
#ifndef CANDMC_LU_OFFLOAD_H
#define CANDMC_LU_OFFLOAD_H

#include <stdint.h>

enum OFF_MAT { OFF_A, OFF_L, OFF_U };

double* get_mat_handle(OFF_MAT omat);

void set_mic_rank(int mic_rank);

void wait_gemm();

void offload_gemm_A(char tA, char tB, int m, int n, int k, double alpha, int offset_A, OFF_MAT omat_A, int lda_A,
                    int offset_B, OFF_MAT omat_B, int lda_B, double beta, int offset_C, OFF_MAT omat_C, int lda_C);

void download_lda_cpy(int nrow, int ncol, int lda_A, int lda_B, int offset_A, double* B, OFF_MAT omat_A);

void upload_lda_cpy(int nrow, int ncol, int lda_A, int lda_B, double const* A, int offset_B, OFF_MAT omat_B);

void offload_sparse_rw(int nrow, int ncol, int lda_B, double* A, int lda_A, int* offsets_transfer, OFF_MAT omat_B,
                       char rw);

void alloc_A(int64_t size, double* ptr);
void alloc_L(int64_t size);
void alloc_U(int64_t size);
void alloc_transfer(int64_t size);
void free_offload_A();
void free_offload_L();
void free_offload_U();
void free_offload_transfer();

#endif
