/* candmc/lapack.h — cdgemm with the reference's by-value signature (alg/shared/lapack.h:10-16, lapack.cxx:425-434),
 * executed by the sm_100a DMMA kernel.  A, B, C may be host or device pointers.  The other LAPACK pass-through
 * wrappers of alg/shared/lapack.h are outside the CANMM hot path. */
#ifndef CANDMC_LAPACK_H
#define CANDMC_LAPACK_H

void cdgemm(char transa, char transb, int m, int n, int k, double a, const double* A, int lda, const double* B, int ldb,
            double b, double* C, int ldc);

/* Single-precision companion of cdgemm (no reference counterpart: the reference multiplies in double only): same by-value
 * argument list with float data, DEVICE pointers, executed by the tcgen05 kind::tf32 kernel with split operands
 * (candmc_sgemm in candmc_b200.h states the error bound). */
void csgemm(char transa, char transb, int m, int n, int k, float a, const float* A, int lda, const float* B, int ldb, float b,
            float* C, int ldc);

#endif
