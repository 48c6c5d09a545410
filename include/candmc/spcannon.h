/* candmc/spcannon.h — split-dimensional Cannon with the reference's signatures
 * (alg/MM/splitdim_cannon/spcannon.h:31-59).  Like the reference the `comm` argument is informational: the
 * algorithm runs on the world communicator (spcannon.cxx:270-273 creates its windows on MPI_COMM_WORLD).  Unlike the
 * reference, A and B are preserved. */
#ifndef CANDMC_SPCANNON_H
#define CANDMC_SPCANNON_H

#include "mpi.h"

void kput_cannon(int const rank, int const kary, int const ndim, MPI_Comm const comm, int const n, int const m,
                 int const k, char const transp_A, double const alpha, double* A, char const transp_B,
                 double const beta, double* B, double* C);

void kuni_cannon(int const rank, int const kary, int const ndim, MPI_Comm const comm, int const n, int const m,
                 int const k, char const transp_A, double const alpha, double* A, char const transp_B,
                 double const beta, double* B, double* C);

#endif
