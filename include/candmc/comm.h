/* candmc/comm.h — processor-grid descriptors and communicator macros with the reference's names and semantics
 * (alg/shared/comm.h:32-201).  CommData_t is passed BY VALUE to the multiplies, exactly as in the reference; its `cm`
 * is an MPI_Comm, which in this build is a candmc_comm_t* (an NCCL communicator on this process's GPU).
 */
#ifndef CANDMC_COMM_H
#define CANDMC_COMM_H

#include <assert.h>
#include <stdint.h>

#include "mpi.h"
#include "util.h"

/* comm.h:32-63 (the unused alpha-beta cost-model methods are not reproduced) */
typedef class CommData {
 public:
  MPI_Comm cm;
  int np;
  int rank;
  int color;
  int alive;
} CommData_t;

/* 2d processor grid local processor view, comm.h:66-84 */
class pview {
 public:
  int rrow;          /* current root row */
  int rcol;          /* current root col */
  CommData_t crow;   /* row communicator */
  CommData_t ccol;   /* column communicator */
  CommData_t cdiag;  /* diagonal communicator */
  CommData_t cworld; /* world communicator */
};

/* 3d processor grid local processor view, comm.h:88-101 */
class pview_3d {
 public:
  pview prect;
  pview plyr;
  CommData_t clyr;
  CommData_t cworld;
};

/* comm.h:110-114: a blocking broadcast of a HOST or DEVICE buffer of doubles; WAIT_BCAST is a no-op there too */
#define POST_BCAST(buf, sz, type, root, cdt, bcast_req) \
  do {                                                  \
    MPI_Bcast(buf, sz, type, root, cdt.cm);             \
  } while (0)
#define WAIT_BCAST(cdt, bcast_req)

#define SET_COMM(_cm, _rank, _np, _cdt) \
  do {                                  \
    _cdt.cm = _cm;                      \
    _cdt.rank = _rank;                  \
    _cdt.np = _np;                      \
    _cdt.alive = 1;                     \
  } while (0)

#define RINIT_COMM(numPes, myRank, nr, nb, cdt) \
  do {                                          \
    INIT_COMM(numPes, myRank, nr, cdt);         \
  } while (0)

#define INIT_COMM(numPes, myRank, nr, cdt)         \
  do {                                             \
    MPI_Init(&argc, &argv);                        \
    MPI_Comm_size(MPI_COMM_WORLD, &numPes);        \
    MPI_Comm_rank(MPI_COMM_WORLD, &myRank);        \
    SET_COMM(MPI_COMM_WORLD, myRank, numPes, cdt); \
  } while (0)

#define COMM_EXIT   \
  do {              \
    MPI_Finalize(); \
  } while (0)

#define SETUP_SUB_COMM(cdt_master, cdt, commrank, bcolor, p)      \
  do {                                                            \
    cdt.rank = commrank;                                          \
    cdt.np = p;                                                   \
    cdt.color = bcolor;                                           \
    cdt.alive = 1;                                                \
    MPI_Comm_split(cdt_master.cm, bcolor, commrank, &cdt.cm);     \
  } while (0)

#define SETUP_SUB_COMM_SHELL(cdt_master, cdt, commrank, bcolor, p) \
  do {                                                             \
    cdt.rank = commrank;                                           \
    cdt.np = p;                                                    \
    cdt.color = bcolor;                                            \
    cdt.alive = 0;                                                 \
  } while (0)

#define SHELL_SPLIT(cdt_master, cdt)                              \
  do {                                                            \
    cdt.alive = 1;                                                \
    MPI_Comm_split(cdt_master.cm, cdt.color, cdt.rank, &cdt.cm);  \
  } while (0)

#define RSETUP_KDIR_COMM(myRank, p, c, cdt, commrank, color)        \
  do {                                                              \
    commrank = myRank / (p / c);                                    \
    color = myRank % (p / c);                                       \
    cdt.rank = commrank;                                            \
    cdt.np = c;                                                     \
    MPI_Comm_split(MPI_COMM_WORLD, color, commrank, &(cdt.cm));     \
  } while (0)

/* like the reference (comm.h:183-195) this uses the caller's variables literally named myRow / myCol */
#define RSETUP_LAYER_COMM(pesdim, commrank, color, cdt_row, cdt_col, row, col)      \
  do {                                                                              \
    MPI_Comm MPI_INTRALAYER_COMM;                                                   \
    MPI_Comm_split(MPI_COMM_WORLD, commrank, color, &MPI_INTRALAYER_COMM);          \
    row = color / pesdim;                                                           \
    col = color % pesdim;                                                           \
    MPI_Comm_split(MPI_INTRALAYER_COMM, myRow, myCol, &(cdt_row.cm));               \
    MPI_Comm_split(MPI_INTRALAYER_COMM, myCol, myRow, &(cdt_col.cm));               \
    cdt_row.np = pesdim;                                                            \
    cdt_row.rank = col;                                                             \
    cdt_col.np = pesdim;                                                            \
    cdt_col.rank = row;                                                             \
  } while (0)

#define FREE_CDT(cdt)           \
  do {                          \
    MPI_Comm_free(&(cdt->cm));  \
  } while (0)

#endif
