/* TEST INFRASTRUCTURE (oracle/): the four ScaLAPACK wrappers alg/SE/dmatrix.cxx references.  The image has no ScaLAPACK /
 * BLACS; the DMatrix pack operations the oracle is pinned to (replicate_*, reduce_scatter_horizontal, transpose_data,
 * foldcols / foldrows, slice, get_contig) never call them — only the constructor fills a descriptor. */
#include <stdio.h>
#include <stdlib.h>
void cdescinit(int* desc, const int m, const int n, const int mb, const int nb, const int irsrc, const int icsrc,
               const int ictxt, const int LLD, int* info) {
  desc[0] = 1; desc[1] = ictxt; desc[2] = m; desc[3] = n; desc[4] = mb; desc[5] = nb; desc[6] = irsrc; desc[7] = icsrc;
  desc[8] = LLD; *info = 0;
}
static void no_scalapack(const char* what) {
  fprintf(stderr, "oracle: %s needs ScaLAPACK, which this image does not have\n", what);
  abort();
}
void cpdgemm(char, char, int, int, int, double, double*, int, int, int*, double*, int, int, int*, double, double*, int, int,
             int*) { no_scalapack("cpdgemm"); }
void cpdsyrk(char, char, int, int, double, const double*, int, int, int*, double, double*, int, int, int*) { no_scalapack("cpdsyrk"); }
void cpdtrsm(char, char, char, char, int, int, double, const double*, int, int, int*, double*, int, int, int*) { no_scalapack("cpdtrsm"); }
