// candmc_b200 — block-cyclic <-> blocked redistribution: the index plan shared by host code, kernels and CPU tests.
//
// Layouts of an m x n FP64 matrix on an nprow x npcol grid (both give every rank an (m/nprow) x (n/npcol) local piece):
//   block-cyclic, block nb, root (rrow, rcol): what the reference's QR / SE drivers use — global block row I lives on grid
//     row (I + rrow) mod nprow at local block row I div nprow (test/QR/test_qr_2d.cxx:87-94, alg/QR/qr_2d/qr_2d.cxx:140-147,
//     alg/SE/dmatrix.cxx:194-203); columns likewise.
//   blocked: what CANMM uses — grid row i owns rows [i*m/nprow, (i+1)*m/nprow) (test/MM/topo_pdgemm_unit.cxx:250-256).
// The 2-D redistribution is two 1-D exchanges (rows over the column communicator, columns over the row communicator).
// On one axis with P ranks, K = (local extent / nb) blocks per rank and global block index I in [0, P*K):
//   the blocks rank `me` holds CYCLICALLY and blocked-owner p needs form ONE contiguous local range  [lo[p], lo[p]+ccnt[p]),
//   the blocks rank `me` holds BLOCKED   and cyclic-owner  p needs form a STRIDED set  first[p] + t*P, t < scnt[p].
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CANDMC_HD __host__ __device__ __forceinline__
#else
#define CANDMC_HD inline
#endif

namespace candmc {

constexpr int REDIST_MAX_P = 64;

struct AxisPlan {
  int P;                          // ranks on the axis
  int me;                         // my rank on the axis
  int root;                       // cyclic root: rank `root` holds global block 0
  int nb;                         // block size (elements)
  int64_t K;                      // blocks per rank
  unsigned base_mod;              // (me*K + root) mod P: cyclic owner of my blocked-local block blk is (base_mod + blk) mod P
  int lo[REDIST_MAX_P];           // contiguous side: first local (cyclic) block exchanged with p
  int ccnt[REDIST_MAX_P];         //                  number of blocks
  int first[REDIST_MAX_P];        // strided side: first local (blocked) block exchanged with p (stride P)
  int scnt[REDIST_MAX_P];         //               number of blocks
  int64_t coff[REDIST_MAX_P];     // exclusive prefix sums of ccnt / scnt (in blocks): segment starts
  int64_t soff[REDIST_MAX_P];
};

inline int64_t ceil_div_i64(int64_t a, int64_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

// Fills `pl`; returns false on bad arguments.
inline bool axis_plan(int P, int me, int root, int64_t K, int nb, AxisPlan* pl) {
  if (P < 1 || P > REDIST_MAX_P || me < 0 || me >= P || root < 0 || root >= P || K < 0 || nb < 1 ||
      K > (int64_t)0x3fffffff)
    return false;
  pl->P = P;
  pl->me = me;
  pl->root = root;
  pl->nb = nb;
  pl->K = K;
  pl->base_mod = (unsigned)(((int64_t)me * K + root) % P);
  const int mc = (me - root + P) % P;  // my cyclic index
  int64_t cacc = 0, sacc = 0;
  for (int p = 0; p < P; ++p) {
    // contiguous: my cyclic blocks Lb (global I = Lb*P + mc) with I in rank p's blocked range [p*K, (p+1)*K)
    int64_t lo = ceil_div_i64(p * K - mc, P), hi = ceil_div_i64((p + 1) * K - mc, P);
    if (lo < 0) lo = 0;
    if (hi > K) hi = K;
    if (hi < lo) hi = lo;
    pl->lo[p] = (int)lo;
    pl->ccnt[p] = (int)(hi - lo);
    pl->coff[p] = cacc;
    cacc += hi - lo;
    // strided: blocks of my blocked range [me*K, (me+1)*K) whose cyclic owner is rank p, i.e. I == cyc(p) (mod P)
    const int pc = (p - root + P) % P;
    const int64_t base = (int64_t)me * K;
    const int64_t i0 = base + (((pc - base) % P) + P) % P;
    const int64_t cnt = i0 < base + K ? ceil_div_i64(base + K - i0, P) : 0;
    pl->first[p] = (int)(i0 - base);
    pl->scnt[p] = (int)cnt;
    pl->soff[p] = sacc;
    sacc += cnt;
  }
  return cacc == K && sacc == K;
}

// Strided side: which peer owns (cyclically) my blocked-local block `blk`, and which of that peer's blocks it is.
// 32-bit arithmetic on purpose (blk < 2^30): this runs once per element in the row-permuting kernel.
CANDMC_HD void strided_owner(const AxisPlan& pl, unsigned blk, int* p, unsigned* t) {
  const unsigned P = (unsigned)pl.P;
  const int owner = (int)((pl.base_mod + blk) % P);
  *p = owner;
  *t = (blk - (unsigned)pl.first[owner]) / P;
}

// Strided side, element level.  `blk` = local blocked block index on the permuted axis, `w` = offset inside the block,
// `o` = index along the other axis, `other` = local extent of the other axis.  Returns the element's position in the
// segmented exchange buffer (segment p = everything exchanged with rank p, in rank order):
//   rows axis (permuting block ROWS of a column-major matrix): segment p is (scnt[p]*nb) x other, column-major;
//   cols axis (permuting block COLUMNS): segment p is other x (scnt[p]*nb), column-major.
CANDMC_HD int64_t strided_segment_index(const AxisPlan& pl, bool rows_axis, int64_t blk, int w, int64_t o, int64_t other) {
  int p;
  unsigned t;
  strided_owner(pl, (unsigned)blk, &p, &t);
  const int64_t seg = pl.soff[p] * pl.nb * other;
  if (rows_axis) return seg + o * ((int64_t)pl.scnt[p] * pl.nb) + (int64_t)t * pl.nb + w;
  return seg + ((int64_t)t * pl.nb + w) * other + o;
}

}  // namespace candmc
