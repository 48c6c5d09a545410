// candmc_b200 — FP64 GEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C   (column-major, Fortran dgemm
// semantics).  Replaces the vendor `dgemm_` behind the reference's `cdgemm` wrapper
// (reference alg/shared/lapack.cxx:425-434) at its hot call sites: summa.cxx:96-97, d25_summa.cxx:138-148,200-201,
// dual_cannon.cxx:164-193, spcannon.cxx:117,195 ('N','T'), qr_2d.cxx:259 ('T','N') and :275 ('N','N').
//
// Design (B200-first, not a translation of any BLAS):
//   * FP64 has no tcgen05/UMMA kind on sm_100a; the FP64 tensor atom is mma.sync.m8n8k4 (SASS DMMA.8x8x4) with
//     register accumulators.  One persistent CTA per SM owns 128x128 C tiles; 8 consumer warps each hold a 64x32
//     register tile (64 accumulator doubles/lane), one producer warp drives TMA and claims tiles from an atomic
//     counter (dynamic scheduling keeps the tail short and lets the kernel co-run with NCCL kernels).
//   * Operand tiles (128 x 16 doubles = 16 KiB each) are brought in by TMA (`cp.async.bulk.tensor.2d`, 128-byte
//     swizzle) into a 6-stage mbarrier ring (192 KiB smem).  Two smem layouts exist, chosen per operand by the
//     transpose flag:
//       K-major  (A with 'T', B with 'N'): one 128 B line per m/n index holding the tile's 16 k values; one TMA box.
//       MN-major (A with 'N', B with 'T'): 8 slabs of 16 m/n indices; in a slab one 128 B line per k; 8 TMA boxes.
//   * The DMMA slot -> (index, k) assignment is permuted so that every fragment `ld.shared.f64` is bank-conflict
//     free under the hardware swizzle for BOTH layouts (see frag maps below): index slots g=0..7 map to
//     {0,1,4,5,2,3,6,7}; the 4 DMMA k-steps of a 16-wide k tile use k = 8*(s>>1) + 2j + ((j&1)^(s&1)).
//     Sums over k and the set of output elements are unchanged — only which lane/step touches which element.
//   * Tiles are rasterised in groups of 8 tile-rows so a wave of 148 CTAs re-uses A/B panels out of the 126 MB L2.
//   * Small tile counts (< 3 waves) are cut along k as well (split-K): the partial tiles are parked in an L2-resident
//     scratch and the last unit of a tile to arrive sums them in a fixed order, so the result is deterministic.
//   * Out-of-range rows/cols/k are zero-filled by TMA (exact), stores are predicated: any m,n,k >= 0 works on the
//     TMA path as long as the base pointers are 16 B aligned and lda/ldb are even.  Otherwise a plain tiled
//     CUDA-core kernel (`gemm_f64_generic`) is used — still on the GPU; there is no CPU path.
#include <algorithm>

#include "common.cuh"
#include "ipc.h"
#include "runtime.h"

namespace candmc {

namespace {

constexpr int BM = 128;            // CTA tile rows
constexpr int BK = 16;             // k per stage (one 128 B swizzle span of doubles)
constexpr int OPER_A_BYTES = BM * BK * 8;        // 16 KiB of A per stage
// Two CTA shapes share one kernel body (TBN = tile columns):
//   128: one CTA per SM, 8 consumer warps (2 along M x 4 along N), 6 stages of 32 KiB — the least shared-memory traffic per
//        flop; the shape of long-k launches, where the epilogue is a fraction of a percent of a tile.
//    64: TWO CTAs per SM, 4 consumer warps each (2 x 2), 4 stages of 24 KiB.  While one CTA of an SM is in its epilogue
//        (a beta != 0 read-modify-write of C costs several k-tiles of a 128-wide tile's time during which the DMMA pipe of a
//        lone CTA idles) the other one is in its main loop and has the pipe to itself: the epilogues of short-k launches
//        (k-chunks of a SUMMA panel, the k = 512 update of the CAQR trailing matrix) disappear behind the neighbour's DMMAs.
//        The second CTA of every SM starts half a tile late so the two stay out of phase.
template <int TBN>
struct Shape {
  static constexpr int BN = TBN;
  static constexpr int NSTAGE = TBN == 128 ? 6 : 4;
  static constexpr int OPER_B_BYTES = TBN * BK * 8;
  static constexpr int STAGE_BYTES = OPER_A_BYTES + OPER_B_BYTES;
  static constexpr int NCW = 2 * (TBN / 32);                 // consumer warps, each a 64 x 32 register tile
  static constexpr int NTHREADS = (NCW + 4) * 32;            // + 1 producer warpgroup (setmaxnreg works per 4 warps)
  static constexpr int CTAS_PER_SM = TBN == 128 ? 1 : 2;
  static constexpr int PRODUCER_REGS = 40;                   // 128: 4 x 40 + 8 x 232 regs per lane = 64512 <= 65536 per SM
  static constexpr int CONSUMER_REGS = TBN == 128 ? 232 : 216;   // 64: 2 x (4 x 40 + 4 x 216) = 65536
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers + tile ids*/;
};
constexpr int RASTER_GROUP = 8;
constexpr int PF_WINDOW = 64;      // k-tiles before the end of a unit over which the C tile's L2 prefetch is spread
constexpr int kSplitSemCount = 4096;  // tiles a split-K launch may have (it is only used for small tile counts)

// index-slot permutation shared by A rows and B cols (see header comment)
__device__ __forceinline__ int cf_map(int g) { return ((g & 1) | ((g & 2) << 1) | ((g & 4) >> 1)); }
// k handled by lane slot j in DMMA step s of a 16-wide k tile
__device__ __forceinline__ int k_map(int s, int j) { return 8 * (s >> 1) + 2 * j + ((j & 1) ^ (s & 1)); }

// Byte offset of (index slot cf in an 8-aligned group, k) inside an operand tile; `par` = bit 3 of the index
// (only matters for the MN-major layout where 16 indices share a slab).
template <bool KMAJ>
__device__ __forceinline__ uint32_t lane_off(int cf, int k, int par) {
  if (KMAJ) {
    return cf * 128 + ((((k >> 1) ^ cf) & 7) << 4) + (k & 1) * 8;
  } else {
    return k * 128 + ((((par * 4 + (cf >> 1)) ^ (k & 7)) & 7) << 4) + (cf & 1) * 8;
  }
}
// Byte offset of fragment f (8 indices each) relative to the warp's first index (a multiple of 32).
template <bool KMAJ>
__device__ __forceinline__ constexpr uint32_t frag_off(int f) {
  return KMAJ ? f * 8 * 128 : (f >> 1) * 2048;
}
template <bool KMAJ>
__device__ __forceinline__ constexpr uint32_t warp_off(int first_index /*multiple of 32*/) {
  return KMAJ ? first_index * 128 : (first_index >> 4) * 2048;
}

struct TileCoord {
  int tm, tn;
};

__device__ __forceinline__ TileCoord tile_coord(int t, int tilesM, int tilesN) {
  // groups of RASTER_GROUP tile-rows; inside a group walk down the rows first, then across columns
  const int per_group = RASTER_GROUP * tilesN;
  const int grp = t / per_group;
  const int first_m = grp * RASTER_GROUP;
  const int rows = min(RASTER_GROUP, tilesM - first_m);
  const int r = t - grp * per_group;
  TileCoord c;
  c.tm = first_m + r % rows;
  c.tn = r / rows;
  return c;
}

// Tile of work unit `u`.  Plain: the rastered tile grid.  FUSED (GEMM + depth all-reduce): the tile columns are split
// into `c` contiguous owner ranges and every rank walks the ranges of the OTHER owners first ((me+1)%c, (me+2)%c, ...)
// and its own last, so the partial tiles it must send leave early and the partials it needs have long arrived.
struct UnitTile {
  int tm, tn, owner, within;
};
template <bool FUSED>
__device__ __forceinline__ UnitTile unit_tile(int tile, int tilesM, int tilesN, const FusedParams& fp) {
  UnitTile u;
  if (!FUSED) {
    const TileCoord tc = tile_coord(tile, tilesM, tilesN);
    u.tm = tc.tm; u.tn = tc.tn; u.owner = 0; u.within = tile;
  } else {
    const int per_owner = tilesM * fp.tiles_n_per_owner;
    const int phase = tile / per_owner;
    u.within = tile - phase * per_owner;
    u.owner = (fp.me + 1 + phase) % fp.c;
    const TileCoord tc = tile_coord(u.within, tilesM, fp.tiles_n_per_owner);
    u.tm = tc.tm;
    u.tn = tc.tn + u.owner * fp.tiles_n_per_owner;
  }
  return u;
}

template <bool A_KMAJ, bool B_KMAJ, bool FUSED, int TBN>
__global__ void __launch_bounds__(Shape<TBN>::NTHREADS, Shape<TBN>::CTAS_PER_SM)
gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    double* __restrict__ C, int64_t ldc, int M, int N, int K, double alpha, double beta,
                    int tilesM, int tilesN, int* __restrict__ tile_counter, int ksplit, double* __restrict__ part,
                    int* __restrict__ tile_sem, const __grid_constant__ FusedParams fp, int b_kc, int b_cs, int pf_c,
                    int* __restrict__ sm_slots, unsigned stagger_ns) {
  using S = Shape<TBN>;
  constexpr int BN = S::BN, NSTAGE = S::NSTAGE, STAGE_BYTES = S::STAGE_BYTES, OPER_BYTES = OPER_A_BYTES;
  constexpr int NCONSUMER_WARPS = S::NCW, NCT = S::NCW * 32;
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128 B swizzle atom; pointer arithmetic keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGE;
  // tile id published with the first k-stage of every tile (dynamic scheduler: tiles are claimed with an atomic
  // counter so CTAs that start late — e.g. behind an NCCL kernel holding their SM — simply take fewer tiles)
  volatile int* stage_tile = reinterpret_cast<volatile int*>(empty_bar + NSTAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KT = (K + BK - 1) / BK;
  // split-K: work unit u = tile * ksplit + s covers k-tiles [s*KTS, min(KT, (s+1)*KTS)); ksplit == 1 is the plain GEMM
  const int KTS = (KT + ksplit - 1) / ksplit;
  const int ntiles = tilesM * tilesN * ksplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NCONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= NCONSUMER_WARPS) {
    // ===================== TMA producer warpgroup (one lane works; the rest only donate registers) =========
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(S::PRODUCER_REGS));
    if (warp == NCONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (sm_slots != nullptr) {
        // two CTAs per SM: the one that arrives second on its SM starts half a tile late, so that its epilogues fall into
        // the other one's main loops (and vice versa) instead of coinciding with them
        if (atomicAdd(sm_slots + sm_id(), 1) & 1) {
          const unsigned long long t0 = global_timer_ns();
          while (global_timer_ns() - t0 < stagger_ns) nanosleep_ns(2000);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int iter = 0;; ++iter) {
        // dynamic: claim the next tile; static (tile_counter == nullptr): round-robin over the grid
        const int t = tile_counter ? atomicAdd(tile_counter, 1) : static_cast<int>(blockIdx.x + iter * gridDim.x);
        if (t >= ntiles) {  // sentinel: an empty stage whose tile id is -1
          mbar_wait(&empty_bar[stage], phase ^ 1);
          stage_tile[stage] = -1;
          mbar_arrive(&full_bar[stage]);
          break;
        }
        const UnitTile tc = unit_tile<FUSED>(t / ksplit, tilesM, tilesN, fp);
        const int m0 = tc.tm * BM, n0 = tc.tn * BN;
        const int kt0 = (t % ksplit) * KTS, kt1 = min(KT, kt0 + KTS);
        // beta != 0: the tile of C the epilogue will read is pulled into L2 while the last k-tiles of the unit are being
        // loaded (one bulk prefetch per column, spread over up to PF_WINDOW k-tiles), so the epilogue's loads hit L2
        // instead of waiting on HBM with the DMMA pipe idle
        const int pf_w = pf_c ? min(kt1 - kt0, PF_WINDOW) : 0;
        const int pf_per = pf_w > 0 ? (BN + pf_w - 1) / pf_w : 0;
        const double* pf_src = FUSED ? fp.Cin : C;
        const int64_t pf_ld = FUSED ? fp.ldin : ldc;
        const uint32_t pf_bytes = static_cast<uint32_t>(min(BM, M - m0) & ~1) * 8u;
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kt == kt0) stage_tile[stage] = t;
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + OPER_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = kt * BK;
          if (A_KMAJ) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) tma_load_2d(sa + s * 2048, &tmA, &full_bar[stage], m0 + 16 * s, k0);
          }
          if (B_KMAJ) {
            // chunk-major B (b_kc > 0; what the SUMMA pipeline receives its panels as): k-chunk ch is its own b_kc x b_cs
            // matrix behind chunk ch-1, so in tmB the chunks stand side by side as one b_kc x (b_cs * chunks) matrix; this
            // launch multiplies N <= b_cs of every chunk's columns (tmB's base is the first of them)
            if (b_kc > 0) {
              const int ch = k0 / b_kc;
              tma_load_2d(sb, &tmB, &full_bar[stage], k0 - ch * b_kc, n0 + ch * b_cs);
            } else {
              tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
            }
          } else {
#pragma unroll
            for (int s = 0; s < BN / 16; ++s) tma_load_2d(sb + s * 2048, &tmB, &full_bar[stage], n0 + 16 * s, k0);
          }
          if (pf_w > 0 && kt1 - kt <= pf_w && pf_bytes > 0) {
            const int c0 = (pf_w - (kt1 - kt)) * pf_per;
            for (int c = c0; c < min(BN, c0 + pf_per); ++c)
              if (n0 + c < N) l2_prefetch_bulk(pf_src + m0 + static_cast<int64_t>(n0 + c) * pf_ld, pf_bytes);
          }
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ===================== DMMA consumers =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(S::CONSUMER_REGS));
  const int g = lane >> 2;  // index slot
  const int j = lane & 3;   // k slot
  const int cf = cf_map(g);
  const int wm = warp & 1;   // 2 warps along M (64 rows each)
  const int wn = warp >> 1;  // BN / 32 warps along N (32 cols each)

  // per-lane smem byte offsets for the 4 k-steps (and, for MN-major, the two slab halves)
  const uint32_t smem_base = smem_u32(smem);
  uint32_t aoff[4][2], boff[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int k = k_map(s, j);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      aoff[s][p] = smem_base + warp_off<A_KMAJ>(wm * 64) + lane_off<A_KMAJ>(cf, k, p);
      boff[s][p] = smem_base + OPER_BYTES + warp_off<B_KMAJ>(wn * 32) + lane_off<B_KMAJ>(cf, k, p);
    }
  }

  double acc[8][4][2];
  double fa[2][8], fb[2][4];

  // `st` = byte offset of the stage inside the ring
  auto load_frags = [&](int buf, uint32_t st, int s) {
#pragma unroll
    for (int f = 0; f < 8; ++f)
      fa[buf][f] = lds_f64(aoff[s][A_KMAJ ? 0 : (f & 1)] + st + frag_off<A_KMAJ>(f));
#pragma unroll
    for (int f = 0; f < 4; ++f)
      fb[buf][f] = lds_f64(boff[s][B_KMAJ ? 0 : (f & 1)] + st + frag_off<B_KMAJ>(f));
  };
  auto mma_step = [&](int buf) {
#pragma unroll
    for (int f = 0; f < 8; ++f)
#pragma unroll
      for (int h = 0; h < 4; ++h) dmma884(acc[f][h][0], acc[f][h][1], fa[buf][f], fb[buf][h]);
  };

  int stage = 0;
  uint32_t phase = 0;
  for (;;) {
    mbar_wait(&full_bar[stage], phase);
    const int t = stage_tile[stage];
    if (t < 0) break;
    const int tile = t / ksplit, ks = t - tile * ksplit;
    const UnitTile tc = unit_tile<FUSED>(tile, tilesM, tilesN, fp);
    const int nkt = min(KT, (ks + 1) * KTS) - ks * KTS;
#pragma unroll
    for (int f = 0; f < 8; ++f)
#pragma unroll
      for (int h = 0; h < 4; ++h) acc[f][h][0] = acc[f][h][1] = 0.0;
    load_frags(0, stage * STAGE_BYTES, 0);
    for (int kt = 0; kt < nkt; ++kt) {
      const uint32_t st = stage * STAGE_BYTES;
      load_frags(1, st, 1);
      mma_step(0);
      load_frags(0, st, 2);
      mma_step(1);
      load_frags(1, st, 3);
      mma_step(0);
      // all of this warp's reads of `stage` are now in registers (fa/fb[1] are consumed below; the loads were
      // issued above and mbarrier.arrive has release semantics)
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == NSTAGE) {
        nstage = 0;
        nphase ^= 1;
      }
      if (kt + 1 < nkt) {
        mbar_wait(&full_bar[nstage], nphase);
        load_frags(0, nstage * STAGE_BYTES, 0);
      }
      mma_step(1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      stage = nstage;
      phase = nphase;
    }

    if (ksplit > 1) {
      // ---- split-K: park the partial tile (thread-major, fully coalesced), the last of the `ksplit` units of this tile
      // to arrive sums all partials in the fixed order s = 0..ksplit-1 (deterministic) and runs the epilogue ----
      double* mine = part + (static_cast<int64_t>(tile) * ksplit + ks) * (BM * BN) + threadIdx.x;
#pragma unroll
      for (int f = 0; f < 8; ++f)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          __stcg(mine + (f * 8 + h * 2 + 0) * NCT, acc[f][h][0]);
          __stcg(mine + (f * 8 + h * 2 + 1) * NCT, acc[f][h][1]);
        }
      __threadfence();
      consumer_barrier<NCT>();
      __shared__ int s_last;
      if (threadIdx.x == 0) {
        const int old = atomicAdd(&tile_sem[tile], 1);
        s_last = (old == ksplit - 1);
        if (s_last) tile_sem[tile] = 0;  // ready for the next launch
      }
      consumer_barrier<NCT>();
      if (!s_last) continue;
      __threadfence();
      const double* all = part + static_cast<int64_t>(tile) * ksplit * (BM * BN) + threadIdx.x;
#pragma unroll
      for (int f = 0; f < 8; ++f)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          double s0 = 0.0, s1 = 0.0;
          for (int q2 = 0; q2 < ksplit; ++q2) {
            s0 += __ldcg(all + static_cast<int64_t>(q2) * (BM * BN) + (f * 8 + h * 2 + 0) * NCT);
            s1 += __ldcg(all + static_cast<int64_t>(q2) * (BM * BN) + (f * 8 + h * 2 + 1) * NCT);
          }
          acc[f][h][0] = s0;
          acc[f][h][1] = s1;
        }
    }
    // ---- epilogue: registers -> global, predicated, alpha/beta (beta==0 never reads C) ----
    // beta != 0: the 16 C values of one 8-column group are fetched together (L2 path, 16 loads in flight per lane — the
    // fragment registers are dead here) before they are combined; a load-use-store chain per element would leave the
    // DMMA pipe idle for tens of microseconds per tile.
    //
    // FUSED: the depth all-reduce happens here, tile by tile, over peer memory (NVLink P2P stores into the IPC-mapped
    // windows of the other depth ranks).  A tile owned by another rank: v = alpha*acc + beta*Cin is stored into the
    // owner's stage slot, then one release-store raises the tile's flag there.  A tile I own: wait for the flags of the
    // c-1 other sources (their partials were sent in THEIR first phases), add their partials, write the final value to
    // my C and into every peer's final slab, then bump the peers' delivery counters.
    const int row_base = tc.tm * BM + wm * 64 + cf;
    const int col_base = tc.tn * BN + wn * 32 + cf_map(2 * j);  // slots 2j, 2j+1 -> adjacent columns
    const bool owned = !FUSED || tc.owner == fp.me;
    const double* Cin = FUSED ? fp.Cin : C;
    const int64_t ldin = FUSED ? fp.ldin : ldc;
    const int owner_col0 = FUSED ? tc.owner * fp.tiles_n_per_owner * BN : 0;
    const int64_t slab = FUSED ? fp.ld * static_cast<int64_t>(fp.tiles_n_per_owner) * BN : 0;
    if (FUSED && owned) {
      if (threadIdx.x == 0) {
        const int per_owner = tilesM * fp.tiles_n_per_owner;
        for (int sidx = 0; sidx < fp.c - 1; ++sidx) {
          const uint32_t* fl = fp.sflag_local + static_cast<int64_t>(sidx) * per_owner + tc.within;
          uint32_t v;
          do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(fl) : "memory");
          } while (v != fp.epoch);
        }
      }
      consumer_barrier<NCT>();
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      double cv[2][8];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int f = 0; f < 8; ++f) cv[c][f] = 0.0;
      if (beta != 0.0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = col_base + h * 8 + c;
          const double* cp = Cin + static_cast<int64_t>(col) * ldin;
#pragma unroll
          for (int f = 0; f < 8; ++f) {
            const int row = row_base + f * 8;
            cv[c][f] = (col < N && row < M) ? beta * __ldcg(cp + row) : 0.0;
          }
        }
      }
      if (FUSED && owned) {
        for (int sidx = 0; sidx < fp.c - 1; ++sidx) {
          double sv[2][8];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col = col_base + h * 8 + c;
            const double* sp = fp.stage_local + sidx * slab + static_cast<int64_t>(col - owner_col0) * fp.ld;
#pragma unroll
            for (int f = 0; f < 8; ++f) sv[c][f] = __ldcg(sp + row_base + f * 8);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int f = 0; f < 8; ++f) cv[c][f] += sv[c][f];
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = col_base + h * 8 + c;
        if (col < N) {
#pragma unroll
          for (int f = 0; f < 8; ++f) {
            const int row = row_base + f * 8;
            if (row < M) {
              const double v = alpha * acc[f][h][c] + cv[c][f];
              if (!FUSED) {
                C[row + static_cast<int64_t>(col) * ldc] = v;
              } else if (!owned) {
                fp.stage[tc.owner][row + static_cast<int64_t>(col - owner_col0) * fp.ld] = v;
              } else {
                C[row + static_cast<int64_t>(col) * ldc] = v;
                for (int p = 0; p < fp.c; ++p)
                  if (p != fp.me) fp.cfinal[p][row + static_cast<int64_t>(col - owner_col0) * fp.ld] = v;
              }
            }
          }
        }
      }
    }
    if (FUSED) {
      __threadfence_system();
      consumer_barrier<NCT>();
      if (threadIdx.x == 0) {
        if (!owned) {
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(fp.sflag[tc.owner] + tc.within), "r"(fp.epoch) : "memory");
        } else {
          for (int p = 0; p < fp.c; ++p)
            if (p != fp.me) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(fp.done[p]) : "memory");
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Generic CUDA-core kernel: any alignment / leading dimension.  64x64 tile, 16x16 threads, 4x4 micro-tile.
// Only used when the TMA path's alignment preconditions do not hold (odd lda, 8-byte-aligned pointers).
// ---------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256)
gemm_f64_generic_kernel(int transA, int transB, int M, int N, int K, double alpha, const double* __restrict__ A,
                        int64_t lda, const double* __restrict__ B, int64_t ldb, double beta,
                        double* __restrict__ C, int64_t ldc) {
  __shared__ double sA[GK][GT + 1];
  __shared__ double sB[GK][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      int i, kk;
      if (!transA) { i = e % GT; kk = e / GT; } else { kk = e % GK; i = e / GK; }
      const int gm = m0 + i, gk = k0 + kk;
      double v = 0.0;
      if (gm < M && gk < K) v = transA ? A[gk + static_cast<int64_t>(gm) * lda] : A[gm + static_cast<int64_t>(gk) * lda];
      sA[kk][i] = v;
    }
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      int i, kk;
      if (!transB) { kk = e % GK; i = e / GK; } else { i = e % GT; kk = e / GT; }
      const int gn = n0 + i, gk = k0 + kk;
      double v = 0.0;
      if (gn < N && gk < K) v = transB ? B[gn + static_cast<int64_t>(gk) * ldb] : B[gk + static_cast<int64_t>(gn) * ldb];
      sB[kk][i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][tx + 16 * i];
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = sB[kk][ty + 16 * i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(a[i], b[jj], acc[i][jj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int col = n0 + ty + 16 * jj;
    if (col >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + tx + 16 * i;
      if (row >= M) continue;
      double* cp = C + row + static_cast<int64_t>(col) * ldc;
      double v = alpha * acc[i][jj];
      if (beta != 0.0) v += beta * *cp;
      *cp = v;
    }
  }
}

// C = beta * C (k == 0 or alpha == 0 degenerate case of dgemm)
__global__ void scale_c_kernel(int M, int N, double beta, double* __restrict__ C, int64_t ldc) {
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e % M, c = e / M;
    double* cp = C + r + c * ldc;
    *cp = (beta == 0.0) ? 0.0 : beta * *cp;
  }
}

bool is_trans(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }
bool is_notrans(char t) { return t == 'N' || t == 'n'; }

// split-K factor for a launch with 128 x 128 tiles (1 = none): when the tiles alone cannot fill a few waves, pick the split
// that wastes the least of the last wave
int pick_ksplit(int M, int N, int K) {
  const int64_t tiles = static_cast<int64_t>((M + BM - 1) / BM) * ((N + 127) / 128);
  const int sms = runtime().num_sms;
  const int KT = (K + BK - 1) / BK;
  int ksplit = 1;
  if (runtime().splitk && tiles < 3 * sms && tiles <= kSplitSemCount) {
    // cost model in k-tile units: waves x (k-tiles per unit + per-unit overhead); the overhead of a split unit (park the
    // partial tile, the last arriver re-reads `sp` of them) was measured at ~9 k-tiles + 1 per partial, an unsplit
    // tile's epilogue at ~3 (profiles/r01_gemm_probe_speed*.jsonl)
    auto cost = [&](int sp) {
      const int64_t u = tiles * sp;
      const double waves = static_cast<double>((u + sms - 1) / sms);
      const double per_unit = static_cast<double>((KT + sp - 1) / sp) + (sp == 1 ? 3.0 : 8.0 + sp);
      return waves * per_unit;
    };
    double best = cost(1);
    for (int sp = 2; sp <= 16; ++sp) {
      if (KT / sp < 8) break;
      if ((KT + sp - 1) / sp * (sp - 1) >= KT) continue;  // would leave an empty unit
      const double c = cost(sp);
      if (c < 0.97 * best) {
        best = c;
        ksplit = sp;
      }
    }
  }
  return ksplit;
}

// CTA tile width of a launch (see Shape).  Measured on a B200 (profiles/r02_gemm_probe_tile.jsonl, TFLOP/s 128-wide / 64-wide):
// 16384^2 x 2048 beta=1 on 146 SMs 33.6 / 35.1, beta=0 35.1 / 35.8; 8192^2 x 1024 beta=1 31.8 / 33.8; 65536 x 8192 x 512 beta=1
// 31.0 / 32.6; 16384^3 36.0 / 36.1; 8192^3 35.6 / 35.8 — but 4096^3 35.4 / 35.1 and 2048^3 30.6 / 29.5.  So: two CTAs per SM
// with 128 x 64 tiles whenever the 128-wide tiles fill at least eight waves or a short-k beta != 0 launch makes the epilogue
// a large part of a tile; 128 x 128 tiles for small products, split-K launches and the fused depth-sum epilogue.
int pick_tile_n(int M, int N, int K, double beta, bool fused, int ksplit) {
  if (fused || ksplit > 1) return 128;
  if (runtime().gemm_tile_n != 0) return runtime().gemm_tile_n;
  const int64_t tiles = static_cast<int64_t>((M + BM - 1) / BM) * ((N + 127) / 128);
  const int KT = (K + BK - 1) / BK;
  if (beta != 0.0 && KT <= 256) return 64;
  return tiles >= 8 * runtime().num_sms ? 64 : 128;
}

template <bool AK, bool BK_, bool FUSED, int TBN>
int launch_tma(const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, int64_t ldc, int M, int N, int K,
               double alpha, double beta, cudaStream_t stream, const FusedParams* fused, int b_kc, int b_cs, int ksplit) {
  using S = Shape<TBN>;
  constexpr int BN = S::BN;
  static bool configured = false;  // per template instantiation
  auto kern = gemm_f64_tma_kernel<AK, BK_, FUSED, TBN>;
  if (!configured) {
    CANDMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    configured = true;
  }
  const int tilesM = (M + BM - 1) / BM, tilesN = (N + BN - 1) / BN;
  const int64_t tiles = static_cast<int64_t>(tilesM) * tilesN;
  const int sms = runtime().num_sms;
  const int KT = (K + BK - 1) / BK;
  double* part = nullptr;
  int* sem = nullptr;
  if (ksplit > 1) CANDMC_TRY(splitk_buffers(tiles * ksplit * BM * BN, &part, &sem, stream));
  const int64_t ntiles = tiles * ksplit;
  // leave `gemm_reserve_sms` SMs free when a schedule wants NCCL kernels to run beside this persistent kernel
  const int avail = (sms - runtime().gemm_reserve_sms > 0 ? sms - runtime().gemm_reserve_sms : 1) * S::CTAS_PER_SM;
  const int grid = static_cast<int>(ntiles < avail ? ntiles : avail);
  int* counter = nullptr;
  if (!runtime().static_schedule) CANDMC_TRY(next_tile_counter(&counter, stream));
  // two CTAs per SM: per-SM arrival counters tell the second CTA of an SM to start half a tile late (a 128 x 64 tile's k-tile
  // takes about 2.08 us while two CTAs share the DMMA pipe: 148 SMs x 64 FMA/clk at 1.965 GHz)
  int* sm_slots = nullptr;
  unsigned stagger_ns = 0;
  if (S::CTAS_PER_SM == 2 && grid > sms) {
    CANDMC_TRY(next_sm_slots(&sm_slots, stream));
    stagger_ns = static_cast<unsigned>(std::min<int64_t>((KT + ksplit - 1) / ksplit, 8192) * 1040);
  }
  if (runtime().profile) CANDMC_TRY(profile_begin_launch(stream, 2.0 * M * (double)N * (double)K));
  FusedParams fp;
  if (FUSED) fp = *fused;
  // L2 prefetch of the C tile in front of a beta != 0 epilogue (bulk prefetches need 16-byte aligned column segments)
  const double* cin = FUSED ? fp.Cin : C;
  const int64_t ldin = FUSED ? fp.ldin : ldc;
  const int pf_c = (runtime().prefetch_c && beta != 0.0 && ksplit == 1 && cin != nullptr &&
                    reinterpret_cast<uintptr_t>(cin) % 16 == 0 && ldin % 2 == 0) ? 1 : 0;
  kern<<<grid, S::NTHREADS, S::SMEM_BYTES, stream>>>(tmA, tmB, C, ldc, M, N, K, alpha, beta, tilesM, tilesN, counter, ksplit,
                                                     part, sem, fp, b_kc, b_cs, pf_c, sm_slots, stagger_ns);
  CANDMC_CUDA(cudaGetLastError());
  if (ksplit > 1) CANDMC_TRY(splitk_release(stream));
  if (runtime().profile) CANDMC_TRY(profile_end_launch(stream));
  runtime().launches++;
  return OK;
}

template <bool FUSED, int TBN>
int launch_tma_layouts(bool AK, bool BKm, const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, int64_t ldc, int M, int N,
                       int K, double alpha, double beta, cudaStream_t stream, const FusedParams* fused, int b_kc, int b_cs, int ksplit) {
  if (AK && BKm) return launch_tma<true, true, FUSED, TBN>(tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, fused, b_kc, b_cs, ksplit);
  if (AK && !BKm) return launch_tma<true, false, FUSED, TBN>(tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, fused, b_kc, b_cs, ksplit);
  if (!AK && BKm) return launch_tma<false, true, FUSED, TBN>(tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, fused, b_kc, b_cs, ksplit);
  return launch_tma<false, false, FUSED, TBN>(tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, fused, b_kc, b_cs, ksplit);
}

}  // namespace

int gemm_f64(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
             int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
             cudaStream_t stream) {
  return gemm_f64_fused(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream, nullptr);
}

int gemm_f64_fused(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                   int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream,
                   const FusedParams* fused) {
  return gemm_f64_ex(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream, fused, 0, 0);
}

bool gemm_f64_bchunked_ok(const double* A, int64_t lda, const double* B, int64_t n, int64_t k, int64_t b_kc) {
  // (n = columns per chunk: the chunk stride of the tensor map, whatever column range of it a launch multiplies)
  return b_kc > 0 && k > 0 && k % b_kc == 0 && b_kc % BK == 0 && n * (k / b_kc) < (1LL << 31) && lda % 2 == 0 &&
         reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0 && !runtime().force_generic;
}

int gemm_f64_bchunked(char transa, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                      const double* B, int64_t b_kc, double beta, double* C, int64_t ldc, cudaStream_t stream) {
  CANDMC_CHECK(alpha != 0.0 && gemm_f64_bchunked_ok(A, lda, B, n, k, b_kc),
               "dgemm(chunk-major B): needs k a multiple of the chunk depth, the chunk depth a multiple of %d, aligned operands", BK);
  return gemm_f64_ex(transa, 'N', m, n, k, alpha, A, lda, B, b_kc, beta, C, ldc, stream, nullptr, b_kc, n);
}

int gemm_f64_bchunked_cols(char transa, int64_t m, int64_t ncols, int64_t k, double alpha, const double* A, int64_t lda,
                           const double* B, int64_t b_kc, int64_t chunk_cols, int64_t col0, double beta, double* C, int64_t ldc,
                           cudaStream_t stream) {
  CANDMC_CHECK(alpha != 0.0 && col0 >= 0 && ncols > 0 && col0 + ncols <= chunk_cols && (col0 * b_kc) % 2 == 0 &&
                   gemm_f64_bchunked_ok(A, lda, B, chunk_cols, k, b_kc),
               "dgemm(chunk-major B, column range): needs k a multiple of the chunk depth, the chunk depth a multiple of %d, aligned operands", BK);
  return gemm_f64_ex(transa, 'N', m, ncols, k, alpha, A, lda, B + col0 * b_kc, b_kc, beta, C, ldc, stream, nullptr, b_kc, chunk_cols);
}

int gemm_f64_ex(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t stream,
                const FusedParams* fused, int64_t b_kc, int64_t b_cs) {
  NvtxRange nvtx_range("DGEMM");   // spcannon.cxx:116,194; mcdgemm_compute at d25_summa.cxx:185
  CANDMC_TRY(runtime_require());
  CANDMC_CHECK(is_trans(transa) || is_notrans(transa), "dgemm: bad transa '%c'", transa);
  CANDMC_CHECK(is_trans(transb) || is_notrans(transb), "dgemm: bad transb '%c'", transb);
  const bool tA = is_trans(transa), tB = is_trans(transb);
  CANDMC_CHECK(m >= 0 && n >= 0 && k >= 0, "dgemm: negative dimension m=%lld n=%lld k=%lld", (long long)m,
               (long long)n, (long long)k);
  CANDMC_CHECK(m < (1LL << 31) && n < (1LL << 31) && k < (1LL << 31), "dgemm: dimension exceeds 2^31-1");
  const int64_t rowsA = tA ? k : m, rowsB = tB ? n : k;
  CANDMC_CHECK(lda >= (rowsA > 1 ? rowsA : 1), "dgemm: lda=%lld < %lld", (long long)lda, (long long)rowsA);
  CANDMC_CHECK(b_kc > 0 || ldb >= (rowsB > 1 ? rowsB : 1), "dgemm: ldb=%lld < %lld", (long long)ldb, (long long)rowsB);
  CANDMC_CHECK(ldc >= (m > 1 ? m : 1), "dgemm: ldc=%lld < %lld", (long long)ldc, (long long)m);
  if (m == 0 || n == 0) return OK;
  const int M = (int)m, N = (int)n, K = (int)k;

  CANDMC_CHECK(fused == nullptr || (k > 0 && alpha != 0.0), "fused GEMM+all-reduce needs k > 0");
  if (k == 0 || alpha == 0.0) {
    if (beta == 1.0) return OK;
    const int64_t total = m * n;
    int grid = (int)((total + 255) / 256);
    const int cap = runtime().num_sms * 8;
    if (grid > cap) grid = cap;
    scale_c_kernel<<<grid, 256, 0, stream>>>(M, N, beta, C, ldc);
    CANDMC_CUDA(cudaGetLastError());
    runtime().launches++;
    return OK;
  }

  const bool aligned = (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 16 == 0) &&
                       (lda % 2 == 0) && (ldb % 2 == 0) && !runtime().force_generic;
  CANDMC_CHECK(fused == nullptr || (aligned && m == n && m % (128 * fused->c) == 0),
               "fused GEMM+all-reduce needs aligned square blocks with whole tile columns per depth rank");
  if (aligned) {
    // tensor maps: dim0 is the contiguous (stored-row) dimension of the operand
    CUtensorMap tmA, tmB;
    const bool AK = tA;    // op(A)(m,kk) = A[kk + m*lda]  -> K contiguous
    const bool BKm = !tB;  // op(B)(kk,n) = B[kk + n*ldb]  -> K contiguous
    CANDMC_TRY(encode_tmap_f64(&tmA, A, AK ? k : m, AK ? m : k, lda, 16, AK ? BM : 16));
    const int ksplit = fused ? 1 : pick_ksplit(M, N, K);
    const int tbn = pick_tile_n(M, N, K, beta, fused != nullptr, ksplit);
    if (b_kc > 0) {   // chunk-major B: the k / b_kc chunks (b_kc x n, ld = b_kc) side by side
      // B points at column col0 of chunk 0: the map ends where the last chunk ends
      CANDMC_TRY(encode_tmap_f64(&tmB, B, b_kc, b_cs * (k / b_kc - 1) + n, b_kc, 16, tbn));
    } else {
      CANDMC_TRY(encode_tmap_f64(&tmB, B, BKm ? k : n, BKm ? n : k, ldb, 16, BKm ? tbn : 16));
    }
    const int kc = static_cast<int>(b_kc), cs = static_cast<int>(b_cs);
    if (fused) {
      CANDMC_TRY((launch_tma_layouts<true, 128>(AK, BKm, tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, fused, kc, cs, 1)));
      runtime().fused_launches++;   // the epoch this launch belongs to is now in flight (FusedEpochGuard, ipc.h)
      return OK;
    }
    if (tbn == 64) return launch_tma_layouts<false, 64>(AK, BKm, tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, nullptr, kc, cs, ksplit);
    return launch_tma_layouts<false, 128>(AK, BKm, tmA, tmB, C, ldc, M, N, K, alpha, beta, stream, nullptr, kc, cs, ksplit);
  }

  CANDMC_CHECK(b_kc == 0, "dgemm(chunk-major B): operands not TMA-able");
  dim3 grid((M + GT - 1) / GT, (N + GT - 1) / GT);
  CANDMC_CHECK(grid.y <= 65535, "dgemm(generic path): n too large for unaligned operands");
  gemm_f64_generic_kernel<<<grid, 256, 0, stream>>>(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  CANDMC_CUDA(cudaGetLastError());
  runtime().launches++;
  return OK;
}

}  // namespace candmc
