/* candmc_b200 — C ABI of the B200-native CANMM hot path (libcandmc_b200.so).
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, no CUDA / torch / NCCL types.  Each entry
 * point names the reference interface (solomonik/CANDMC, path:line) it replaces.  The C++ headers next to this
 * file (CANDMC.h and candmc/...) re-export the reference's own C++ signatures on top of these calls; the Python
 * package `candmc_b200` binds the same symbols with ctypes.  See INTEGRATION.md.
 *
 * Conventions
 *  - All matrices are FP64, column-major, exactly as in the reference.
 *  - Matrix pointers may be DEVICE pointers (fast path; what bench.py's device-resident leg uses) or HOST
 *    pointers (what the reference's own tests pass); host operands are staged through device memory inside the
 *    call.  The kind is detected with cudaPointerGetAttributes.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls with device pointers are
 *    asynchronous on that stream; calls with any host operand return after the result is back in host memory.
 *  - Every function returns CANDMC_OK (0) or an error code; candmc_last_error() gives the text.  There is NO
 *    CPU fallback: without a compute-capability-10.x GPU every compute call returns CANDMC_ERR_NODEVICE.
 *    (The reference has no return codes — it asserts/ABORTs, alg/shared/util.h:127-138; the C++ shim keeps that.)
 */
#ifndef CANDMC_B200_H
#define CANDMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CANDMC_B200_VERSION 100 /* 0.1.0 */

enum {
  CANDMC_OK = 0,
  CANDMC_ERR_INVALID = 1,
  CANDMC_ERR_CUDA = 2,
  CANDMC_ERR_NCCL = 3,
  CANDMC_ERR_NOMEM = 4,
  CANDMC_ERR_NODEVICE = 5
};

/* ---- runtime ---------------------------------------------------------------------------------------------- */
int candmc_version(void);
const char* candmc_last_error(void);
/* Bind this process to one GPU (device < 0: the current CUDA device).  One process per GPU, like one MPI rank
 * per grid point in the reference (alg/shared/comm.h:125-136 INIT_COMM). */
int candmc_init(int device);
int candmc_finalize(void);
int candmc_device_sm_count(int* out);
/* Number of kernels this library has launched so far in this process (bench.py reports the delta). */
unsigned long long candmc_launch_count(void);
/* Depth sum of the replicated-grid multiplies: 1 (default) fuses the all-reduce over the depth communicator into the
 * epilogue of the last GEMM on 1 x 1 x c grids — partial tiles travel as P2P stores over NVLink into CUDA-IPC-mapped
 * windows of the other depth ranks while the GEMM is still running; 2 also on q x q x c grids (parity-green on 8 B200s, but
 * measured slower there than the NCCL all-reduce it replaces — 277.4 against 280.9 TFLOP/s on 2x2x2, DESIGN.md 4 — hence
 * opt-in); 0 uses ncclAllReduce after the GEMM (also the automatic fallback when IPC is
 * unavailable or the block is not a multiple of 128*c).  Must be the same on all ranks. */
int candmc_set_fused_reduce(int on);
/* Host operands on q x q x c grids: 1 (default) skips the upload of an A (B) block whose grid column (row) is not one of the
 * layer's panels — the multiply never reads it there; 0 uploads both blocks on every rank. */
int candmc_set_skip_unused_uploads(int on);
/* candmc_d25_summa on grids with c > 1 picks the protocol of its depth sum from the kind of mat_C it is given (a host block
 * leaves in column slabs, each summed by its own all-reduce; a device block takes one all-reduce or the fused epilogue), so
 * ALL RANKS OF A DEPTH GROUP MUST PASS THE SAME KIND OF POINTER FOR mat_C — as every caller of the reference does.  A mixed
 * group issues mismatched collectives and hangs the GPUs.  1 = verify it first (one 8-byte all-reduce over the depth
 * communicator and a stream synchronisation per call; a mismatch returns CANDMC_ERR_INVALID on every rank of the group);
 * 0 (default) = trust the caller.  Also the environment variable CANDMC_CHECK_PEER_ARGS. */
int candmc_set_check_peer_args(int on);
/* Host C blocks in candmc_d25_summa: 1 (default) = the last launch group of the multiply is cut into column slabs of b/2, b/4,
 * b/8, b/8 columns (each still over all of the group's k), each slab is summed over the depth and downloaded while the next
 * ones multiply; 0 = one download at the end. */
int candmc_set_early_c_download(int on);
/* Test hooks (pure host arithmetic, no device needed): the launch groups a sweep cuts a panel's k-chunks into — grp_hi[t] = end
 * of the group that starts at chunk t (mode = candmc_set_merge_panels' value) — and the column slabs a host C block is
 * finalised in (candmc_set_early_c_download; fin_slabs = 8 / 4 / 2 equal slabs is the fallback when b is not a multiple of 1024). */
int candmc_debug_launch_groups(int nchunks, int mode, int first_panel, int last_panel, int host_ops, int all_dma, int fused, int nn,
                               int* grp_hi);
int candmc_debug_fin_slab_widths(int64_t b, int fin_slabs, int64_t* widths, int cap, int* count);
/* Pinned (page-locked) host B blocks on grids: 1 (default) = every k-chunk is gathered straight out of host memory by the pack
 * kernel — coalesced reads over PCIe, chunk-major on arrival — so the first multiply starts after 1/8 of the block instead
 * of after all of it; 0 = one copy of the whole block, re-laid out on the device (also what pageable memory gets). */
int candmc_set_host_gather(int on);
/* SUMMA panel chunks (candmc_summa, candmc_d25_summa, the inner level of candmc_bcast_cannon_4d): 1 = the root writes them
 * into the consumers' CUDA-IPC-mapped windows with copy engines (cudaMemcpyAsync over NVLink + a 4-byte flag DMA, consumers
 * wait with cuStreamWaitValue32) — no SM, no NCCL kernel, the GEMM keeps all 148 SMs (default since round 2: 4 and 8 B200s,
 * DESIGN.md 4; the Cannon staggers and shifts of candmc_bcast_cannon_4d / candmc_spcannon travel the same way); 0 =
 * ncclBroadcast on the CTA-capped background communicators, grouped ncclSend/ncclRecv for the shifts.  Must be the same on all
 * ranks; falls back to NCCL by itself when peer windows or stream memory operations are unavailable. */
int candmc_set_panel_transport(int on);
/* Panel chunks this process has shipped that way so far (0 = the transport is off or fell back to NCCL). */
unsigned long long candmc_panel_transport_sends(void);
/* Merged last-panel launches so far (candmc_set_merge_last_panel): with B read chunk-major through one tensor map (1) or
 * as a plain matrix on the panel's root (0).  Diagnostics for tests and benches. */
unsigned long long candmc_merged_panel_launches(int chunk_major_b);
/* Host B blocks: 1 = the rows of the first k-chunk are uploaded ahead of the rest so the first multiply starts earlier
 * (default 0: one copy of the whole block, whose wide rows keep the 2-D DMA efficient).  Only matters for PAGEABLE host
 * memory or with candmc_set_host_gather(0): pinned blocks are gathered chunk by chunk anyway. */
int candmc_set_b_first_chunk_early(int on);
/* Test/measurement hook: 0 disables the split-K path the GEMM takes for small tile counts (default on). */
int candmc_debug_splitk(int on);
/* Test/measurement hook: CTA tile of the TMA + DMMA GEMM — 128 (128 x 128 tiles, one CTA per SM), 64 (128 x 64 tiles, two CTAs
 * per SM whose epilogues hide behind each other's main loops) or 0 = chosen per launch (default). */
int candmc_debug_gemm_tile(int tile_n);
/* Test/measurement hook: 0 routes candmc_transpose through the LDG/STG kernel instead of the TMA load / TMA store kernel
 * (which is also the automatic choice for operands that are not 16-byte aligned with even leading dimensions). */
int candmc_debug_transpose_tma(int on);
/* Measurement hook: 1 = GEMM launches with beta != 0 pull the C tile their epilogue will read into L2 under the tile's last
 * k-tiles (bulk L2 prefetches issued by the TMA producer lane); 0 = the epilogue's loads go to HBM. */
int candmc_debug_prefetch_c(int on);
/* Measurement hook: GEMM launches leave `sms` SMs free (what the SUMMA sweeps do while NCCL panel traffic is in flight). */
int candmc_debug_gemm_reserve_sms(int sms);
/* Test/measurement hook: 1 makes the GEMM walk its tiles round-robin instead of claiming them from an atomic counter. */
int candmc_debug_static_schedule(int on);
/* Measurement hook (bench.py's roofline leg): while enabled every TMA+DMMA GEMM launch is bracketed by CUDA events on
 * its own stream; candmc_profile_gemm_stats synchronises and returns the number of launches, the sum of their device
 * durations and of their algorithmic flops (2*m*n*k) since the last enable. */
int candmc_profile_enable(int on);
int candmc_profile_gemm_stats(int64_t* launches, double* total_ms, double* total_flops);
/* Start / end (ms, relative to the first profiled launch) of up to `cap` profiled GEMM launches; *n = how many. */
int candmc_profile_gemm_timeline(double* start_ms, double* end_ms, int64_t cap, int64_t* n);
/* Tuning: CTA cap of the second NCCL communicator each grid axis uses for panel traffic that runs under a GEMM
 * (default 2; 0 = use the full-width communicator).  Must be set before the first multiply on a communicator. */
int candmc_set_background_ctas(int max_ctas);
/* Test hook: 1 routes candmc_dgemm through the generic CUDA-core kernel instead of the TMA+DMMA kernel. */
int candmc_debug_force_generic_gemm(int on);

/* ---- local kernels ---------------------------------------------------------------------------------------- */
/* C = alpha*op(A)*op(B) + beta*C.  Replaces cdgemm (alg/shared/lapack.h:10-16, lapack.cxx:425-434) and the
 * direct dgemm_ wrapper of split-dim Cannon (alg/MM/splitdim_cannon/spcannon_internal.h:51-63). */
int candmc_dgemm(char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream);
/* The same multiply with op(B) = B (k x n) given in the SUMMA pipeline's CHUNK-MAJOR layout: k / kc consecutive chunks, chunk t
 * holding rows [t*kc, (t+1)*kc) of B as a kc x n column-major matrix with leading dimension kc — what a rank holds after the
 * chunk-wise panel broadcasts that replace the MPI_Bcast of summa.cxx:66-89 / d25_summa.cxx:124-136.  One launch over all of k
 * (candmc_set_merge_last_panel uses it).  Device pointers only; k a multiple of kc, kc a multiple of 16, A and B 16-byte
 * aligned, lda even, alpha != 0; anything else is an error (the callers inside the library fall back to one launch per chunk). */
int candmc_dgemm_chunked_b(char transa, int64_t m, int64_t n, int64_t k, int64_t kc, double alpha, const double* A, int64_t lda,
                           const double* B, double beta, double* C, int64_t ldc, void* stream);
/* Single-precision companion (BASELINE north star: "optional FP32"; the reference has no single-precision multiply, so
 * this mirrors cdgemm's argument list with float data): C = alpha*op(A)*op(B) + beta*C on the 5th-generation tensor cores
 * (tcgen05.mma kind::tf32, accumulator in tensor memory).  Device pointers only; asynchronous on `stream`.  Operands that
 * are not K-major and TMA-readable in place are packed into the library workspace first, so calls on different streams must
 * not overlap.  candmc_set_f32_mode: TF32 products per FP32 product — 3 (default): operands split into TF32 high and low
 * parts, relative error per product <= 3 * 2^-20 before the FP32 accumulation (which adds what every single-precision GEMM
 * has: of the order of sqrt(k) * 2^-24 of the summed terms); 1: plain TF32 inputs (2^-10 per product). */
int candmc_sgemm(char transa, char transb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* stream);
int candmc_set_f32_mode(int tf32_products);
/* B(nrow x ncol, lda_B) = A(nrow x ncol, lda_A).  Replaces lda_cpy (alg/shared/util.h:459-471). */
int candmc_lda_cpy(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                   void* stream);
/* B = b*B + a*A.  Replaces the scaled lda_cpy overload (alg/shared/util.h:484-501). */
int candmc_lda_cpy_scaled(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, double* B,
                          double a, double b, void* stream);
/* B(cols x rows, ldb) = A(rows x cols, lda)^T, out of place.  Replaces TRANSPOSE/naive_transp
 * (alg/MM/splitdim_cannon/spcannon_internal.h:66-72; that one copies back in place). */
int candmc_transpose(int64_t rows, int64_t cols, const double* A, int64_t lda, double* B, int64_t ldb,
                     void* stream);
/* Fill X(nrow x ncol, ld) with the reference unit test's per-element generator
 * (test/MM/topo_pdgemm_unit.cxx:250-256,301-307): element at global (row0+r, col0+c) is the `which`-th (0=A, 1=B)
 * drand48() draw after srand48((col0+c)*n_global + (row0+r)).  X must be a device pointer. */
int candmc_fill_drand48(double* X, int64_t nrow, int64_t ncol, int64_t ld, int64_t row0, int64_t col0,
                        int64_t n_global, int which, void* stream);
/* out[0] = ||X - Y||_F^2, out[1] = ||Y||_F^2 (host doubles; synchronises).  X, Y device pointers. */
int candmc_frob_diff(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t nrow, int64_t ncol,
                     double* out_host2, void* stream);

/* ---- processor-grid communicators ------------------------------------------------------------------------- */
/* Opaque handle standing in for the reference's CommData_t.cm / MPI_Comm (alg/shared/comm.h:32-37). */
typedef struct candmc_comm candmc_comm_t;

#define CANDMC_UNIQUE_ID_BYTES 128
/* Rank 0 calls this and ships the 128 bytes to every rank by any out-of-band channel (bench.py uses
 * torch.distributed; the native launcher uses a file) — the role MPI_Init plays in INIT_COMM (comm.h:125-136). */
int candmc_get_unique_id(void* out_128_bytes);
/* Collective over all ranks: builds the world communicator (an NCCL communicator on this process's GPU). */
int candmc_comm_init_rank(const void* unique_id_128_bytes, int nranks, int rank, candmc_comm_t** out);
/* Collective over `parent`: MPI_Comm_split semantics (same color -> same child, ordered by key).  Replaces the
 * MPI_Comm_split calls inside SETUP_SUB_COMM / RSETUP_KDIR_COMM / RSETUP_LAYER_COMM (comm.h:145-195). */
int candmc_comm_split(candmc_comm_t* parent, int color, int key, candmc_comm_t** out);
int candmc_comm_free(candmc_comm_t* comm); /* FREE_CDT, comm.h:199-201 */
int candmc_comm_rank(const candmc_comm_t* comm, int* rank);
int candmc_comm_size(const candmc_comm_t* comm, int* size);
int candmc_comm_barrier(candmc_comm_t* comm); /* device-side barrier + host sync */
/* Collectives on FP64 buffers (device or host pointers) — the MPI calls of the hot path:
 * MPI_Bcast (summa.cxx:63-84 via POST_BCAST comm.h:110-112), MPI_Allreduce(SUM) (d25_summa.cxx:149,221; qr_2d.cxx:265). */
int candmc_comm_bcast(candmc_comm_t* comm, double* buf, int64_t count, int root, void* stream);
int candmc_comm_allreduce_sum(candmc_comm_t* comm, const double* sendbuf, double* recvbuf, int64_t count,
                              void* stream);

/* ---- distributed multiplies (alg/MM) ---------------------------------------------------------------------- */
/* Same fields and meaning as ctb_args_t (alg/MM/topo_pdgemm/topo_pdgemm_algs.h:6-15). */
typedef struct candmc_ctb_args {
  char trans_A;        /* 'N' / 'T': as in the reference the flags reach the LOCAL multiply only (summa.cxx:97, d25_summa.cxx:185, */
  char trans_B;        /* dual_cannon.cxx:163-166) — blocks travel as stored, each block product is op(A blk)*op(B blk); pinned  */
                       /* to the unmodified reference's outputs (golden *_TN / *_NT / *_TT).  On a 1 x 1 x c grid: op(A)*op(B).   */
  int64_t n;           /* global matrix dimension */
  int64_t lda_A;
  int64_t lda_B;
  int64_t lda_C;
  int64_t buffer_size; /* BYTES available in `buffer`; checked like the reference's ASSERTs */
  int ovp;
} candmc_ctb_args_t;

/* 2D SUMMA, C = A*B on a q x q grid.  Replaces summa (topo_pdgemm_algs.h:17-23, summa.cxx:26-101).
 * cdt_row: communicator along my grid row (rank = my column); cdt_col: along my grid column (rank = my row).
 * `buffer` may be NULL (an internal device workspace is used); if non-NULL it must hold buffer_size >= 4*b*b*8 bytes. */
int candmc_summa(const candmc_ctb_args_t* args, const double* mat_A, const double* mat_B, double* mat_C,
                 double* buffer, candmc_comm_t* cdt_row, candmc_comm_t* cdt_col, void* stream);
/* 2.5D SUMMA on q x q x c: layer l multiplies k-panels [l*q/c, (l+1)*q/c), then the depth sum leaves the full C
 * block on every layer.  Replaces d25_summa / d25_summa_ovp (topo_pdgemm_algs.h:25-49, d25_summa.cxx:33-281);
 * `ovp` picks the reference's buffer-size rule (3 vs 5 b^2 doubles) — on this implementation communication is
 * always overlapped.  Unlike the reference (SURVEY App. A-1,A-3) the first panel of every layer uses beta = 0 and
 * mat_A / mat_B are preserved.  Extension: cdt_row/cdt_col of size 1 with cdt_kdir of size c splits k across
 * the c ranks (the 1 x 1 x c grid the reference's q % c == 0 assert forbids). */
int candmc_d25_summa(const candmc_ctb_args_t* args, const double* mat_A, const double* mat_B, double* mat_C,
                     double* buffer, candmc_comm_t* cdt_row, candmc_comm_t* cdt_col, candmc_comm_t* cdt_kdir,
                     int ovp, void* stream);
/* SUMMA over (x1,y1) nested in Cannon over (x2,y2).  Replaces bcast_cannon_4d (topo_pdgemm_algs.h:51-59,
 * dual_cannon.cxx:40-215) with its *intended* semantics (the reference deadlocks for x2_np > 1). */
int candmc_bcast_cannon_4d(const candmc_ctb_args_t* args, const double* mat_A, const double* mat_B, double* mat_C,
                           double* buffer, candmc_comm_t* cdt_x1, candmc_comm_t* cdt_y1, candmc_comm_t* cdt_x2,
                           candmc_comm_t* cdt_y2, void* stream);
/* Split-dimensional Cannon, C <- alpha*A*B + beta*C with rectangular local blocks (A m x k, B k x n or n x k
 * per transp_B, C m x n).  Replaces kput_cannon (bidir != 0) / kuni_cannon (bidir == 0)
 * (alg/MM/splitdim_cannon/spcannon.h:31-59, spcannon.cxx:237-347).  `world` must contain kary^ndim ranks laid out
 * as in spcannon.cxx:59-62; any even ndim (ndim = 4 needs 16 ranks and has run on the simulator only — an
 * NVSwitch crossbar has no torus dimensions to split over).  A and B are preserved (the reference destroys them). */
int candmc_spcannon(int bidir, int rank, int kary, int ndim, candmc_comm_t* world, int n, int m, int k,
                    char transp_A, double alpha, const double* A, char transp_B, double beta, const double* B,
                    double* C, void* stream);
/* CAQR trailing update A <- A - Y * (T^-1 * (Y^T A)) on one grid column.  Replaces upd_A
 * (alg/QR/qr_2d/qr_2d.cxx:224-282): cdgemm('T','N') :259, MPI_Allreduce over ccol :265, cdtrsm('L','L','N','N') with the
 * b x b lower-triangular T (ld = b) :271, cdgemm('N','N', alpha=-1, beta=1) :275.  Y is mb x b, A is mb x kb (local
 * extents).  T != NULL is the W_is_T form (what the pipelined drivers pass, :447-620); T == NULL is the reference's W == NULL
 * form (:241-246): T^-1 = lower triangle of the grid column's sum of Y^T Y with the diagonal halved (compute_invT_from_Y
 * :22-60), formed on the device.  The third form (W = the panel QR's factor) needs the whole grid view: candmc_update_A.
 * Device pointers are used in place and the call is asynchronous; HOST pointers (what the reference's own QR drivers hold)
 * are staged for the call, which then returns with A written back — that is what integration/qr_2d_upd_A_gpu.cxx, the
 * upd_A a maintainer links in front of the reference's, passes.  ccol may be NULL (single process column). */
int candmc_upd_A(const double* Y, int64_t lda_Y, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                 const double* T, candmc_comm_t* ccol, void* stream);
/* Tuning: the triangular solve inside upd_A / update_A — 1 (default since round 2: 8.77 against 9.04 ms for BASELINE config 5
 * on 4 B200s) = one warp per right-hand side (shuffles inside 32-row blocks, two block barriers per T tile), 0 = a block
 * barrier per row of T. */
int candmc_set_trsm_variant(int variant);
/* Processor-grid view of the CAQR drivers: mirror of `pview` (alg/shared/comm.h:66-84; the diagonal communicator is not
 * used on this path).  crow: ranks of my grid row (rank = my column); ccol: ranks of my grid column (rank = my row). */
typedef struct candmc_pview {
  int rrow; /* current root row */
  int rcol; /* current root column */
  candmc_comm_t* crow;
  candmc_comm_t* ccol;
  candmc_comm_t* cworld;
} candmc_pview_t;
/* (I - Y T^-1 Y^T) A on a block-cyclic nprow x npcol grid.  Replaces update_A (alg/QR/qr_2d/qr_2d.cxx:124-177): local
 * extents by the block-cyclic formulas (:140-147), Y panel packed with zeroed upper triangle and unit diagonal on the root
 * row (:155-165), MPI_Bcast along the grid row (:168), then upd_A (:224-282); with W == NULL the triangular factor is formed
 * from Y (compute_invT_from_Y, :22-60), with W_is_T != 0 W is the b x b lower-triangular T, and with W != NULL, W_is_T == 0
 * (what QR_2D hands in, :325) W is the b x b upper-triangular factor of the panel QR, read on the root rank (rrow, rcol)
 * only: T = lower(-W^-T Y1) as comp_bcast_T_from_W forms it (:179-208; alg/QR/hh_recon/hh_recon.cxx:26-31), delivered to
 * every rank along the root's grid row and then down the grid columns (pv->cworld is not used).  The three forms must be
 * chosen alike on every rank.  aggreg_Y may be NULL.  Device pointers are used in place and the call is asynchronous; HOST
 * pointers — what QR_2D itself holds (:325) — are staged for the call (the panel on the root column, W where it is read, A and
 * aggreg_Y where the rank has rows) and A / aggreg_Y are written back before it returns. */
int candmc_update_A(const double* Y, int64_t lda_Y, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                    const double* W, const candmc_pview_t* pv, double* aggreg_Y, int64_t lda_aY, int W_is_T,
                    void* stream);
/* Yamamoto form of the CAQR trailing update, A <- A + Qm * (T * (Qm^T A)) with T = (Q1 - S)^-1 kept explicitly.
 * candmc_upd_Yamamoto_A replaces upd_Yamamoto_A (alg/QR/qr_2d/qr_y2d.cxx:123-169): cdgemm('T','N') :140, MPI_Allreduce over
 * ccol :146, cdgemm('N','N', alpha=-1) with the b x b T (ld = b) :156, cdgemm('N','N', alpha=-1, beta=1) :160.
 * candmc_update_Yamamoto_A replaces update_Yamamoto_A (:68-120) with agg == NULL: block-cyclic local extents (:81-88), the
 * root column's panel packed and MPI_Bcast along the grid row (:101-110), T MPI_Bcast along the grid row IN PLACE (:112,
 * so T is an output on the other columns), then the update.  Device pointers; Qm is read on the root column only.
 * Both also take HOST pointers (staged for the call; A — and for candmc_update_Yamamoto_A the broadcast T — written back on
 * return): candmc_upd_Yamamoto_A is what integration/qr_2d_upd_A_gpu.cxx passes for the reference's unmodified QR_Yamamoto
 * drivers.  The aggregator's arrays always live in device memory. */
int candmc_upd_Yamamoto_A(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t mb, int64_t kb, int64_t b,
                          const double* T, candmc_comm_t* ccol, void* stream);
int candmc_update_Yamamoto_A(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                             double* T, const candmc_pview_t* pv, void* stream);
/* The reference's `aggregator` (alg/QR/qr_2d/qr_y2d.h:4-46, qr_y2d.cxx:13-62) with its arrays in device memory: the panels of a
 * block column appended side by side in aQm (lda_aQm x lda_aT, panel s at column n, row `shift`) and their aggregated T in aT
 * (lda_aT x lda_aT), so that the whole block column is applied to the trailing matrix by ONE candmc_upd_Yamamoto_A(aQm,
 * lda_aQm, ..., b = n, aT) — as QR_Yamamoto_2D_2D does (:370).  create zero-fills (the constructor, :13-23), reset = :25-31,
 * shift_down = :34-36 (host-side field).  `scratch` is private to the library. */
typedef struct candmc_aggregator {
  int64_t lda_aQm, lda_aT, shift, n;
  double* aQm;
  double* aT;
  double* scratch;
} candmc_aggregator_t;
int candmc_aggregator_create(int64_t lda_aQm, int64_t lda_aT, candmc_aggregator_t* out);
int candmc_aggregator_reset(candmc_aggregator_t* agg);
int candmc_aggregator_shift_down(candmc_aggregator_t* agg, int64_t b);
int candmc_aggregator_free(candmc_aggregator_t* agg);
/* update_Yamamoto_A with agg != NULL (qr_y2d.cxx:68-120): the update above, then aggregator::append (:38-62) of the broadcast
 * panel and T — aT[n.., 0..n] = T ((Qm^T aQm[shift.., 0..n], all-reduced over the grid column) aT[0..n, 0..n]), aT[n.., n..] = T.
 * update == 0 skips the trailing update: the last panel of a block column is only broadcast and appended (QR_Yamamoto_2D
 * :266-271).  A rank without rows of the panel contributes zeros to the sum (the reference adds an uninitialised buffer
 * there, :54-56).  agg may be NULL (= candmc_update_Yamamoto_A). */
int candmc_update_Yamamoto_A_agg(const double* Qm, int64_t lda_Qm, double* A, int64_t lda_A, int64_t m, int64_t k, int64_t b,
                                 double* T, const candmc_pview_t* pv, candmc_aggregator_t* agg, int update, void* stream);
/* Tuning: the SUMMA pipeline cuts each b-wide panel into up to 8 k-chunks of at least this many columns
 * (default 1024) so the broadcast of chunk t+1 runs under the GEMM of chunk t.  Tests lower it to exercise the
 * chunked path on small matrices. */
int candmc_set_min_kchunk(int64_t min_kchunk);
/* Earlier experiment, kept for A/B runs (= candmc_set_merge_panels(1); also CANDMC_MERGE_LAST_PANEL=1): in summa / d25_summa sweeps with at least two panels the LAST
 * panel's k-chunks — all broadcast under the previous panel's multiplies — are multiplied in one launch (plus one for the chunk
 * whose buffer slot came free last) instead of one launch per chunk; the reference has no counterpart (its summa.cxx:59-99
 * multiplies a panel with one blocking dgemm_ after blocking broadcasts). */
int candmc_set_merge_last_panel(int on);
/* How the k-chunks of a panel are grouped into launches (also CANDMC_MERGE_PANELS=0..3).  2 (default since round 2: 16 -> 3
 * launches per step on 2x2x1, 135 -> 143 TFLOP/s on 4 B200s) = as few launches as the arrival of the data allows: the first
 * panel of a sweep as chunk 0 alone — the only chunk whose transfer nothing hides — followed by ONE launch over chunks
 * 1 .. nc-1, which arrive while chunk 0 is multiplied; every later panel in ONE launch when it travelled by copy engines
 * (chunk 0 + the rest on the NCCL path, whose buffer slots are handed back launch by launch); chunk 0, chunks 1-2, the rest
 * while operands are still coming up from host memory; with the fused depth sum the last chunk keeps its own launch.
 * 0 = one launch per chunk (round 1's schedule); 1 = candmc_set_merge_last_panel(1); 3 = groups that double (chunk 0, chunk 1,
 * chunks 2-3, chunks 4-7), for links only a few times faster than the multiply. */
int candmc_set_merge_panels(int mode);
/* Tuning: on a 1x1x1 grid with HOST operands and n >= this (default 2048) the multiply is streamed through PCIe in
 * column panels (upload of panel j+1 and download of panel j-1 under the GEMM of panel j) instead of staged whole. */
int candmc_set_host_pipeline_min(int64_t min_n);
/* ... and how it is cut: n > 0 = n equal column panels (and n equal k-chunks of the first panel; 8 is what B200s have
 * measured), -1 = graduated (first panel n/4 wide with k-chunks growing by a tenth from k/16, then n/8 panels, last panels
 * shrinking to n/32: 1/16 of A is uploaded before the first multiply, each further chunk under the previous chunk's multiply,
 * 1/32 of C downloaded after the last), 0 = automatic (graduated when n, k >= 8192, else 8 equal panels). */
int candmc_set_host_pipeline_panels(int panels);
/* The cut itself for an n-column, k-deep product (host arithmetic only; exposed for the CPU-side tests): panel widths and
 * the first panel's k-chunks, at most `cap` of each. */
int candmc_host_pipeline_cut(int64_t n, int64_t k, int panels, int64_t* widths, int64_t* kchunks, int cap, int* npanels,
                             int* nchunks);

/* ---- symmetric full -> band reduction, trailing update ----------------------------------------
 * One level of sym_full2band (alg/SE/full_to_band.cxx:28-250) after its panel QR (:96): given the aggregated Householder
 * panel Y the QR left on every rank (mb x b, replicated along grid rows), forms invT (compute_invT_from_Y,
 * alg/QR/qr_2d/qr_2d.cxx:22-60), W = Y^T A (:122) summed over the grid column (:126-130), on the diagonal ranks
 * Z = Y^T W^T (:157, all-reduced over `cdiag` :161), U = Y invT^-1 (:172), V' = W - Z U^T / 2 (:186), broadcasts V' down
 * the columns and U along the rows (:205-207), multiplies UV' (:213), swaps it with the transposed grid partner (:222-224)
 * and applies A -= UV' + VU' to the trailing block (:239-241).
 * A points at the level's working corner (the reference's `A` argument, NOT the trailing block), pv holds the roots at
 * that corner as on entry to the reference routine (it is not modified: the caller rotates rrow / rcol by b / b_sub for
 * the next level, :90,245), cdiag is the communicator of the diagonal ranks (ignored elsewhere; test/SE/test_full2band.cxx:
 * 161-173).  Square grid, rank = row + col * np.  Requires (b / b_sub) and ((n - b) / b_sub) to be multiples of the grid
 * dimension (outside that the reference itself corrupts its heap).  Device pointers; asynchronous on `stream`. */
int candmc_sym_full2band_update(double* A, int64_t lda_A, int64_t n, int64_t b, int64_t b_sub, const candmc_pview_t* pv,
                                candmc_comm_t* cdiag, const double* Y, int64_t lda_Y, void* stream);
/* The offsets and extents of that level on grid position (myrow, mycol) (full_to_band.cxx:57-79): where the trailing block
 * starts inside the local array (rows, columns) and its local size.  Host arithmetic only. */
int candmc_sym_full2band_extents(int64_t n, int64_t b, int64_t b_sub, int np, int myrow, int mycol, int rrow, int rcol,
                                 int64_t* loc_row_offset, int64_t* loc_col_offset, int64_t* mb, int64_t* kb);

/* ---- block-cyclic <-> blocked redistribution ------------------------------------------------
 * Converts the local piece of an m x n matrix between the ScaLAPACK-style block-cyclic layout of the reference's QR / SE
 * drivers (block nb, global block row I on grid row (I + pv->rrow) mod nprow at local block row I div nprow; columns
 * likewise with rcol — test/QR/test_qr_2d.cxx:87-94, alg/QR/qr_2d/qr_2d.cxx:140-147, alg/SE/dmatrix.cxx:194-203) and the
 * blocked layout of the CANMM multiplies (grid row i owns rows [i*m/nprow, (i+1)*m/nprow), test/MM/topo_pdgemm_unit.cxx:
 * 250-256).  to_cyclic = 0: src is block-cyclic, dst blocked; 1: the reverse.  Both local pieces are (m/nprow) x
 * (n/npcol), column-major, device pointers, src != dst.  Requires m % (nb*nprow) == 0 and n % (nb*npcol) == 0.
 * pv->crow / pv->ccol are the row / column communicators (rank = my column / my row); pv->cworld is not used.
 * Collective over both communicators; asynchronous on `stream`. */
int candmc_redistribute(int to_cyclic, int64_t m, int64_t n, int64_t nb, const double* src, int64_t ld_src, double* dst,
                        int64_t ld_dst, const candmc_pview_t* pv, void* stream);
/* The redistribution's index plan on one grid axis (pure host code; exposed for the CPU-side tests).  P ranks, K blocks
 * per rank, cyclic root `root`.  For every peer p: [lo[p], lo[p]+ccnt[p]) = my cyclic-local blocks that rank p owns in
 * the blocked layout; first[p] + t*P (t < scnt[p]) = my blocked-local blocks that rank p owns in the cyclic layout. */
int candmc_redist_axis_plan(int P, int me, int root, int64_t K, int nb, int* lo, int* ccnt, int* first, int* scnt);
/* Where the kernels put element (block blk, offset w, other-axis index o) of the blocked-side matrix inside the segmented
 * exchange buffer, and which peer's segment that is. */
int candmc_redist_strided_index(int P, int me, int root, int64_t K, int nb, int rows_axis, int64_t blk, int w, int64_t o,
                                int64_t other, int64_t* idx, int* peer);

/* ---- DMatrix pack / replication operations ----------------------------------------------------
 * candmc_dmat_t is the reference's DMatrix (alg/SE/dmatrix.h:7-33) without its ScaLAPACK descriptor: an nrow x ncol matrix
 * in block-cyclic layout (block b) on the grid of `pv`, whose current roots pv.rrow / pv.rcol own global block (0, 0);
 * `data` is the DEVICE pointer to the local piece, `lda` its leading dimension.  Local extents follow dmatrix.cxx:194-203.
 * Outputs are caller-allocated device buffers; sizes are given per call in doubles (mr x mc = the local extents). */
typedef struct candmc_dmat {
  int64_t nrow, ncol, b, lda;
  double* data;
  candmc_pview_t pv;
} candmc_dmat_t;
/* get_mynrow / get_myncol (dmatrix.cxx:194-203); host arithmetic only. */
int candmc_dmat_local_extents(const candmc_dmat_t* A, int64_t* mynrow, int64_t* myncol);
/* slice (dmatrix.cxx:367-394): the numrows x numcols sub-matrix at (firstrow, firstcol), by reference — rotated roots and a
 * moved data pointer, no copy; host arithmetic only. */
int candmc_dmat_slice(const candmc_dmat_t* A, int64_t firstrow, int64_t numrows, int64_t firstcol, int64_t numcols,
                      candmc_dmat_t* out);
/* get_contig (dmatrix.cxx:470-484): out (mr x mc, ld = mr) = the local piece. */
int candmc_dmat_get_contig(const candmc_dmat_t* A, double* out, void* stream);
/* replicate_vertical (dmatrix.cxx:268-289): rep (nrow * mc doubles) = the packed local pieces of the ranks of my grid
 * column, in rank order (MPI_Allgather over ccol).  Needs (nrow / b) % nprow == 0. */
int candmc_dmat_replicate_vertical(const candmc_dmat_t* A, double* rep, void* stream);
/* replicate_horizontal (dmatrix.cxx:294-304): rep (ncol * mr doubles), MPI_Allgather over crow. */
int candmc_dmat_replicate_horizontal(const candmc_dmat_t* A, double* rep, void* stream);
/* reduce_scatter_horizontal (dmatrix.cxx:310-355): cntrb holds ncol * mr doubles = npcol chunks of mr x mc; afterwards
 * A.data += the sum over my grid row of everyone's chunk number <my column>.  cntrb is scratch on return.  (Any npcol, any
 * lda; the reference needs a power of two and lda == mr.  Summation order differs from the reference's butterfly.) */
int candmc_dmat_reduce_scatter_horizontal(const candmc_dmat_t* A, double* cntrb, void* stream);
/* transpose_data (dmatrix.cxx:252-263): out (mr x mc) = the packed local piece of my transposed grid partner, world rank
 * crow.rank + ccol.rank * npcol (square grids, world rank = myrow + mycol * nprow as in test/QR/test_qr_2d.cxx:367-374). */
int candmc_dmat_transpose_data(const candmc_dmat_t* A, double* out, void* stream);
/* foldcols (dmatrix.cxx:527-552): out = the (mr/factor) x (mc*factor) local piece (ld = mr/factor) of the nrow/factor x
 * ncol*factor matrix whose column group i holds local block rows j*factor + i.  Needs mr % (b*factor) == 0. */
int candmc_dmat_foldcols(const candmc_dmat_t* A, int64_t factor, double* out, void* stream);
/* foldrows (dmatrix.cxx:560-584): the inverse regrouping, out = (mr*factor) x (mc/factor), ld = mr*factor. */
int candmc_dmat_foldrows(const candmc_dmat_t* A, int64_t factor, double* out, void* stream);

/* The fold kernels' index map (pure host code, for the CPU-side tests): source offset in the mr x mc (lda) local piece of
 * output element (rr, cc) of foldcols (foldcols != 0) or foldrows. */
int candmc_debug_fold_src_index(int foldcols, int64_t mr, int64_t mc, int64_t b, int64_t f, int64_t lda, int64_t rr, int64_t cc,
                                int64_t* src);
/* Test hook: one launch of the redistribution's permute kernel for the plan of rank `me` of `P` (no communicator), so a
 * single GPU can stand in for every rank of an axis.  gather != 0: SEG <- X (blocked side -> segments), else X <- SEG. */
int candmc_debug_redist_permute(int P, int me, int root, int64_t K, int nb, int rows_axis, int gather, double* X,
                                int64_t ldx, double* SEG, int64_t rows, int64_t cols, void* stream);

/* ---- accelerator seam of the 2.5D LU -------------------------------------------------------
 * The reference keeps three matrices on an accelerator, enum OFF_MAT { OFF_A, OFF_L, OFF_U } (alg/LU/lu_offload.h:19),
 * and addresses sub-blocks by (matrix, element offset, leading dimension).  Here they live in HBM.  `mat` is 0, 1, 2
 * for OFF_A, OFF_L, OFF_U.  Offsets and sizes count doubles and are 64-bit (the reference's are int).  All calls come
 * from one host thread.  The observable results are those of executing the calls one after the other (the
 * reference's host fallback); internally GEMMs and transfers run on two streams and are ordered only where they touch
 * overlapping elements. */
#define CANDMC_OFF_A 0
#define CANDMC_OFF_L 1
#define CANDMC_OFF_U 2
/* Binds the offload runtime to GPU (rank mod #GPUs).  Replaces set_mic_rank (lu_offload.cxx:126-128).  Optional: without
 * it the device is LOCAL_RANK mod #GPUs (CANDMC_OFF_DEVICE overrides), else the current device. */
int candmc_off_set_device(int rank);
/* Allocates `size` doubles of HBM for `mat` (an existing allocation is released first).  Replaces alloc_A / alloc_L /
 * alloc_U (lu_offload.cxx:479-531).  host_init, if not NULL, is uploaded — what alloc_A(size, ptr) does on the
 * reference's accelerator build (:482-486); the contents are undefined otherwise. */
int candmc_off_alloc(int mat, int64_t size, const double* host_init);
/* Replaces free_offload_A / _L / _U (lu_offload.cxx:545-573).  Waits for queued work first. */
int candmc_off_free(int mat);
/* Capacity hint (doubles) for the pinned + device staging used by sparse row traffic.  Replaces alloc_transfer /
 * free_offload_transfer (lu_offload.cxx:533-543, :575-585); the staging also grows on demand. */
int candmc_off_alloc_transfer(int64_t size);
int candmc_off_free_transfer(void);
/* Device pointer and size of `mat`: what get_mat_handle returns ON the accelerator (lu_offload.cxx:159-175).  The pointer
 * can be handed to candmc_dgemm / candmc_d25_summa etc.; call candmc_off_sync before using it on another stream. */
int candmc_off_device_ptr(int mat, double** out, int64_t* size);
/* Allocated size of `mat` in doubles. */
int candmc_off_size(int mat, int64_t* size);
/* Host-side get_mat_handle (the reference's caller memcpy's the local matrix through it, lu_25d_pvt.cxx:1600-1602):
 * returns a pinned host mirror holding the current contents; whatever the host writes there is uploaded before the next
 * operation on `mat`.  Later device-side changes are NOT reflected in a pointer obtained earlier — call again. */
int candmc_off_host_mirror(int mat, double** out);
/* C <- alpha * op(A) * op(B) + beta * C between blocks of offloaded matrices; asynchronous.  Replaces offload_gemm_A
 * (lu_offload.cxx:216-251).  C must not overlap A or B. */
int candmc_off_gemm(char tA, char tB, int64_t m, int64_t n, int64_t k, double alpha, int64_t offset_A, int mat_A,
                    int64_t lda_A, int64_t offset_B, int mat_B, int64_t lda_B, double beta, int64_t offset_C,
                    int mat_C, int64_t lda_C);
/* Blocks until every GEMM issued so far has finished.  Replaces wait_gemm (lu_offload.cxx:130-134). */
int candmc_off_wait_gemm(void);
/* Host (or device) block A, nrow x ncol with leading dimension lda_A  ->  mat_B at offset_B with leading dimension
 * lda_B.  Replaces upload_lda_cpy (lu_offload.cxx:366-392).  A may be reused when the call returns. */
int candmc_off_upload(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, const double* A, int64_t offset_B,
                      int mat_B);
/* mat_A at offset_A, leading dimension lda_A  ->  host (or device) block B with leading dimension lda_B.  Replaces
 * download_lda_cpy (lu_offload.cxx:338-364).  B holds the data when the call returns. */
int candmc_off_download(int64_t nrow, int64_t ncol, int64_t lda_A, int64_t lda_B, int64_t offset_A, double* B,
                        int mat_A);
/* Row traffic for pivoting.  Row i lives at elements offsets[i] + j*lda_B (j < ncol) of mat_B and at A + i*lda_A
 * (ncol contiguous doubles) on the host.  rw = 'r': matrix -> host, 'w': host -> matrix, 's': swap.  Replaces
 * offload_sparse_rw (lu_offload.cxx:424-476). */
int candmc_off_sparse_rw(int64_t nrow, int64_t ncol, int64_t lda_B, double* A, int64_t lda_A, const int* offsets,
                         int mat_B, char rw);
/* Waits for every queued GEMM and transfer. */
int candmc_off_sync(void);
/* 0: run GEMMs and transfers on one stream (strictly serial); 1 (default): two streams + conflict scoreboard. */
int candmc_off_set_overlap(int enable);
/* Counters since load: {GEMMs, uploads, downloads, sparse calls, cross-stream waits inserted, GEMMs still tracked}. */
int candmc_off_stats(int64_t* out6);
/* The scoreboard's overlap test on two strided blocks of one matrix (exposed for the CPU-side tests). */
int candmc_off_blocks_overlap(int64_t off_x, int64_t ld_x, int64_t rows_x, int64_t cols_x, int64_t off_y,
                              int64_t ld_y, int64_t rows_y, int64_t cols_y);

#ifdef __cplusplus
}
#endif
#endif /* CANDMC_B200_H */
