"""Host-side mirror of the reference's CANMM operator interface, over the C ABI.

Names, argument order and meaning follow the reference headers (alg/MM/topo_pdgemm/topo_pdgemm_algs.h:6-59,
alg/MM/splitdim_cannon/spcannon.h:31-59, alg/shared/lapack.h:10-16, alg/shared/util.h:403-412); matrices are torch
CUDA tensors (float64, column-major storage: a b x b block is `torch.empty((cols, rows)).T` or any tensor whose
`data_ptr()` addresses column-major data), numpy arrays (host path, staged inside the library) or raw integer
addresses.  Errors raise CandmcError where the reference would assert/ABORT.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import _lib
from ._lib import check, lib


def _ptr(x):
    """Address of a matrix operand: torch tensor, numpy array, ctypes pointer or int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)!r}")


def _stream(stream):
    if stream is None:
        try:
            import torch

            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream
        except Exception:  # pragma: no cover - torch always present in this image
            pass
        return None
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def _ch(c):
    return c.encode("ascii") if isinstance(c, str) else c


@dataclass
class CommData_t:
    """Mirror of `CommData_t` (alg/shared/comm.h:32-63): communicator handle + cached np / rank."""
    cm: int = 0          # candmc_comm_t* (stands where the reference has an MPI_Comm)
    np: int = 1
    rank: int = 0
    color: int = 0
    alive: int = 1

    def free(self):       # FREE_CDT (comm.h:199-201)
        if self.cm:
            check(lib().candmc_comm_free(self.cm))
            self.cm = 0


@dataclass
class ctb_args_t:
    """Mirror of `ctb_args_t` (alg/MM/topo_pdgemm/topo_pdgemm_algs.h:6-15)."""
    n: int
    lda_A: int
    lda_B: int
    lda_C: int
    buffer_size: int = 0
    trans_A: str = "N"
    trans_B: str = "N"
    ovp: int = 0

    def _c(self):
        return _lib.CtbArgs(_ch(self.trans_A), _ch(self.trans_B), self.n, self.lda_A, self.lda_B, self.lda_C,
                            self.buffer_size, self.ovp)


# ---- local kernels ------------------------------------------------------------------------------------------------
def cdgemm(transa, transb, m, n, k, a, A, lda, B, ldb, b, Cm, ldc, stream=None):
    """cdgemm (alg/shared/lapack.h:10-16): C = a*op(A)*op(B) + b*C."""
    check(lib().candmc_dgemm(_ch(transa), _ch(transb), m, n, k, a, _ptr(A), lda, _ptr(B), ldb, b, _ptr(Cm), ldc,
                             _stream(stream)))


def cdgemm_chunked_b(transa, m, n, k, kc, a, A, lda, B, b, Cm, ldc, stream=None):
    """cdgemm with B (k x n, not transposed) in the SUMMA pipeline's chunk-major layout: chunk t = rows [t*kc, (t+1)*kc) stored
    as a kc x n matrix with leading dimension kc (include/candmc_b200.h: candmc_dgemm_chunked_b)."""
    check(lib().candmc_dgemm_chunked_b(_ch(transa), m, n, k, kc, a, _ptr(A), lda, _ptr(B), b, _ptr(Cm), ldc, _stream(stream)))


def csgemm(transa, transb, m, n, k, a, A, lda, B, ldb, b, Cm, ldc, stream=None):
    """Single-precision companion of cdgemm (float32 device operands, tcgen05 kind::tf32 with split operands; the
    reference has no counterpart): C = a*op(A)*op(B) + b*C."""
    check(lib().candmc_sgemm(_ch(transa), _ch(transb), m, n, k, a, _ptr(A), lda, _ptr(B), ldb, b, _ptr(Cm), ldc,
                             _stream(stream)))


def lda_cpy(nrow, ncol, lda_A, lda_B, A, B, a=None, b=None, stream=None):
    """lda_cpy and its scaled overload (alg/shared/util.h:459-501)."""
    if a is None:
        check(lib().candmc_lda_cpy(nrow, ncol, lda_A, lda_B, _ptr(A), _ptr(B), _stream(stream)))
    else:
        check(lib().candmc_lda_cpy_scaled(nrow, ncol, lda_A, lda_B, _ptr(A), _ptr(B), a, b, _stream(stream)))


def transpose(rows, cols, A, lda, B, ldb, stream=None):
    """Out-of-place TRANSPOSE (alg/MM/splitdim_cannon/spcannon_internal.h:66-72)."""
    check(lib().candmc_transpose(rows, cols, _ptr(A), lda, _ptr(B), ldb, _stream(stream)))


def fill_drand48(X, nrow, ncol, ld, row0, col0, n_global, which, stream=None):
    check(lib().candmc_fill_drand48(_ptr(X), nrow, ncol, ld, row0, col0, n_global, which, _stream(stream)))


def frob_diff(X, ldx, Y, ldy, nrow, ncol, stream=None):
    """(||X-Y||_F^2, ||Y||_F^2) computed on the device."""
    out = (C.c_double * 2)()
    check(lib().candmc_frob_diff(_ptr(X), ldx, _ptr(Y), ldy, nrow, ncol, out, _stream(stream)))
    return out[0], out[1]


def set_min_kchunk(v):
    check(lib().candmc_set_min_kchunk(v))


# ---- distributed multiplies -----------------------------------------------------------------------------------------
def summa(args: ctb_args_t, mat_A, mat_B, mat_C, buffer, cdt_row: CommData_t, cdt_col: CommData_t, stream=None):
    """summa (topo_pdgemm_algs.h:17-23)."""
    a = args._c()
    check(lib().candmc_summa(C.byref(a), _ptr(mat_A), _ptr(mat_B), _ptr(mat_C), _ptr(buffer), cdt_row.cm, cdt_col.cm,
                             _stream(stream)))


def _d25(args, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir, ovp, stream):
    a = args._c()
    check(lib().candmc_d25_summa(C.byref(a), _ptr(mat_A), _ptr(mat_B), _ptr(mat_C), _ptr(buffer), cdt_row.cm,
                                 cdt_col.cm, cdt_kdir.cm, ovp, _stream(stream)))


def d25_summa(args, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir, stream=None):
    """d25_summa (topo_pdgemm_algs.h:25-36)."""
    _d25(args, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir, 0, stream)


def d25_summa_ovp(args, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir, stream=None):
    """d25_summa_ovp (topo_pdgemm_algs.h:38-49)."""
    _d25(args, mat_A, mat_B, mat_C, buffer, cdt_row, cdt_col, cdt_kdir, 1, stream)


def bcast_cannon_4d(args, mat_A, mat_B, mat_C, buffer, cdt_x1, cdt_y1, cdt_x2, cdt_y2, stream=None):
    """bcast_cannon_4d (topo_pdgemm_algs.h:51-59)."""
    a = args._c()
    check(lib().candmc_bcast_cannon_4d(C.byref(a), _ptr(mat_A), _ptr(mat_B), _ptr(mat_C), _ptr(buffer), cdt_x1.cm,
                                       cdt_y1.cm, cdt_x2.cm, cdt_y2.cm, _stream(stream)))


def kput_cannon(rank, kary, ndim, comm: CommData_t, n, m, k, transp_A, alpha, A, transp_B, beta, B, Cm, stream=None):
    """kput_cannon (alg/MM/splitdim_cannon/spcannon.h:31-44), bidirectional."""
    check(lib().candmc_spcannon(1, rank, kary, ndim, comm.cm, n, m, k, _ch(transp_A), alpha, _ptr(A), _ch(transp_B),
                                beta, _ptr(B), _ptr(Cm), _stream(stream)))


def kuni_cannon(rank, kary, ndim, comm: CommData_t, n, m, k, transp_A, alpha, A, transp_B, beta, B, Cm, stream=None):
    """kuni_cannon (alg/MM/splitdim_cannon/spcannon.h:46-59), unidirectional."""
    check(lib().candmc_spcannon(0, rank, kary, ndim, comm.cm, n, m, k, _ch(transp_A), alpha, _ptr(A), _ch(transp_B),
                                beta, _ptr(B), _ptr(Cm), _stream(stream)))


def upd_A(Y, lda_Y, A, lda_A, mb, kb, b, T, ccol: CommData_t | None = None, stream=None):
    """The GEMM pair + allreduce + trsm of upd_A (alg/QR/qr_2d/qr_2d.cxx:224-282): T given = the W_is_T form, T None = the
    reference's W == NULL form (T^-1 from Y, compute_invT_from_Y :22-60).  numpy (host) operands are staged for the call."""
    check(lib().candmc_upd_A(_ptr(Y), lda_Y, _ptr(A), lda_A, mb, kb, b, _ptr(T), ccol.cm if ccol else None,
                             _stream(stream)))


@dataclass
class pview:
    """Mirror of `pview` (alg/shared/comm.h:66-84)."""
    rrow: int
    rcol: int
    crow: CommData_t
    ccol: CommData_t
    cworld: CommData_t | None = None


def update_A(Y, lda_Y, A, lda_A, m, k, b, W, pv: pview, aggreg_Y=None, lda_aY=0, W_is_T=False, stream=None):
    """update_A (alg/QR/qr_2d/qr_2d.h:74-85): (I - Y T^-1 Y^T) A on a block-cyclic grid; W None -> T from Y."""
    cpv = _lib.PView(pv.rrow, pv.rcol, pv.crow.cm, pv.ccol.cm, pv.cworld.cm if pv.cworld else None)
    check(lib().candmc_update_A(_ptr(Y), lda_Y, _ptr(A), lda_A, m, k, b, _ptr(W), C.byref(cpv), _ptr(aggreg_Y), lda_aY,
                                1 if W_is_T else 0, _stream(stream)))


def upd_Yamamoto_A(Qm, lda_Qm, A, lda_A, mb, kb, b, T, ccol: CommData_t | None = None, stream=None):
    """upd_Yamamoto_A (alg/QR/qr_2d/qr_y2d.cxx:123-169): A <- A + Qm (T (Qm^T A)) summed over one grid column."""
    check(lib().candmc_upd_Yamamoto_A(_ptr(Qm), lda_Qm, _ptr(A), lda_A, mb, kb, b, _ptr(T), ccol.cm if ccol else None,
                                      _stream(stream)))


class aggregator:
    """The reference's aggregator (alg/QR/qr_2d/qr_y2d.h:4-46, qr_y2d.cxx:13-62) with device arrays: the panels of a block column
    side by side in aQm (lda_aQm x lda_aT), their aggregated T in aT (lda_aT x lda_aT); n panels' columns so far, `shift` rows down."""

    def __init__(self, lda_aQm, lda_aT):
        self._c = _lib.Aggregator()
        check(lib().candmc_aggregator_create(lda_aQm, lda_aT, C.byref(self._c)))

    lda_aQm = property(lambda self: self._c.lda_aQm)
    lda_aT = property(lambda self: self._c.lda_aT)
    n = property(lambda self: self._c.n)
    shift = property(lambda self: self._c.shift)
    aQm = property(lambda self: self._c.aQm)
    aT = property(lambda self: self._c.aT)

    def reset(self):
        check(lib().candmc_aggregator_reset(C.byref(self._c)))

    def shift_down(self, b):
        check(lib().candmc_aggregator_shift_down(C.byref(self._c), b))

    def free(self):
        check(lib().candmc_aggregator_free(C.byref(self._c)))


def update_Yamamoto_A(Qm, lda_Qm, A, lda_A, m, k, b, T, pv: pview, agg: aggregator | None = None, stream=None, update=True):
    """update_Yamamoto_A (alg/QR/qr_2d/qr_y2d.h:59-68); T is broadcast along the grid row in place.  With an aggregator the
    broadcast panel and T are appended to it afterwards (aggregator::append, qr_y2d.cxx:38-62); update=False only broadcasts
    and appends — the last panel of a block column (QR_Yamamoto_2D, :266-271)."""
    cpv = _lib.PView(pv.rrow, pv.rcol, pv.crow.cm, pv.ccol.cm, pv.cworld.cm if pv.cworld else None)
    if agg is None:
        check(lib().candmc_update_Yamamoto_A(_ptr(Qm), lda_Qm, _ptr(A), lda_A, m, k, b, _ptr(T), C.byref(cpv), _stream(stream)))
    else:
        check(lib().candmc_update_Yamamoto_A_agg(_ptr(Qm), lda_Qm, _ptr(A), lda_A, m, k, b, _ptr(T), C.byref(cpv), C.byref(agg._c),
                                                 1 if update else 0, _stream(stream)))


def sym_full2band_extents(n, b, b_sub, np_, myrow, mycol, rrow, rcol):
    """(loc_row_offset, loc_col_offset, mb, kb) of one level of sym_full2band (alg/SE/full_to_band.cxx:57-79)."""
    out = [C.c_int64() for _ in range(4)]
    check(lib().candmc_sym_full2band_extents(n, b, b_sub, np_, myrow, mycol, rrow, rcol, *[C.byref(o) for o in out]))
    return tuple(o.value for o in out)


def sym_full2band_update(A, lda_A, n, b, b_sub, pv: pview, cdiag: CommData_t | None, Y, lda_Y, stream=None):
    """The trailing update of one level of sym_full2band (alg/SE/full_to_band.cxx:90-239) given the panel QR's aggregated Y:
    A (pointer at the level's working corner) -= U V' + V U' on the trailing block.  pv is NOT rotated (the caller does,
    :90,245); the panel QR itself (QR_2D_pipe, :96) is host-side TSQR and outside this library's scope."""
    cpv = _lib.PView(pv.rrow, pv.rcol, pv.crow.cm, pv.ccol.cm, pv.cworld.cm if pv.cworld else None)
    check(lib().candmc_sym_full2band_update(_ptr(A), lda_A, n, b, b_sub, C.byref(cpv), cdiag.cm if cdiag else None, _ptr(Y),
                                            lda_Y, _stream(stream)))


def cyclic_to_blocked(m, n, nb, A_cyc, lda_cyc, A_blk, lda_blk, pv: pview, stream=None):
    """Local piece of an m x n block-cyclic matrix (block nb, roots pv.rrow / pv.rcol: the layout of the reference's
    QR / SE drivers, test/QR/test_qr_2d.cxx:87-94) -> the blocked layout of the CANMM multiplies."""
    cpv = _lib.PView(pv.rrow, pv.rcol, pv.crow.cm, pv.ccol.cm, pv.cworld.cm if pv.cworld else None)
    check(lib().candmc_redistribute(0, m, n, nb, _ptr(A_cyc), lda_cyc, _ptr(A_blk), lda_blk, C.byref(cpv), _stream(stream)))


def blocked_to_cyclic(m, n, nb, A_blk, lda_blk, A_cyc, lda_cyc, pv: pview, stream=None):
    """Inverse of cyclic_to_blocked."""
    cpv = _lib.PView(pv.rrow, pv.rcol, pv.crow.cm, pv.ccol.cm, pv.cworld.cm if pv.cworld else None)
    check(lib().candmc_redistribute(1, m, n, nb, _ptr(A_blk), lda_blk, _ptr(A_cyc), lda_cyc, C.byref(cpv), _stream(stream)))
