"""Host-side mirror of the pack / replication half of the reference's DMatrix (alg/SE/dmatrix.h:7-253): same method
names and meaning, data on the GPU (torch tensors), every method a call into the C ABI (candmc_dmat_*).  The ScaLAPACK
half of DMatrix (pdgemm / pdsyrk / pdtrsm wrappers, QR) is outside this library's scope."""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import lib
from .mm import _stream, check, pview


class DMatrix:
    def __init__(self, nrow: int, ncol: int, b: int, pv: pview, lda: int = 0, data=None, _keep=None):
        """DMatrix(nrow, ncol, b, pv[, lda, data]) — dmatrix.cxx:51-91.  Without `data` a device buffer of
        get_myncol() * lda doubles is allocated (lda defaults to get_mynrow())."""
        self.nrow, self.ncol, self.b, self.pv = int(nrow), int(ncol), int(b), pv
        self._c = _lib.DMat(self.nrow, self.ncol, self.b, 0, None,
                            _lib.PView(pv.rrow, pv.rcol, pv.crow.cm, pv.ccol.cm, pv.cworld.cm if pv.cworld else None))
        self.lda = int(lda) if lda else self.get_mynrow()
        self._c.lda = self.lda
        if data is None:
            import torch

            self.tensor = torch.zeros(max(self.get_myncol() * self.lda, 1), dtype=torch.float64, device="cuda")
            self._ptr = self.tensor.data_ptr()
        else:
            self.tensor = _keep if _keep is not None else data
            self._ptr = data if isinstance(data, int) else data.data_ptr()
        self._c.data = self._ptr

    # dmatrix.cxx:194-210
    def _extents(self):
        mr, mc = C.c_int64(), C.c_int64()
        check(lib().candmc_dmat_local_extents(C.byref(self._c), C.byref(mr), C.byref(mc)))
        return mr.value, mc.value

    def get_mynrow(self) -> int:
        return self._extents()[0]

    def get_myncol(self) -> int:
        return self._extents()[1]

    def get_mysize(self) -> int:
        mr, mc = self._extents()
        return mr * mc

    def slice(self, firstrow, numrows, firstcol, numcols) -> "DMatrix":
        """dmatrix.cxx:367-394 — by reference."""
        out = _lib.DMat()
        check(lib().candmc_dmat_slice(C.byref(self._c), firstrow, numrows, firstcol, numcols, C.byref(out)))
        pv = pview(out.pv.rrow, out.pv.rcol, self.pv.crow, self.pv.ccol, self.pv.cworld)
        return DMatrix(numrows, numcols, self.b, pv, lda=self.lda, data=int(out.data), _keep=self.tensor)

    def _new(self, n):
        import torch

        return torch.empty(max(n, 1), dtype=torch.float64, device="cuda")

    def get_contig(self, stream=None) -> "DMatrix":
        """dmatrix.cxx:470-484."""
        out = self._new(self.get_mysize())
        check(lib().candmc_dmat_get_contig(C.byref(self._c), out.data_ptr(), _stream(stream)))
        return DMatrix(self.nrow, self.ncol, self.b, self.pv, lda=self.get_mynrow(), data=out)

    def replicate_vertical(self, stream=None):
        """dmatrix.cxx:268-289 -> tensor of nrow * get_myncol() doubles."""
        out = self._new(self.nrow * self.get_myncol())
        check(lib().candmc_dmat_replicate_vertical(C.byref(self._c), out.data_ptr(), _stream(stream)))
        return out

    def replicate_horizontal(self, stream=None):
        """dmatrix.cxx:294-304 -> tensor of ncol * get_mynrow() doubles."""
        out = self._new(self.ncol * self.get_mynrow())
        check(lib().candmc_dmat_replicate_horizontal(C.byref(self._c), out.data_ptr(), _stream(stream)))
        return out

    def reduce_scatter_horizontal(self, cntrb, stream=None):
        """dmatrix.cxx:310-355; `cntrb` (ncol * get_mynrow() doubles on the device) is scratch afterwards."""
        check(lib().candmc_dmat_reduce_scatter_horizontal(C.byref(self._c), cntrb.data_ptr(), _stream(stream)))

    def transpose_data(self, stream=None) -> "DMatrix":
        """dmatrix.cxx:252-263."""
        out = self._new(self.get_mysize())
        check(lib().candmc_dmat_transpose_data(C.byref(self._c), out.data_ptr(), _stream(stream)))
        return DMatrix(self.nrow, self.ncol, self.b, self.pv, lda=self.get_mynrow(), data=out)

    def foldcols(self, factor, stream=None) -> "DMatrix":
        """dmatrix.cxx:527-552."""
        out = self._new(self.get_mysize())
        check(lib().candmc_dmat_foldcols(C.byref(self._c), factor, out.data_ptr(), _stream(stream)))
        return DMatrix(self.nrow // factor, self.ncol * factor, self.b, self.pv, lda=self.get_mynrow() // factor, data=out)

    def foldrows(self, factor, stream=None) -> "DMatrix":
        """dmatrix.cxx:560-584."""
        out = self._new(self.get_mysize())
        check(lib().candmc_dmat_foldrows(C.byref(self._c), factor, out.data_ptr(), _stream(stream)))
        return DMatrix(self.nrow * factor, self.ncol // factor, self.b, self.pv, lda=self.get_mynrow() * factor, data=out)
