"""Host-side mirror of the reference's LU accelerator seam (alg/LU/lu_offload.h:19-101) over the C ABI candmc_off_*.

Same names, argument order and meaning as the reference: the three offloaded matrices OFF_A / OFF_L / OFF_U live in HBM,
`offload_gemm_A` multiplies sub-blocks addressed by (matrix, element offset, leading dimension) asynchronously,
`wait_gemm` joins, `upload_lda_cpy` / `download_lda_cpy` move strided blocks and `offload_sparse_rw` moves pivot rows.
Host operands are numpy float64 arrays (any writable buffer for outputs).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib, CandmcError, last_error  # noqa: F401

OFF_A, OFF_L, OFF_U = 0, 1, 2


def _ck(rc):
    if rc:
        raise CandmcError(rc, last_error())


def _hptr(a: np.ndarray) -> int:
    if a.dtype != np.float64:
        raise TypeError("host operands must be float64")
    return a.ctypes.data


def set_mic_rank(mic_rank: int):
    """lu_offload.cxx:126-128 — here: bind to GPU (rank mod #GPUs)."""
    _ck(lib().candmc_off_set_device(int(mic_rank)))


def alloc_A(size: int, ptr: np.ndarray | None = None):
    """lu_offload.cxx:479-491; `ptr` (if given) is uploaded."""
    _ck(lib().candmc_off_alloc(OFF_A, int(size), None if ptr is None else _hptr(ptr)))


def alloc_L(size: int):
    _ck(lib().candmc_off_alloc(OFF_L, int(size), None))


def alloc_U(size: int):
    _ck(lib().candmc_off_alloc(OFF_U, int(size), None))


def alloc_transfer(size: int):
    _ck(lib().candmc_off_alloc_transfer(int(size)))


def free_offload_A():
    _ck(lib().candmc_off_free(OFF_A))


def free_offload_L():
    _ck(lib().candmc_off_free(OFF_L))


def free_offload_U():
    _ck(lib().candmc_off_free(OFF_U))


def free_offload_transfer():
    _ck(lib().candmc_off_free_transfer())


def get_mat_handle(omat: int) -> np.ndarray:
    """Host-side get_mat_handle (lu_offload.cxx:159-175): a writable float64 view of the pinned mirror holding the
    matrix's current contents; what is written there is uploaded before the next operation on the matrix."""
    p = C.c_void_p()
    _ck(lib().candmc_off_host_mirror(int(omat), C.byref(p)))
    n = C.c_int64()
    _ck(lib().candmc_off_size(int(omat), C.byref(n)))
    n = n.value
    buf = (C.c_double * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64, count=n)


def device_handle(omat: int):
    """(device pointer, size in doubles) of an offloaded matrix — get_mat_handle as seen ON the accelerator."""
    p = C.c_void_p()
    n = C.c_int64()
    _ck(lib().candmc_off_device_ptr(int(omat), C.byref(p), C.byref(n)))
    return p.value, n.value


def wait_gemm():
    _ck(lib().candmc_off_wait_gemm())


def offload_gemm_A(tA, tB, m, n, k, alpha, offset_A, omat_A, lda_A, offset_B, omat_B, lda_B, beta, offset_C, omat_C,
                   lda_C):
    """lu_offload.cxx:216-251."""
    _ck(lib().candmc_off_gemm(tA.encode()[:1], tB.encode()[:1], m, n, k, alpha, offset_A, omat_A, lda_A, offset_B, omat_B,
                              lda_B, beta, offset_C, omat_C, lda_C))


def download_lda_cpy(nrow, ncol, lda_A, lda_B, offset_A, B: np.ndarray, omat_A):
    """lu_offload.cxx:338-364."""
    _ck(lib().candmc_off_download(nrow, ncol, lda_A, lda_B, offset_A, _hptr(B), omat_A))


def upload_lda_cpy(nrow, ncol, lda_A, lda_B, A: np.ndarray, offset_B, omat_B):
    """lu_offload.cxx:366-392."""
    _ck(lib().candmc_off_upload(nrow, ncol, lda_A, lda_B, _hptr(A), offset_B, omat_B))


def offload_sparse_rw(nrow, ncol, lda_B, A: np.ndarray, lda_A, offsets_transfer, omat_B, rw: str):
    """lu_offload.cxx:424-476."""
    offs = np.ascontiguousarray(offsets_transfer, dtype=np.int32)
    _ck(lib().candmc_off_sparse_rw(nrow, ncol, lda_B, _hptr(A), lda_A, offs.ctypes.data_as(C.POINTER(C.c_int)), omat_B,
                                   rw.encode()[:1]))


def sync():
    _ck(lib().candmc_off_sync())


def set_overlap(enable: bool):
    _ck(lib().candmc_off_set_overlap(1 if enable else 0))


def stats() -> dict:
    out = (C.c_int64 * 6)()
    _ck(lib().candmc_off_stats(out))
    keys = ("gemms", "uploads", "downloads", "sparse", "cross_stream_waits", "tracked_gemms")
    return dict(zip(keys, list(out)))
