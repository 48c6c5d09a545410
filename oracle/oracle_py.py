"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module; the
product package `candmc_b200` never does (tests/test_boundary.py checks that).
All matrices are numpy float64, column-major ("F" order) like the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_pd = C.POINTER(C.c_double)
_ppd = C.POINTER(_pd)
_i64 = C.c_int64


def build() -> str:
    """Compile the C restatement (gcc only; no GPU, no reference sources needed)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return os.path.join(_HERE, "liboracle.so")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "candmc_oracle.c")

        def stale():
            return not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src)

        if stale():
            # several ranks of one job may get here together (a snapshot that did not keep modification times): one builds, the
            # others wait for it — a rank must never dlopen a half-written file
            import fcntl

            with open(os.path.join(_HERE, ".liboracle.lock"), "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    if stale():
                        build()
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
        L = C.CDLL(path)
        L.oracle_unit_elem.restype = C.c_double
        L.oracle_unit_elem.argtypes = [_i64, _i64, _i64, C.c_int]
        L.oracle_fill_unit_block.argtypes = [_pd, _i64, _i64, _i64, _i64, _i64, _i64, C.c_int]
        L.oracle_dgemm.argtypes = [C.c_char, C.c_char, _i64, _i64, _i64, C.c_double, _pd, _i64, _pd, _i64,
                                   C.c_double, _pd, _i64]
        L.oracle_lda_cpy.argtypes = [_i64, _i64, _i64, _i64, _pd, _pd]
        L.oracle_lda_cpy_scaled.argtypes = [_i64, _i64, _i64, _i64, _pd, _pd, C.c_double, C.c_double]
        L.oracle_transpose.argtypes = [_i64, _i64, _pd, _i64, _pd, _i64]
        L.oracle_summa.argtypes = [_i64, C.c_int, C.c_char, C.c_char, _ppd, _i64, _ppd, _i64, _ppd, _i64]
        L.oracle_d25_summa.argtypes = [_i64, C.c_int, C.c_int, C.c_int, C.c_char, C.c_char, _ppd, _ppd, _ppd]
        L.oracle_bcast_cannon_4d.argtypes = [_i64, C.c_int, C.c_int, C.c_int, _ppd, _ppd, _ppd]
        L.oracle_bcast_cannon_4d_t.argtypes = [_i64, C.c_int, C.c_int, C.c_int, C.c_char, C.c_char, _ppd, _ppd, _ppd]
        L.oracle_spcannon.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char, C.c_double,
                                      _ppd, C.c_char, C.c_double, _ppd, _ppd]
        L.oracle_upd_A.argtypes = [C.c_int, C.POINTER(_i64), _i64, _i64, _ppd, C.POINTER(_i64), _ppd,
                                   C.POINTER(_i64), _pd]
        L.oracle_update_A.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i64, _i64, _i64, _ppd, _ppd, _pd, C.c_int,
                                      C.POINTER(_i64), C.POINTER(_i64)]
        L.oracle_update_Yamamoto_A.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i64, _i64, _i64, _ppd, _ppd, _pd]
        L.oracle_update_A_extents.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i64, _i64, _i64,
                                              C.POINTER(_i64), C.POINTER(_i64)]
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_pd)


def _pp(blocks):
    arr = (_pd * len(blocks))(*[_p(b) for b in blocks])
    return arr


def _ch(c: str) -> bytes:
    return c.encode("ascii")


# ---- generators ---------------------------------------------------------------------------------------------
def unit_block(nrow, ncol, row0, col0, n, which):
    """Block of the reference unit-test matrices (test/MM/topo_pdgemm_unit.cxx:250-256), column-major."""
    X = np.zeros((nrow, ncol), dtype=np.float64, order="F")
    lib().oracle_fill_unit_block(_p(X), nrow, ncol, nrow, row0, col0, n, which)
    return X


def drand48_stream(seed, count):
    """`count` successive drand48() draws after srand48(seed) — vectorised LCG (test/MM/test_spc.cxx:66-76)."""
    a, c, mask = 0x5DEECE66D, 0xB, (1 << 48) - 1
    x = ((seed & 0xFFFFFFFF) << 16) | 0x330E
    out = np.empty(count, dtype=np.float64)
    for i in range(count):
        x = (a * x + c) & mask
        out[i] = x / 281474976710656.0
    return out


# ---- local kernels --------------------------------------------------------------------------------------------
def dgemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
    lib().oracle_dgemm(_ch(ta), _ch(tb), m, n, k, alpha, _p(A), lda, _p(B), ldb, beta, _p(Cm), ldc)


def lda_cpy(nrow, ncol, lda_A, lda_B, A, B, a=None, b=None):
    if a is None:
        lib().oracle_lda_cpy(nrow, ncol, lda_A, lda_B, _p(A), _p(B))
    else:
        lib().oracle_lda_cpy_scaled(nrow, ncol, lda_A, lda_B, _p(A), _p(B), a, b)


def transpose(rows, cols, A, lda, B, ldb):
    lib().oracle_transpose(rows, cols, _p(A), lda, _p(B), ldb)


# ---- distributed (all ranks simulated in-process; blocks = list of per-rank arrays) -------------------------------
def summa(n, q, A, B, Cb, lda_A=None, lda_B=None, lda_C=None, trans_A="N", trans_B="N"):
    b = n // q
    rc = lib().oracle_summa(n, q, _ch(trans_A), _ch(trans_B), _pp(A), lda_A or b, _pp(B), lda_B or b, _pp(Cb),
                            lda_C or b)
    assert rc == 0, "oracle_summa: bad grid"


def d25_summa(n, q, c, ovp, A, B, Cb, trans_A="N", trans_B="N"):
    rc = lib().oracle_d25_summa(n, q, c, ovp, _ch(trans_A), _ch(trans_B), _pp(A), _pp(B), _pp(Cb))
    assert rc == 0, "oracle_d25_summa: bad grid"


def bcast_cannon_4d(n, x1_np, x2_np, ovp, A, B, Cb, trans_A="N", trans_B="N"):
    rc = lib().oracle_bcast_cannon_4d_t(n, x1_np, x2_np, ovp, _ch(trans_A), _ch(trans_B), _pp(A), _pp(B), _pp(Cb))
    assert rc == 0, "oracle_bcast_cannon_4d: bad grid"


def spcannon(bidir, kary, ndim, n, m, k, tA, alpha, A, tB, beta, B, Cb):
    rc = lib().oracle_spcannon(bidir, kary, ndim, n, m, k, _ch(tA), alpha, _pp(A), _ch(tB), beta, _pp(B), _pp(Cb))
    assert rc == 0, "oracle_spcannon: bad grid"


def upd_A(mb, kb, b, Y, lda_Y, A, lda_A, T):
    nprow = len(mb)
    I = _i64 * nprow
    rc = lib().oracle_upd_A(nprow, I(*mb), kb, b, _pp(Y), I(*lda_Y), _pp(A), I(*lda_A), _p(T))
    assert rc == 0, "oracle_upd_A: bad arguments"


# ---- the reference tests' data layouts ------------------------------------------------------------------------------
def d25_blocks(n, q, c):
    """Per-rank A, B blocks of d25_unit (test/MM/topo_pdgemm_unit.cxx:250-256): rank = layer*q*q + row*q + col holds
    global rows [row*b,(row+1)*b) x cols [col*b,(col+1)*b) of both A and B (replicated over layers)."""
    b = n // q
    A, B = [], []
    for layer in range(c):
        for row in range(q):
            for col in range(q):
                A.append(unit_block(b, b, row * b, col * b, n, 0))
                B.append(unit_block(b, b, row * b, col * b, n, 1))
    return A, B


def dcn_blocks(n, x1_np, x2_np):
    """Per-rank blocks of dcn_unit (test/MM/topo_pdgemm_unit.cxx:87-94): col = (x1*x2_np+x2)*b, row = (y1*x2_np+y2)*b."""
    b = n // (x1_np * x2_np)
    A, B = [], []
    for r in range(x1_np * x1_np * x2_np * x2_np):
        x1 = r % x1_np
        y1 = (r // x1_np) % x1_np
        x2 = (r // (x1_np * x1_np)) % x2_np
        y2 = r // (x1_np * x1_np * x2_np)
        A.append(unit_block(b, b, (y1 * x2_np + y2) * b, (x1 * x2_np + x2) * b, n, 0))
        B.append(unit_block(b, b, (y1 * x2_np + y2) * b, (x1 * x2_np + x2) * b, n, 1))
    return A, B


def spc_blocks(kary, ndim, seed, n, m, k, tB="N"):
    """Per-rank A (m x k), B (k x n, or n x k when tB == 'T'), C (m x n) blocks and the full matrices of test_spc
    (test/MM/test_spc.cxx:56-102)."""
    khalf = kary ** (ndim // 2)
    s = drand48_stream(seed, (m * k + k * n + m * n) * khalf * khalf)
    o = 0
    full_A = s[o:o + m * khalf * k * khalf].reshape((m * khalf, k * khalf), order="F"); o += full_A.size
    full_B = s[o:o + k * khalf * n * khalf].reshape((k * khalf, n * khalf), order="F"); o += full_B.size
    full_C = s[o:o + m * khalf * n * khalf].reshape((m * khalf, n * khalf), order="F")
    A, B, Cb = [], [], []
    for rank in range(kary ** ndim):
        px = py = 0
        sc, tr = 1, rank
        for _ in range(ndim // 2):
            px += (tr % kary) * sc; tr //= kary
            py += (tr % kary) * sc; tr //= kary
            sc *= kary
        A.append(np.asfortranarray(full_A[py * m:(py + 1) * m, px * k:(px + 1) * k]))
        Bb = full_B[py * k:(py + 1) * k, px * n:(px + 1) * n]
        B.append(np.asfortranarray(Bb.T if tB == "T" else Bb))
        Cb.append(np.asfortranarray(full_C[py * m:(py + 1) * m, px * n:(px + 1) * n]))
    return A, B, Cb, (full_A, full_B, full_C)


# ---- CAQR trailing update (N1): update_A on a block-cyclic nprow x npcol grid -------------------------------------------
def update_A_extents(nprow, npcol, rrow, rcol, myrow, mycol, m, k, b):
    mb, kb = _i64(), _i64()
    lib().oracle_update_A_extents(nprow, npcol, rrow, rcol, myrow, mycol, m, k, b, C.byref(mb), C.byref(kb))
    return mb.value, kb.value


def _lcg48(seed):
    x = ((seed & 0xFFFFFFFF) << 16) | 0x330E
    x = (0x5DEECE66D * x + 0xB) & ((1 << 48) - 1)
    return x / 281474976710656.0


def update_A_blocks(nprow, npcol, rrow, rcol, m, k, b):
    """Per-rank Y (mb x b) and A (mb x kb) blocks of oracle/ref_dump.cxx's `upda` mode (rank = myrow + mycol*nprow):
    elements are seeded by their global coordinates in the remaining matrix."""
    Y, A = [], []
    for rank in range(nprow * npcol):
        myrow, mycol = rank % nprow, rank // nprow
        mb, kb = update_A_extents(nprow, npcol, rrow, rcol, myrow, mycol, m, k, b)
        Yb = np.zeros((max(mb, 1), b), order="F")
        Ab = np.zeros((max(mb, 1), max(kb, 1)), order="F")
        for r in range(mb):
            gr = ((r // b) * nprow + (myrow - rrow + nprow) % nprow) * b + r % b
            for j in range(b):
                Yb[r, j] = (_lcg48(7000 + gr * b + j) - .5) * 0.25
            for cc in range(kb):
                gc = ((cc // b) * npcol + (mycol - rcol - 1 + npcol) % npcol) * b + cc % b
                Ab[r, cc] = _lcg48(900000 + gc * m + gr) - .5
        Y.append(Yb[:mb] if mb else Yb[:0])
        A.append(Ab[:mb, :kb] if (mb and kb) else np.zeros((mb, kb), order="F"))
    Y = [np.asfortranarray(y) for y in Y]
    A = [np.asfortranarray(a) for a in A]
    return Y, A


def update_A(nprow, npcol, rrow, rcol, m, k, b, Y, A, W=None, W_is_T=True):
    """All ranks simulated; A blocks are updated in place.  W None -> T from Y; W_is_T -> W is the lower-triangular T;
    otherwise W is the upper-triangular factor of the panel QR and T comes from comp_bcast_T_from_W (qr_2d.cxx:179-208)."""
    Yp = [y if y.size else np.zeros(1) for y in Y]
    Ap = [a if a.size else np.zeros(1) for a in A]
    rc = lib().oracle_update_A(nprow, npcol, rrow, rcol, m, k, b, _pp(Yp), _pp(Ap),
                               _p(W) if W is not None else None, 0 if W is None else (1 if W_is_T else 2), None, None)
    assert rc == 0, "oracle_update_A: bad arguments"


def panel_W(b):
    """The b x b upper-triangular W of oracle/ref_dump.cxx's `updw` mode (diagonal in [1, 1.5), small off-diagonal entries;
    the strict lower triangle is never read: cdtrsm('L','U',...))."""
    W = np.zeros((b, b), order="F")
    for j in range(b):
        for i in range(j + 1):
            v = _lcg48(333000 + i + j * b)
            W[i, j] = 1.0 + 0.5 * v if i == j else (v - .5) * 0.2
    return W


def yamamoto_T(b):
    """The b x b T of oracle/ref_dump.cxx's `updy` mode (a generic dense matrix: the update never assumes structure)."""
    T = np.zeros((b, b), order="F")
    for j in range(b):
        for i in range(b):
            T[i, j] = (_lcg48(555000 + i + j * b) - .5) * 0.5
    return T


def update_Yamamoto_A(nprow, npcol, rrow, rcol, m, k, b, Qm, A, T):
    """All ranks simulated; A blocks are updated in place (alg/QR/qr_2d/qr_y2d.cxx:68-169, agg == NULL)."""
    Qp = [q if q.size else np.zeros(1) for q in Qm]
    Ap = [a if a.size else np.zeros(1) for a in A]
    rc = lib().oracle_update_Yamamoto_A(nprow, npcol, rrow, rcol, m, k, b, _pp(Qp), _pp(Ap), _p(T))
    assert rc == 0, "oracle_update_Yamamoto_A: bad arguments"


# ---- update_Yamamoto_A with an aggregator (alg/QR/qr_2d/qr_y2d.cxx:38-62,68-120), driven as QR_Yamamoto_2D drives it (:171-277) ----
# Restated on the GLOBAL block column (every rank's local arrays are block-cyclic pieces of it), pinned to the unmodified
# reference by the `updyagg_*` fixtures of tests/golden/canmm_ref_outputs.npz (oracle/ref_dump.cxx, mode updyagg).
def _lcg48_vec(seeds):
    a, c, mask, lo24 = np.uint64(0x5DEECE66D), np.uint64(0xB), np.uint64((1 << 48) - 1), np.uint64((1 << 24) - 1)
    x = ((np.asarray(seeds, dtype=np.uint64) & np.uint64(0xFFFFFFFF)) << np.uint64(16)) | np.uint64(0x330E)
    xl, xh = x & lo24, x >> np.uint64(24)
    x = (a * xl + (((a * xh) & lo24) << np.uint64(24)) + c) & mask
    return x.astype(np.float64) / 281474976710656.0


def yamamoto_agg_inputs(m, k, b, s=None):
    """ref_dump's `updyagg` generators: the m x k block column (s is None), or step s's panel Qm (rows s*b .. m of the original
    matrix, b columns) and its b x b T."""
    if s is None:
        gr, gc = np.meshgrid(np.arange(m, dtype=np.uint64), np.arange(k, dtype=np.uint64), indexing="ij")
        return np.asfortranarray(_lcg48_vec(np.uint64(900000) + gc * np.uint64(m) + gr) - .5)
    gr, j = np.meshgrid(np.arange(s * b, m, dtype=np.uint64), np.arange(b, dtype=np.uint64), indexing="ij")
    Qm = (_lcg48_vec(np.uint64(7000 + 131 * s) + gr * np.uint64(b) + j) - .5) * 0.25
    i, j = np.meshgrid(np.arange(b, dtype=np.uint64), np.arange(b, dtype=np.uint64), indexing="ij")
    T = (_lcg48_vec(np.uint64(555000 + 977 * s) + i + j * np.uint64(b)) - .5) * 0.5
    return np.asfortranarray(Qm), np.asfortranarray(T)


def yamamoto_aggregate(m, k, b):
    """Global result of the k/b steps: (A after the trailing updates, aQm (m x k), aT (k x k)).
    Step s (qr_y2d.cxx:123-169): W = Qm^T A_trail, W2 = -T W, A_trail -= Qm W2 on rows s*b.., columns (s+1)*b..;
    append (:38-62): aQm[s*b.., n..n+b] = Qm; aT[0:b,0:b] = T for the first panel, afterwards
    aT[n..n+b, 0..n] = T ((Qm^T aQm[s*b.., 0..n]) aT[0..n,0..n]) and aT[n..n+b, n..n+b] = T."""
    A = yamamoto_agg_inputs(m, k, b)
    aQm = np.zeros((m, k), order="F"); aT = np.zeros((k, k), order="F")
    n = 0
    for s in range(k // b):
        Qm, T = yamamoto_agg_inputs(m, k, b, s)
        if k - s * b - b > 0 and m - s * b - b > 0:
            tr = A[s * b:, (s + 1) * b:]
            tr -= Qm @ (-T @ (Qm.T @ tr))
        aQm[s * b:, n:n + b] = Qm
        if n == 0:
            aT[:b, :b] = T
        else:
            aT[n:n + b, :n] = T @ ((Qm.T @ aQm[s * b:, :n]) @ aT[:n, :n])
            aT[n:n + b, n:n + b] = T
        n += b
    return A, aQm, aT


def cyclic_local(G, b, nprow, npcol, rrow, rcol, myrow, mycol):
    """rank (myrow, mycol)'s block-cyclic piece of the global matrix G (blocks of b, roots rrow / rcol own block 0)"""
    rb = [g for g in range(G.shape[0] // b) if (g - (myrow - rrow)) % nprow == 0]
    cb = [g for g in range(G.shape[1] // b) if (g - (mycol - rcol)) % npcol == 0]
    rows = np.concatenate([np.arange(g * b, (g + 1) * b) for g in rb]) if rb else np.zeros(0, dtype=int)
    cols = np.concatenate([np.arange(g * b, (g + 1) * b) for g in cb]) if cb else np.zeros(0, dtype=int)
    return np.asfortranarray(G[np.ix_(rows, cols)])


# ---- DMatrix pack / replication operations (alg/SE/dmatrix.cxx), all ranks simulated in numpy ------------------------------
# Data movement only (plus one sum), so numpy index arithmetic is the restatement; pinned to the unmodified reference by
# tests/golden/dmat_ref_outputs.npz (oracle/ref_dmat_dump.cxx).  Grid rank = myrow + mycol*nprow (test_qr_2d.cxx:367-374).
def dmat_extent(n, b, np_, rank, root):
    """get_mynrow / get_myncol, dmatrix.cxx:194-203"""
    nb = n // b
    return (nb // np_ + (1 if (nb % np_) > (rank + np_ - root) % np_ else 0)) * b


def dmat_local(nrow, ncol, b, nprow, npcol, rrow, rcol, myrow, mycol, value):
    """Local piece of the matrix whose global element (gr, gc) is value(gr, gc) (vectorised callable)."""
    mr, mc = dmat_extent(nrow, b, nprow, myrow, rrow), dmat_extent(ncol, b, npcol, mycol, rcol)
    r, c = np.arange(mr), np.arange(mc)
    gr = ((r // b) * nprow + (myrow - rrow) % nprow) * b + r % b
    gc = ((c // b) * npcol + (mycol - rcol) % npcol) * b + c % b
    return np.asfortranarray(value(gr[:, None], gc[None, :]))


def dmat_replicate_vertical(pieces, nprow, npcol):
    """dmatrix.cxx:268-289: per rank, the packed pieces of its grid column in row-rank order"""
    return [np.concatenate([pieces[pr + (rank // nprow) * nprow].reshape(-1, order="F") for pr in range(nprow)])
            for rank in range(nprow * npcol)]


def dmat_replicate_horizontal(pieces, nprow, npcol):
    """dmatrix.cxx:294-304"""
    return [np.concatenate([pieces[(rank % nprow) + pc * nprow].reshape(-1, order="F") for pc in range(npcol)])
            for rank in range(nprow * npcol)]


def dmat_reduce_scatter_horizontal(pieces, cntrbs, nprow, npcol):
    """dmatrix.cxx:310-355: data += sum over my grid row of chunk <my column> of every contribution"""
    out = []
    for rank in range(nprow * npcol):
        myrow, mycol = rank % nprow, rank // nprow
        d = pieces[rank].reshape(-1, order="F").copy()
        for pc in range(npcol):
            d += cntrbs[myrow + pc * nprow][mycol * d.size:(mycol + 1) * d.size]
        out.append(d)
    return out


def dmat_transpose_data(pieces, nprow, npcol):
    """dmatrix.cxx:252-263: the packed piece of world rank crow.rank + ccol.rank*npcol"""
    return [pieces[(rank // nprow) + (rank % nprow) * npcol].reshape(-1, order="F").copy() for rank in range(nprow * npcol)]


def dmat_foldcols(piece, b, factor):
    """dmatrix.cxx:527-552 on one local piece (mr x mc) -> (mr/f) x (mc*f)"""
    mr, mc = piece.shape
    blocks = piece.reshape(mr // (b * factor), factor, b, mc)            # [j, i, w, c]
    return np.asfortranarray(np.concatenate([blocks[:, i].reshape(mr // factor, mc) for i in range(factor)], axis=1))


def dmat_foldrows(piece, b, factor):
    """dmatrix.cxx:560-584 on one local piece (mr x mc) -> (mr*f) x (mc/f)"""
    mr, mc = piece.shape
    bcol = mc // factor
    out = np.empty((mr // b, factor, b, bcol))
    for i in range(factor):
        out[:, i] = piece[:, i * bcol:(i + 1) * bcol].reshape(mr // b, b, bcol)
    return np.asfortranarray(out.reshape(mr * factor, bcol))


def dmat_slice(piece, nrow, ncol, b, nprow, npcol, rrow, rcol, myrow, mycol, firstrow, numrows, firstcol, numcols):
    """slice, dmatrix.cxx:367-394: returns (view of the local piece, new rrow, new rcol).  The roots rotate to the owner of the
    corner block; the local pointer moves past the rows / columns this rank owns above / left of the corner."""
    nr, nc = (rrow + firstrow // b) % nprow, (rcol + firstcol // b) % npcol
    mr0, mc0 = dmat_extent(nrow, b, nprow, myrow, rrow), dmat_extent(ncol, b, npcol, mycol, rcol)
    mr1, mc1 = dmat_extent(nrow - firstrow, b, nprow, myrow, nr), dmat_extent(ncol - firstcol, b, npcol, mycol, nc)
    mr2, mc2 = dmat_extent(numrows, b, nprow, myrow, nr), dmat_extent(numcols, b, npcol, mycol, nc)
    return piece[mr0 - mr1:mr0 - mr1 + mr2, mc0 - mc1:mc0 - mc1 + mc2], nr, nc


# ---- symmetric full -> band trailing update (alg/SE/full_to_band.cxx:28-250), all ranks simulated ------------
def _cmod(a, p):
    """C's truncating % (the reference's extent formulas subtract before taking the remainder, full_to_band.cxx:70,77)."""
    return int(np.fmod(a, p))


def f2b_level(n, b, b_sub, pr, rrow, rcol, myrow, mycol):
    """Offsets and extents one level of sym_full2band uses on rank (myrow, mycol) of a pr x pr grid (full_to_band.cxx:57-79):
    (loc_row_offset, loc_col_offset in columns, mb, kb)."""
    s = b // b_sub
    ro = b_sub * (b // (b_sub * pr)) + (b_sub if (myrow + pr - rrow) % pr < s % pr else 0)
    co = b_sub * (b // (b_sub * pr)) + (b_sub if (mycol + pr - rcol) % pr < s % pr else 0)
    t = (n - b) // b_sub
    mb = (t // pr + (1 if _cmod(myrow + pr - rrow - s % pr, pr) < t % pr else 0)) * b_sub
    kb = (t // pr + (1 if _cmod(mycol + pr - rcol - s % pr, pr) < t % pr else 0)) * b_sub
    return ro, co, mb, kb


def f2b_invT(Ycol, b):
    """compute_invT_from_Y (alg/QR/qr_2d/qr_2d.cxx:22-60): lower triangle of the column sum of Y_i^T Y_i, diagonal halved."""
    S = np.zeros((b, b))
    for Y in Ycol:
        if Y.shape[0]:
            S += Y.T @ Y
    T = np.tril(S)
    T[np.diag_indices(b)] *= 0.5
    return T


def f2b_update(n, b, b_sub, pr, rrow, rcol, A, Y):
    """One level's trailing update of sym_full2band, in place on A.
    A[r]: rank r's local array with its corner at this level's working corner (any lda; a numpy view), r = myrow + mycol*pr;
    Y[r]: the aggregated panel the QR left on rank r (mb x b).  Follows full_to_band.cxx:90,103-239 step by step."""
    lev = {(i, j): f2b_level(n, b, b_sub, pr, rrow, rcol, i, j) for i in range(pr) for j in range(pr)}
    rk = lambda i, j: i + j * pr  # noqa: E731
    rrow2 = (rrow + b // b_sub) % pr                                     # :90
    invT = f2b_invT([Y[rk(i, rcol)][:lev[(i, rcol)][2]] for i in range(pr)], b)     # root column, then bcast (:103)
    W = {}
    for j in range(pr):                                                  # W = Y^T A reduced onto the diagonal (:121-131)
        acc = None
        for i in range(pr):
            ro, co, mb, kb = lev[(i, j)]
            Wi = Y[rk(i, j)][:mb].T @ A[rk(i, j)][ro:ro + mb, co:co + kb] if (mb and kb) else np.zeros((b, kb))
            acc = Wi if acc is None else acc + Wi
        W[j] = acc
    Z = np.zeros((b, b))
    for d in range(pr):                                                  # Z = Y^T W^T summed over the diagonal (:156-161)
        lb = lev[(d, d)][2]
        assert lb == lev[(d, d)][3]
        if lb:
            Z += Y[rk(d, d)][:lb].T @ W[d].T
    U, V = {}, {}
    for d in range(pr):
        lb = lev[(d, d)][2]
        Yd = Y[rk(d, d)][:lb]
        if lb:
            U[d] = np.linalg.solve(invT.T, Yd.T).T                       # U invT = Y (cdtrsm R,L,N,N, :172)
            V[d] = W[d] - 0.5 * Z @ U[d].T                               # V' = W - .5 Z U' (:186)
        else:
            U[d], V[d] = Yd, W[d]
    UVT = {}
    for i in range(pr):
        for j in range(pr):
            ro, co, mb, kb = lev[(i, j)]
            UVT[(i, j)] = U[i] @ V[j] if (mb and kb) else None           # U along rows, V' along columns (:205-215)
    for i in range(pr):
        for j in range(pr):
            ro, co, mb, kb = lev[(i, j)]
            if mb and kb:                                                # A -= UV' + (partner's UV')^T (:222-239)
                A[rk(i, j)][ro:ro + mb, co:co + kb] -= UVT[(i, j)] + UVT[(j, i)].T
    return rrow2, (rcol + b // b_sub) % pr, invT


def f2b_sym_value(n, seed=23):
    """The dump tool's symmetric test matrix: element (i, j) = off_value(seed, min + max*n)."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "tests"))
    from off_script import off_value
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    table = off_value(seed, n * n)
    return table[lo + hi * n]
