"""Operation scripts for the LU accelerator seam and two of their three interpreters.

A script is a text of one operation per line (grammar in oracle/ref_off_dump.cxx, the third interpreter, which feeds
it to the UNMODIFIED reference alg/LU/lu_offload.cxx).  Here:
  * `make_scripts()`      the seeded scripts committed (inside tests/golden/lu_offload_ref_outputs.npz) with the
                          reference's outputs for them;
  * `run_script(text, backend)`  executes a script against a backend and returns the list of observable outputs in the
                          order ref_off_dump writes them (downloads, sparse reads/swaps, then each matrix whole);
  * `OracleBackend`       oracle/liboracle.so's oracle_off_* (plain-C restatement of the reference's host fallback);
  * `GpuBackend`          candmc_b200.lu_offload (the CUDA path, through the C ABI).
Test infrastructure only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_M64 = (1 << 64) - 1


def off_value(seed: int, n: int) -> np.ndarray:
    """oracle_off_value(seed, 0..n) — splitmix64 finaliser mapped to [-0.5, 0.5) (oracle/candmc_oracle.c)."""
    with np.errstate(over="ignore"):
        x = np.uint64((seed * 0x9E3779B97F4A7C15) & _M64) + np.arange(n, dtype=np.uint64)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


# ---------------------------------------------------------------------------------------------------------------------
class OracleBackend:
    """oracle_off_* of oracle/liboracle.so."""

    class _Off(C.Structure):
        _fields_ = [("mat", C.c_void_p * 3), ("size", C.c_int64 * 3)]

    def __init__(self):
        self.L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        i64, dp = C.c_int64, C.c_void_p
        self.L.oracle_off_handle.restype = C.c_void_p
        self.L.oracle_off_handle.argtypes = [C.c_void_p, C.c_int]
        self.L.oracle_off_alloc.argtypes = [C.c_void_p, C.c_int, i64]
        self.L.oracle_off_gemm.argtypes = [C.c_void_p, C.c_char, C.c_char, i64, i64, i64, C.c_double, i64, C.c_int, i64,
                                           i64, C.c_int, i64, C.c_double, i64, C.c_int, i64]
        self.L.oracle_off_upload.argtypes = [C.c_void_p, i64, i64, i64, i64, dp, i64, C.c_int]
        self.L.oracle_off_download.argtypes = [C.c_void_p, i64, i64, i64, i64, i64, dp, C.c_int]
        self.L.oracle_off_sparse_rw.argtypes = [C.c_void_p, i64, i64, i64, dp, i64, C.POINTER(C.c_int), C.c_int, C.c_char]
        self.o = self._Off()
        self.L.oracle_off_init(C.byref(self.o))
        self.sizes = {}

    def close(self):
        self.L.oracle_off_destroy(C.byref(self.o))

    def alloc(self, mat, size):
        assert self.L.oracle_off_alloc(C.byref(self.o), mat, size) == 0
        self.sizes[mat] = size

    def fill(self, mat, values):
        p = self.L.oracle_off_handle(C.byref(self.o), mat)
        C.memmove(p, values.ctypes.data, values.nbytes)

    def upload(self, nrow, ncol, lda_A, lda_B, A, off_B, mat):
        assert self.L.oracle_off_upload(C.byref(self.o), nrow, ncol, lda_A, lda_B, A.ctypes.data, off_B, mat) == 0

    def download(self, nrow, ncol, lda_A, lda_B, off_A, B, mat):
        assert self.L.oracle_off_download(C.byref(self.o), nrow, ncol, lda_A, lda_B, off_A, B.ctypes.data, mat) == 0

    def gemm(self, tA, tB, m, n, k, alpha, offA, matA, ldA, offB, matB, ldB, beta, offC, matC, ldC):
        assert self.L.oracle_off_gemm(C.byref(self.o), tA.encode(), tB.encode(), m, n, k, alpha, offA, matA, ldA, offB,
                                      matB, ldB, beta, offC, matC, ldC) == 0

    def wait(self):
        pass

    def sparse_rw(self, nrow, ncol, lda_B, A, lda_A, offs, mat, rw):
        o = np.ascontiguousarray(offs, dtype=np.int32)
        assert self.L.oracle_off_sparse_rw(C.byref(self.o), nrow, ncol, lda_B, A.ctypes.data, lda_A,
                                           o.ctypes.data_as(C.POINTER(C.c_int)), mat, rw.encode()) == 0


class GpuBackend:
    """candmc_b200.lu_offload — the product path."""

    def __init__(self, overlap=True):
        from candmc_b200 import lu_offload as lo
        self.lo = lo
        lo.set_overlap(overlap)
        self.sizes = {}

    def close(self):
        self.lo.sync()
        for mat, fn in ((0, self.lo.free_offload_A), (1, self.lo.free_offload_L), (2, self.lo.free_offload_U)):
            if mat in self.sizes:
                fn()
        self.lo.free_offload_transfer()

    def alloc(self, mat, size):
        (lambda s: self.lo.alloc_A(s, None), self.lo.alloc_L, self.lo.alloc_U)[mat](size)
        self.sizes[mat] = size

    def fill(self, mat, values):
        # exactly what the reference's driver does: memcpy through get_mat_handle (lu_25d_pvt.cxx:1600-1602)
        h = self.lo.get_mat_handle(mat)
        h[:] = values

    def upload(self, nrow, ncol, lda_A, lda_B, A, off_B, mat):
        self.lo.upload_lda_cpy(nrow, ncol, lda_A, lda_B, A, off_B, mat)
        A[:] = np.nan  # the source may be reused as soon as the call returns

    def download(self, nrow, ncol, lda_A, lda_B, off_A, B, mat):
        self.lo.download_lda_cpy(nrow, ncol, lda_A, lda_B, off_A, B, mat)

    def gemm(self, *a):
        self.lo.offload_gemm_A(*a)

    def wait(self):
        self.lo.wait_gemm()

    def sparse_rw(self, nrow, ncol, lda_B, A, lda_A, offs, mat, rw):
        self.lo.offload_sparse_rw(nrow, ncol, lda_B, A, lda_A, offs, mat, rw)


# ---------------------------------------------------------------------------------------------------------------------
def run_script(text: str, be) -> list:
    outs = []
    for line in text.strip().splitlines():
        t = line.split()
        if not t:
            continue
        op = t[0]
        if op == "alloc":
            be.alloc(int(t[1]), int(t[2]))
        elif op == "fill":
            mat = int(t[1])
            be.fill(mat, off_value(int(t[2]), be.sizes[mat]))
        elif op == "up":
            nrow, ncol, lda_A, lda_B, off_B, mat, seed = map(int, t[1:8])
            A = off_value(seed, lda_A * ncol + 1)
            be.upload(nrow, ncol, lda_A, lda_B, A, off_B, mat)
        elif op == "down":
            nrow, ncol, lda_A, lda_B, off_A, mat = map(int, t[1:7])
            B = np.full(lda_B * ncol + 1, -7.0)
            be.download(nrow, ncol, lda_A, lda_B, off_A, B, mat)
            outs.append(B)
        elif op == "gemm":
            tA, tB = t[1], t[2]
            m, n, k = map(int, t[3:6])
            alpha = float(t[6])
            offA, matA, ldA, offB, matB, ldB = map(int, t[7:13])
            beta = float(t[13])
            offC, matC, ldC = map(int, t[14:17])
            be.gemm(tA, tB, m, n, k, alpha, offA, matA, ldA, offB, matB, ldB, beta, offC, matC, ldC)
        elif op == "wait":
            be.wait()
        elif op == "sp":
            rw = t[1]
            nrow, ncol, lda_B, lda_A, mat, seed = map(int, t[2:8])
            offs = np.array(list(map(int, t[8:8 + nrow])), dtype=np.int32)
            A = off_value(seed, nrow * lda_A + 1)
            be.sparse_rw(nrow, ncol, lda_B, A, lda_A, offs, mat, rw)
            if rw != "w":
                outs.append(A)
        else:
            raise ValueError(f"unknown op {op}")
    be.wait()
    for mat in (0, 1, 2):
        if mat in be.sizes:
            n = be.sizes[mat]
            whole = np.full(n + 1, -7.0)
            if n > 0:
                be.download(n, 1, n, n, 0, whole, mat)
            outs.append(whole[:n].copy())
    return outs


# ---------------------------------------------------------------------------------------------------------------------
def _lu_like(ld: int, b: int, kpan: int, seed: int, ragged: bool) -> str:
    """The call pattern of the reference's fat-GEMM LU step (lu_25d_pvt.cxx:1227-1391, tnmt_pvt.cxx:640-706): for each big
    block upload the L and U panels, update the trailing block of A with alpha=-1 / beta=1, move pivot rows with sparse
    reads and writes, and download the next panel — without a wait between the GEMM and the transfers that follow."""
    rng = np.random.RandomState(seed)
    L = []
    size_A = ld * ld
    nb = ld // b
    pan = kpan * b
    L.append(f"alloc 0 {size_A}")
    L.append(f"alloc 1 {ld * pan}")
    L.append(f"alloc 2 {pan * ld}")
    L.append(f"fill 0 {seed}")
    L.append(f"fill 1 {seed + 1}")
    L.append(f"fill 2 {seed + 2}")
    s = seed + 10
    for ib in range(nb - 1):
        act = ld - (ib + 1) * b          # trailing extent
        off = (ib + 1) * b * ld + (ib + 1) * b
        if ragged and ib % 2 == 1:
            act -= 1                     # odd extents: the GEMM's unaligned (generic) path
        # panels: L is act x pan (ld = act), U is pan x act (ld = pan)
        L.append(f"up {act} {pan} {act + (3 if ragged else 0)} {act} 0 1 {s}"); s += 1
        L.append(f"up {pan} {act} {pan} {pan} 0 2 {s}"); s += 1
        L.append(f"gemm N N {act} {act} {pan} -1.0 0 1 {act} 0 2 {pan} 1.0 {off} 0 {ld}")
        # pivot rows of the trailing block: read some, write others (distinct rows), while the GEMM may still run
        nrow = min(b, act)
        rows = rng.permutation(act)[:nrow] + (ib + 1) * b
        col0 = (ib + 1) * b
        offs = " ".join(str(int(r + col0 * ld)) for r in rows)
        L.append(f"sp r {nrow} {act} {ld} {act + 2} 0 {s} {offs}"); s += 1
        rows = rng.permutation(act)[:nrow] + (ib + 1) * b
        offs = " ".join(str(int(r + col0 * ld)) for r in rows)
        L.append(f"sp w {nrow} {act} {ld} {act} 0 {s} {offs}"); s += 1
        if ib % 2 == 0:
            L.append("wait")
        # next panel: b columns of the trailing block, then its first b rows
        L.append(f"down {act} {min(b, act)} {ld} {act + 1} {off} 0")
        L.append(f"down {min(b, act)} {act} {ld} {b} {off} 0")
    return "\n".join(L) + "\n"


def _edge_cases(seed: int) -> str:
    """Everything the seam's argument space allows that the LU pattern above does not reach."""
    L = []
    L += ["alloc 0 4096", "alloc 1 2048", "alloc 2 2048", f"fill 0 {seed}", f"fill 1 {seed + 1}", f"fill 2 {seed + 2}"]
    s = seed + 10
    # contiguous copies (lda == nrow on both sides), single column, zero extents
    L.append(f"up 32 8 32 32 100 0 {s}"); s += 1
    L.append(f"up 17 1 40 64 3 1 {s}"); s += 1
    L.append(f"up 0 5 4 4 0 2 {s}"); s += 1
    L.append(f"up 5 0 5 5 0 2 {s}"); s += 1
    L.append("down 32 8 32 32 100 0")
    L.append("down 9 7 64 11 65 0")
    L.append("down 0 3 8 8 0 1")
    # transposed operands, odd sizes and offsets, beta = 0 and general alpha/beta, k = 0
    L.append("gemm T N 13 9 21 0.5 7 1 21 300 2 21 0.0 1001 0 64")
    L.append("gemm N T 16 24 8 -1.0 512 1 16 1024 2 24 1.0 2048 0 64")
    L.append("gemm T T 5 6 7 2.0 0 2 7 0 1 6 -0.5 77 0 64")
    L.append("gemm N N 32 32 0 1.0 0 1 32 0 2 1 0.25 0 0 64")
    L.append("gemm N N 64 16 64 1.0 0 0 64 0 1 64 0.0 0 2 64")   # A as an INPUT, U as the output
    L.append("down 64 16 64 64 0 2")
    L.append("gemm N N 16 16 16 1.0 0 1 16 256 1 16 1.0 3301 0 16")  # aligned inputs (tensor path), odd C offset
    L.append("down 16 16 16 16 3301 0")
    # back-to-back GEMMs into the same block without wait (in-order on the device)
    L.append("gemm N N 8 8 8 1.0 0 1 8 64 1 8 1.0 3000 0 8")
    L.append("gemm N N 8 8 8 1.0 128 1 8 192 1 8 1.0 3000 0 8")
    L.append("down 8 8 8 8 3000 0")
    # upload into a block a queued GEMM reads, then a GEMM that must see the new data
    L.append(f"up 8 8 8 8 0 1 {s}"); s += 1
    L.append("gemm N N 8 8 8 1.0 0 1 8 64 1 8 0.0 3100 0 8")
    L.append("down 8 8 8 8 3100 0")
    # sparse traffic: swap, duplicate rows on write (last one wins), duplicate rows on swap (chained), ncol == 1
    L.append(f"sp s 4 16 64 16 0 {s} 5 9 70 2"); s += 1
    L.append(f"sp w 3 10 64 12 0 {s} 20 20 21"); s += 1
    L.append(f"sp s 3 10 64 10 0 {s} 30 31 30"); s += 1
    L.append(f"sp w 2 4 8 4 1 {s} 3 19"); s += 1          # rows 3 and 19 = 3 + 2*8 share elements (column shift 2 < ncol)
    L.append(f"sp r 5 1 64 3 0 {s} 0 1 2 3 4095"); s += 1
    L.append(f"sp r 6 12 64 12 0 {s} 63 0 17 17 5 40"); s += 1
    L.append("sp r 0 4 64 4 0 1")
    return "\n".join(L) + "\n"


def random_script(seed: int, nops: int = 60) -> str:
    """A random, valid operation sequence over the three offloaded matrices, almost without waits: uploads, downloads, GEMMs
    (output matrix distinct from both inputs, as every call of the reference has it) and sparse row traffic with in-bounds
    random blocks — what the scoreboard of the two-stream seam has to order correctly.  Checked against the oracle, which is
    pinned to the unmodified lu_offload.cxx by the committed scripts."""
    rng = np.random.RandomState(seed)
    size = {0: 64 * 64, 1: 48 * 40, 2: 40 * 48}
    L = [f"alloc {m} {n}" for m, n in size.items()] + [f"fill {m} {seed + m}" for m in size]
    s = seed + 10

    def block(mat, rows, cols):
        """random (offset, ld) of a rows x cols block inside matrix `mat`"""
        ld = rows + int(rng.randint(0, 5))
        span = (cols - 1) * ld + rows
        return int(rng.randint(0, size[mat] - span + 1)), ld

    for _ in range(nops):
        op = rng.choice(["up", "down", "gemm", "gemm", "sp", "wait"], p=[0.2, 0.25, 0.15, 0.15, 0.2, 0.05])
        if op == "up":
            mat = int(rng.randint(0, 3)); nrow, ncol = int(rng.randint(1, 25)), int(rng.randint(1, 13))
            off, ldb = block(mat, nrow, ncol)
            L.append(f"up {nrow} {ncol} {nrow + int(rng.randint(0, 4))} {ldb} {off} {mat} {s}"); s += 1
        elif op == "down":
            mat = int(rng.randint(0, 3)); nrow, ncol = int(rng.randint(1, 25)), int(rng.randint(1, 13))
            off, lda = block(mat, nrow, ncol)
            L.append(f"down {nrow} {ncol} {lda} {nrow + int(rng.randint(0, 3))} {off} {mat}")
        elif op == "gemm":
            mc = int(rng.randint(0, 3)); others = [m for m in (0, 1, 2) if m != mc]
            ma, mb = int(rng.choice(others)), int(rng.choice(others))
            m, n, k = int(rng.randint(1, 17)), int(rng.randint(1, 17)), int(rng.randint(0, 17))
            tA, tB = rng.choice(["N", "T"]), rng.choice(["N", "T"])
            ra, ca = (m, max(k, 1)) if tA == "N" else (max(k, 1), m)
            rb, cb_ = (max(k, 1), n) if tB == "N" else (n, max(k, 1))
            offA, ldA = block(ma, ra, ca); offB, ldB = block(mb, rb, cb_); offC, ldC = block(mc, m, n)
            alpha, beta = rng.choice([1.0, -1.0, 0.5]), rng.choice([0.0, 1.0, -0.5])
            L.append(f"gemm {tA} {tB} {m} {n} {k} {alpha} {offA} {ma} {ldA} {offB} {mb} {ldB} {beta} {offC} {mc} {ldC}")
        elif op == "sp":
            mat = int(rng.randint(0, 3)); rw = rng.choice(["r", "w", "s"])
            nrow, ncol = int(rng.randint(1, 7)), int(rng.randint(1, 11))
            ld = 64 if mat == 0 else int(rng.choice([40, 48]))
            top = size[mat] - (ncol - 1) * ld            # a row starts below this offset
            starts = rng.permutation(min(top, ld))[:nrow]  # distinct rows inside the first column: no aliasing for w / s
            offs = " ".join(str(int(o)) for o in starts)
            L.append(f"sp {rw} {nrow} {ncol} {ld} {ncol + int(rng.randint(0, 3))} {mat} {s} {offs}"); s += 1
        else:
            L.append("wait")
    return "\n".join(L) + "\n"


def make_scripts() -> dict:
    return {
        "lu_ld96_b16": _lu_like(96, 16, 1, 100, ragged=False),
        "lu_ld128_b32_k2": _lu_like(128, 32, 2, 200, ragged=False),
        "lu_ld90_b18_ragged": _lu_like(90, 18, 1, 300, ragged=True),
        "edge": _edge_cases(400),
    }
