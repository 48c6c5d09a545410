"""GPU worker for tests/test_zz_lu_offload_gpu.py — runs in its own process so that a CUDA fault in the (new) LU seam cannot
poison the test session.  Prints one JSON line.

    python tests/off_worker.py script <name> <overlap 0|1>     golden/oracle parity of one committed script
    python tests/off_worker.py trailing <n> <k> <overlap>      LU trailing-update pattern at size n against numpy
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

if os.environ.get("CANDMC_CPUSIM") == "1":   # CPU suite: the same worker on the functional simulator (tests/cpusim)
    sys.path.insert(0, os.path.join(HERE, "cpusim"))
    import simtorch
    simtorch.install()

from off_script import GpuBackend, OracleBackend, run_script  # noqa: E402


def do_script(name, overlap):
    gold = np.load(os.path.join(HERE, "golden", "lu_offload_ref_outputs.npz"))
    text = str(gold[f"{name}__script"])
    want = [gold[f"{name}__out{i}"] for i in range(int(gold[f"{name}__nout"]))]
    ob = OracleBackend()
    orc = run_script(text, ob)
    ob.close()
    be = GpuBackend(overlap=bool(overlap))
    got = run_script(text, be)
    from candmc_b200 import lu_offload as lo, launch_count
    st = lo.stats()
    be.close()
    res = {"name": name, "overlap": overlap, "records": len(got), "stats": st, "launches": int(launch_count())}
    worst_ref = worst_orc = 0.0
    pad_ok = True
    exact = total = 0
    assert len(got) == len(want) == len(orc)
    for g, w, o in zip(got, want, orc):
        g = g[: w.size]
        o = o[: w.size]
        pad_ok &= bool(np.array_equal(g[w == -7.0], w[w == -7.0]))
        worst_ref = max(worst_ref, float(np.abs(g - w).max()) if w.size else 0.0)
        worst_orc = max(worst_orc, float(np.abs(g - o).max()) if w.size else 0.0)
        exact += int((g == w).sum())
        total += w.size
    res.update(max_abs_vs_reference=worst_ref, max_abs_vs_oracle=worst_orc, padding_untouched=pad_ok,
               exact_fraction=exact / max(total, 1))
    print(json.dumps(res))


def do_fuzz(seed, overlap, count=6):
    """random scripts (tests/off_script.py random_script) on the device path vs the oracle: copies bit-exact, GEMM-touched
    elements to 1e-11 relative to the magnitudes involved"""
    from off_script import random_script
    from candmc_b200 import lu_offload as lo
    worst, waits, exact, total = 0.0, 0, 0, 0
    for i in range(count):
        text = random_script(seed * 100 + i)
        ob = OracleBackend(); want = run_script(text, ob); ob.close()
        be = GpuBackend(overlap=bool(overlap)); got = run_script(text, be)
        waits += lo.stats()["cross_stream_waits"]
        be.close()
        assert len(got) == len(want)
        for g, w in zip(got, want):
            g = g[: w.size]
            scale = max(1.0, float(np.abs(w).max()) if w.size else 1.0)
            worst = max(worst, float(np.abs(g - w).max()) / scale if w.size else 0.0)
            exact += int((g == w).sum()); total += w.size
    print(json.dumps({"seed": seed, "overlap": overlap, "scripts": count, "max_rel_vs_oracle": worst, "cross_stream_waits": waits,
                      "exact_fraction": exact / max(total, 1)}))


class OracleLo:
    """The reference-named API (candmc_b200.lu_offload's surface) on top of oracle_off_*, so that the CPU suite can run
    do_trailing's logic against the checker (tests/test_lu_offload_oracle.py)."""
    OFF_A, OFF_L, OFF_U = 0, 1, 2

    def __init__(self):
        self.be = OracleBackend()

    def set_overlap(self, on):
        pass

    def alloc_A(self, size, ptr=None):
        self.be.alloc(0, size)
        if ptr is not None:
            self.be.fill(0, np.ascontiguousarray(ptr))

    def alloc_L(self, size):
        self.be.alloc(1, size)

    def alloc_U(self, size):
        self.be.alloc(2, size)

    def alloc_transfer(self, size):
        pass

    def upload_lda_cpy(self, nrow, ncol, lda_A, lda_B, A, off_B, mat):
        self.be.upload(nrow, ncol, lda_A, lda_B, A, off_B, mat)

    def download_lda_cpy(self, nrow, ncol, lda_A, lda_B, off_A, B, mat):
        self.be.download(nrow, ncol, lda_A, lda_B, off_A, B, mat)

    def offload_gemm_A(self, *a):
        self.be.gemm(*a)

    def offload_sparse_rw(self, nrow, ncol, lda_B, A, lda_A, offs, mat, rw):
        self.be.sparse_rw(nrow, ncol, lda_B, A, lda_A, offs, mat, rw)

    def wait_gemm(self):
        pass

    def stats(self):
        return {"cross_stream_waits": 0}

    def free_offload_A(self):
        self.be.close()

    free_offload_L = free_offload_U = free_offload_transfer = lambda self: None


def do_trailing(n, k, overlap, lo=None):
    """A22 <- A22 - L21 * U12 on an n x n local matrix with panels of width k, next panel downloaded right behind the
    GEMM, an independent block downloaded concurrently; checked against numpy in float64."""
    if lo is None:
        from candmc_b200 import lu_offload as lo, launch_count
    else:
        launch_count = lambda: 0  # noqa: E731
    lo.set_overlap(bool(overlap))
    rng = np.random.RandomState(5)
    A = np.asfortranarray(rng.rand(n, n) - 0.5)
    Lp = np.asfortranarray(rng.rand(n - k, k) - 0.5)
    Up = np.asfortranarray(rng.rand(k, n - k) - 0.5)
    lo.alloc_A(n * n, A.reshape(-1, order="F"))
    lo.alloc_L((n - k) * k)
    lo.alloc_U(k * (n - k))
    lo.alloc_transfer(k * n)
    lo.upload_lda_cpy(n - k, k, n - k, n - k, Lp.reshape(-1, order="F"), 0, lo.OFF_L)
    lo.upload_lda_cpy(k, n - k, k, k, Up.reshape(-1, order="F"), 0, lo.OFF_U)
    off = k * n + k
    lo.offload_gemm_A("N", "N", n - k, n - k, k, -1.0, 0, lo.OFF_L, n - k, 0, lo.OFF_U, k, 1.0, off, lo.OFF_A, n)
    # (1) the first block column (not touched by the GEMM): may be downloaded while the GEMM runs
    first = np.zeros((n, k), order="F")
    lo.download_lda_cpy(n, k, n, n, 0, first.reshape(-1, order="F"), lo.OFF_A)
    waits_after_independent = lo.stats()["cross_stream_waits"]
    # (2) the next panel (GEMM output): must wait for the GEMM without an explicit wait_gemm
    panel = np.zeros((n - k, k), order="F")
    pv = panel.reshape(-1, order="F")
    lo.download_lda_cpy(n - k, k, n, n - k, off, pv, lo.OFF_A)
    # (3) pivot rows of the updated block: swap two rows through the host
    rows = np.array([k + 3, n - 2], dtype=np.int32)
    offs = rows + k * n
    buf = np.zeros(2 * (n - k))
    lo.offload_sparse_rw(2, n - k, n, buf, n - k, offs, lo.OFF_A, "r")
    # rows a, b are now in buf[0], buf[1]; writing them back with the offsets reversed swaps the two rows on the device
    lo.offload_sparse_rw(2, n - k, n, buf, n - k, offs[::-1].copy(), lo.OFF_A, "w")
    lo.wait_gemm()
    whole = np.zeros(n * n)
    lo.download_lda_cpy(n * n, 1, n * n, n * n, 0, whole, lo.OFF_A)
    st = lo.stats()
    launches = int(launch_count())
    lo.free_offload_A(); lo.free_offload_L(); lo.free_offload_U(); lo.free_offload_transfer()
    ref = A.copy(order="F")
    ref[k:, k:] -= Lp @ Up
    want_panel = ref[k:, k:2 * k].copy()
    want_rows = ref[rows, k:].copy()
    ref[rows[::-1], k:] = want_rows  # swapped
    got = whole.reshape(n, n, order="F")
    rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    res = {
        "n": n, "k": k, "overlap": overlap, "rel_frobenius": rel, "bound": 10 * k * np.finfo(np.float64).eps,
        "first_block_exact": bool(np.array_equal(first, A[:, :k])),
        "panel_rel": float(np.linalg.norm(pv.reshape(n - k, k, order="F") - want_panel) / np.linalg.norm(want_panel)),
        "rows_rel": float(np.linalg.norm(buf.reshape(2, n - k) - want_rows) / np.linalg.norm(want_rows)),
        "waits_after_independent_download": waits_after_independent, "stats": st, "launches": launches,
    }
    print(json.dumps(res))
    return res


if __name__ == "__main__" and sys.argv[1] == "fuzz":
    do_fuzz(int(sys.argv[2]), int(sys.argv[3]))
    sys.exit(0)

if __name__ == "__main__":
    if sys.argv[1] == "script":
        do_script(sys.argv[2], int(sys.argv[3]))
    else:
        do_trailing(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
