#!/usr/bin/env python
"""Generate tests/golden/f2b_ref_outputs.npz from the UNMODIFIED reference sym_full2band (alg/SE/full_to_band.cxx) run
under the mini-MPI by oracle/_ref/ref_f2b_dump (our tap around the reference's own 2D QR, see that file).  For every level
of the reduction: each rank's whole local array after the panel QR (Ain), the aggregated Householder panel (Y) and the whole
local array after the trailing update (Aout).  Also checks, per case, that the reference run really is a similarity
transform (eigenvalues of the final band == eigenvalues of the generated matrix).  Only runnable where /root/reference
exists; the fixture is committed.

    python tests/golden/make_golden_f2b.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
# name -> (ranks, n, b, b_sub).  b / b_sub and (n - b) / b_sub are multiples of the grid dimension: outside that the reference
# itself fails (heap corruption / segfault observed on 2x2 with b/b_sub = 3 and on 3x3 with b/b_sub = 2)
CASES = {
    "f2b_p1_n24_b8_s4": (1, 24, 8, 4),
    "f2b_p4_n48_b8_s4": (4, 48, 8, 4),
    "f2b_p4_n40_b8_s2": (4, 40, 8, 2),
    "f2b_p9_n72_b12_s4": (9, 72, 12, 4),
    "f2b_p16_n64_b16_s4": (16, 64, 16, 4),
}


def main():
    exe = os.path.join(REFDIR, "ref_f2b_dump")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/mpirun", "_ref/ref_f2b_dump"])
    out = {}
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    for name, (P, n, b, bs) in CASES.items():
        pr = int(round(P ** 0.5))
        nl = n // pr
        with tempfile.TemporaryDirectory() as td:
            subprocess.check_call([os.path.join(REFDIR, "mpirun"), "-np", str(P), "-timeout", "120", exe, str(n), str(b), str(bs),
                                   os.path.join(td, "x")], env=env, stdout=subprocess.DEVNULL)
            levels = (n - 1) // b if n % b else n // b - 1
            out[f"{name}.args"] = np.array([P, n, b, bs, levels])
            load = lambda L, r, w: np.fromfile(os.path.join(td, f"x.L{L}.r{r}.{w}"))  # noqa: E731
            final = [load(levels - 1, r, "Aout") for r in range(P)]
            for L in sorted({0, 1, levels - 1} & set(range(levels))):   # first two and the last level keep the fixture small
                for r in range(P):
                    for w in ("Ain", "Y", "Aout"):
                        out[f"{name}.L{L}.r{r}.{w}"] = load(L, r, w)
            full = np.zeros((n, n))
            for r in range(P):
                i, j = r % pr, r // pr
                Af = final[r].reshape(nl, nl, order="F")
                gr = ((np.arange(nl) // bs) * pr + i) * bs + np.arange(nl) % bs
                gc = ((np.arange(nl) // bs) * pr + j) * bs + np.arange(nl) % bs
                full[np.ix_(gr, gc)] = Af
            I, J = np.indices((n, n))
            band = np.tril(np.where(np.abs(I - J) <= b, full, 0.0))
            band = band + np.tril(band, -1).T
            d = np.abs(np.linalg.eigvalsh(orc.f2b_sym_value(n)) - np.linalg.eigvalsh(band)).max()
            assert d < 1e-12 * n, (name, d)
            print(f"{name}: {levels} levels, eigenvalues preserved to {d:.1e}")
    path = os.path.join(ROOT, "tests", "golden", "f2b_ref_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
