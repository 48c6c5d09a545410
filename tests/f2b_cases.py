"""Cases of tests/golden/f2b_ref_outputs.npz (levels of the unmodified reference's sym_full2band, see
tests/golden/make_golden_f2b.py) and the replay helper shared by the oracle test and the GPU / simulator worker.
Test infrastructure."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle_py as orc  # noqa: E402


def load_golden():
    return np.load(os.path.join(HERE, "golden", "f2b_ref_outputs.npz"))


def case_names(gold):
    return sorted(k[:-5] for k in gold.files if k.endswith(".args"))


def stored_levels(gold, name):
    return sorted({int(k.split(".")[1][1:]) for k in gold.files if k.startswith(name + ".L")})


def level_state(n, b, b_sub, pr, level):
    """(n_level, rrow, rcol, per-rank corner (row, col) inside the local array) on entry to `level` — the recurrences of
    full_to_band.cxx:57-66,90,245,247 applied `level` times from (n, 0, 0)."""
    rrow = rcol = 0
    corner = {(i, j): (0, 0) for i in range(pr) for j in range(pr)}
    nn = n
    for _ in range(level):
        for (i, j), (cr, cc) in list(corner.items()):
            ro, co, _, _ = orc.f2b_level(nn, b, b_sub, pr, rrow, rcol, i, j)
            corner[(i, j)] = (cr + ro, cc + co)
        rrow, rcol = (rrow + b // b_sub) % pr, (rcol + b // b_sub) % pr
        nn -= b
    return nn, rrow, rcol, corner
