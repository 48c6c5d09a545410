"""Inputs and oracle results of the DMatrix pack-operation cases of tests/golden/dmat_ref_outputs.npz (the same cases and
element seeds oracle/ref_dmat_dump.cxx feeds to the unmodified reference).  Test infrastructure."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from off_script import off_value  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402


def load_golden():
    return np.load(os.path.join(HERE, "golden", "dmat_ref_outputs.npz"))


def case_names(gold):
    return sorted(k[:-5] for k in gold.files if k.endswith(".args"))


def build_case(op, args):
    """args = [ranks, nrow, ncol, b, nprow, rrow, rcol, factor, sliced] -> dict with per-rank parents (the allocated local
    pieces), slice parameters, contributions (rsh) and the oracle's expected outputs."""
    ranks, nrow, ncol, b, nprow, rrow, rcol, factor, sliced = [int(x) for x in args]
    npcol = ranks // nprow
    table = off_value(11, nrow * ncol)
    value = lambda gr, gc: table[gr + gc * nrow]  # noqa: E731
    parents = [orc.dmat_local(nrow, ncol, b, nprow, npcol, rrow, rcol, r % nprow, r // nprow, value) for r in range(ranks)]
    if sliced:
        fr, fc = nprow * b, npcol * b
        sl = [orc.dmat_slice(parents[r], nrow, ncol, b, nprow, npcol, rrow, rcol, r % nprow, r // nprow, fr, nrow - fr, fc,
                             ncol - fc) for r in range(ranks)]
        pieces = [np.asfortranarray(s[0]) for s in sl]
        xr, xc = sl[0][1], sl[0][2]
        xn, xm = nrow - fr, ncol - fc
        # independent statement of what a slice is: the same global elements, shifted
        for r in range(ranks):
            direct = orc.dmat_local(xn, xm, b, nprow, npcol, xr, xc, r % nprow, r // nprow,
                                    lambda gr, gc: table[(gr + fr) + (gc + fc) * nrow])
            assert np.array_equal(direct, pieces[r])
    else:
        pieces, xr, xc, xn, xm = parents, rrow, rcol, nrow, ncol
    cntrbs = None
    if op == "repv":
        want = orc.dmat_replicate_vertical(pieces, nprow, npcol)
    elif op == "reph":
        want = orc.dmat_replicate_horizontal(pieces, nprow, npcol)
    elif op == "rsh":
        cntrbs = [off_value(100 + r, xm * pieces[r].shape[0]) for r in range(ranks)]
        want = orc.dmat_reduce_scatter_horizontal(pieces, cntrbs, nprow, npcol)
    elif op == "tpd":
        want = orc.dmat_transpose_data(pieces, nprow, npcol)
    elif op == "fc":
        want = [orc.dmat_foldcols(p, b, factor).reshape(-1, order="F") for p in pieces]
    elif op == "fr":
        want = [orc.dmat_foldrows(p, b, factor).reshape(-1, order="F") for p in pieces]
    else:
        raise ValueError(op)
    return dict(ranks=ranks, nrow=nrow, ncol=ncol, b=b, nprow=nprow, npcol=npcol, rrow=rrow, rcol=rcol, factor=factor,
                sliced=sliced, parents=parents, pieces=pieces, cntrbs=cntrbs, want=want, xnrow=xn, xncol=xm)


def op_of(name):
    return name.split("_")[0]
