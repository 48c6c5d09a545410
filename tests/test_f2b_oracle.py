"""The numpy restatement of the full -> band trailing update (oracle/oracle_py.py f2b_*) against the levels
dumped from the UNMODIFIED reference sym_full2band (alg/SE/full_to_band.cxx with the reference's own 2D QR; fixture made by
tests/golden/make_golden_f2b.py).  Bound: |diff| <= 64 * b * eps per element (|values| <= ~4, sums of b..n/pr terms,
different BLAS summation order); observed 4e-16."""
import numpy as np
import pytest

from f2b_cases import case_names, level_state, load_golden, stored_levels, orc

GOLD = load_golden()
EPS = 2.220446049250313e-16


@pytest.mark.parametrize("name", case_names(GOLD))
def test_oracle_reproduces_every_stored_level_of_the_reference_run(name):
    P, n, b, bs, levels = [int(x) for x in GOLD[f"{name}.args"]]
    pr = int(round(P ** 0.5))
    nl = n // pr
    assert stored_levels(GOLD, name)
    for L in stored_levels(GOLD, name):
        nn, rrow, rcol, corner = level_state(n, b, bs, pr, L)
        A = [GOLD[f"{name}.L{L}.r{r}.Ain"].reshape(nl, nl, order="F").copy() for r in range(P)]
        views, Ys = [], []
        for r in range(P):
            i, j = r % pr, r // pr
            _, _, mb, _ = orc.f2b_level(nn, b, bs, pr, rrow, rcol, i, j)
            cr, cc = corner[(i, j)]
            views.append(A[r][cr:, cc:])
            y = GOLD[f"{name}.L{L}.r{r}.Y"]
            assert y.size == mb * b
            Ys.append(y.reshape(mb, b, order="F"))
        orc.f2b_update(nn, b, bs, pr, rrow, rcol, views, Ys)
        for r in range(P):
            want = GOLD[f"{name}.L{L}.r{r}.Aout"].reshape(nl, nl, order="F")
            assert np.abs(A[r] - want).max() <= 64 * b * EPS
            # everything outside the trailing block is untouched by the update
            i, j = r % pr, r // pr
            ro, co, mb, kb = orc.f2b_level(nn, b, bs, pr, rrow, rcol, i, j)
            cr, cc = corner[(i, j)]
            mask = np.ones((nl, nl), bool)
            mask[cr + ro:cr + ro + mb, cc + co:cc + co + kb] = False
            assert np.array_equal(A[r][mask], GOLD[f"{name}.L{L}.r{r}.Ain"].reshape(nl, nl, order="F")[mask])


def test_extent_formulas_keep_the_c_remainder_semantics():
    """full_to_band.cxx:70,77 subtract before taking %; C truncates towards zero, Python floors"""
    assert orc._cmod(-1, 3) == -1 and (-1) % 3 == 2
    ro, co, mb, kb = orc.f2b_level(48, 8, 4, 2, 0, 0, 1, 0)
    assert (ro, co, mb, kb) == (4, 4, 20, 20)
