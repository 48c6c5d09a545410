"""CPU tests: the numpy restatement of the DMatrix pack operations (oracle/oracle_py.py dmat_*) against the outputs of the
UNMODIFIED reference (alg/SE/dmatrix.cxx run by oracle/ref_dmat_dump.cxx; tests/golden/dmat_ref_outputs.npz).  Pure data
movement is bit-exact; reduce_scatter_horizontal sums npcol terms in a different order than the reference's butterfly."""
import numpy as np
import pytest

from dmat_cases import build_case, case_names, load_golden, op_of

GOLD = load_golden()
EPS = 2.220446049250313e-16


@pytest.mark.parametrize("name", case_names(GOLD))
def test_oracle_matches_reference(name):
    case = build_case(op_of(name), GOLD[f"{name}.args"])
    for r in range(case["ranks"]):
        ref = GOLD[f"{name}.r{r}"]
        got = np.asarray(case["want"][r]).reshape(-1, order="F")
        assert got.shape == ref.shape, (name, r)
        if op_of(name) == "rsh":
            assert np.abs(got - ref).max() <= 4 * case["npcol"] * EPS, (name, r)
        else:
            assert np.array_equal(got, ref), (name, r)


def test_folds_are_inverse():
    from oracle import oracle_py as orc
    rng = np.random.RandomState(3)
    for (mr, mc, b, f) in [(24, 12, 2, 3), (16, 8, 4, 2), (12, 12, 1, 4)]:
        X = np.asfortranarray(rng.rand(mr, mc))
        assert np.array_equal(orc.dmat_foldrows(orc.dmat_foldcols(X, b, f), b, f), X)


def test_fold_kernel_index_map_matches_oracle():
    """the index function the CUDA fold kernels evaluate per element (host-callable), against the pinned numpy restatement"""
    import ctypes as C

    from candmc_b200 import lib
    from oracle import oracle_py as orc
    rng = np.random.RandomState(4)
    for (mr, mc, b, f, pad) in [(24, 12, 2, 3, 0), (16, 8, 4, 2, 3), (12, 12, 1, 4, 1), (30, 20, 3, 5, 2)]:
        lda = mr + pad
        X = np.zeros((lda, mc), order="F")
        X[:mr] = rng.rand(mr, mc)
        flat = X.reshape(-1, order="F")
        for foldcols, want in ((1, orc.dmat_foldcols(X[:mr], b, f) if mr % (b * f) == 0 else None),
                               (0, orc.dmat_foldrows(X[:mr], b, f) if mc % f == 0 else None)):
            if want is None:
                continue
            orow, ocol = want.shape
            got = np.empty_like(want)
            src = C.c_int64()
            for cc in range(ocol):
                for rr in range(orow):
                    assert lib().candmc_debug_fold_src_index(foldcols, mr, mc, b, f, lda, rr, cc, C.byref(src)) == 0
                    got[rr, cc] = flat[src.value]
            assert np.array_equal(got, want), (mr, mc, b, f, foldcols)
