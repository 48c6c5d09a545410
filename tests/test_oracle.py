"""CPU tests: oracle/candmc_oracle.c (our restatement) against outputs of the UNMODIFIED reference
(tests/golden/canmm_ref_outputs.npz, made by tests/golden/make_golden.py) and against numpy."""
import numpy as np
import pytest

from oracle import oracle_py as orc

EPS = 2.220446049250313e-16


def rel_frob(x, ref):
    x = np.asarray(x, dtype=np.float64).ravel(order="F")  # column-major, like the reference's raw buffers
    ref = np.asarray(ref, dtype=np.float64).ravel(order="F")
    return np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300)


def zeros_like_blocks(blocks, shape=None):
    return [np.zeros(shape or b.shape, dtype=np.float64, order="F") for b in blocks]


def test_drand48_generator_matches_libc():
    import ctypes

    libc = ctypes.CDLL("libc.so.6")
    libc.drand48.restype = ctypes.c_double
    libc.srand48.argtypes = [ctypes.c_long]
    for seed in (0, 3, 1000, 12345678901):
        libc.srand48(seed)
        want = [libc.drand48() for _ in range(5)]
        got = orc.drand48_stream(seed, 5)
        assert list(got) == want
    n = 96
    for (r, c) in [(0, 0), (5, 7), (95, 95)]:
        libc.srand48(c * n + r)
        a, b = libc.drand48(), libc.drand48()
        assert orc.lib().oracle_unit_elem(r, c, n, 0) == a
        assert orc.lib().oracle_unit_elem(r, c, n, 1) == b


@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
def test_oracle_dgemm_vs_numpy(ta, tb):
    rng = np.random.default_rng(7)
    m, n, k = 37, 29, 53
    A = np.asfortranarray(rng.random((m, k) if ta == "N" else (k, m)))
    B = np.asfortranarray(rng.random((k, n) if tb == "N" else (n, k)))
    Cm = np.asfortranarray(rng.random((m + 3, n)))
    want = 1.2 * (A if ta == "N" else A.T) @ (B if tb == "N" else B.T) + 0.8 * Cm[:m]
    orc.dgemm(ta, tb, m, n, k, 1.2, A, A.shape[0], B, B.shape[0], 0.8, Cm, m + 3)
    assert rel_frob(Cm[:m], want) < 10 * k * EPS
    # beta == 0 must overwrite NaNs (BLAS semantics)
    Cn = np.full((m, n), np.nan, order="F")
    orc.dgemm(ta, tb, m, n, k, 1.0, A, A.shape[0], B, B.shape[0], 0.0, Cn, m)
    assert np.isfinite(Cn).all()


def test_oracle_pack_kernels():
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.random((11, 7)))
    B = np.zeros((9, 7), order="F")
    orc.lda_cpy(5, 7, 11, 9, A, B)
    assert (B[:5] == A[:5]).all() and (B[5:] == 0).all()
    B2 = np.asfortranarray(rng.random((9, 7)))
    want = B2.copy()
    want[:5] = want[:5] * 0.25 + A[:5] * 0.5
    orc.lda_cpy(5, 7, 11, 9, A, B2, 0.5, 0.25)
    assert np.array_equal(B2, want)
    T = np.zeros((7, 11), order="F")
    orc.transpose(11, 7, A, 11, T, 7)
    assert np.array_equal(T, A.T)


D25 = [("d25_n96_q2_c1_ovp0", 96, 2, 1, 0), ("d25_n96_q2_c1_ovp1", 96, 2, 1, 1), ("d25_n64_q2_c2_ovp0", 64, 2, 2, 0),
       ("d25_n64_q2_c2_ovp1", 64, 2, 2, 1), ("d25_n40_q1_c1_ovp0", 40, 1, 1, 0), ("d25_n96_q4_c2_ovp0", 96, 4, 2, 0),
       ("d25_n90_q3_c1_ovp1", 90, 3, 1, 1)]


@pytest.mark.parametrize("name,n,q,c,ovp", D25)
def test_oracle_d25_matches_reference(golden, name, n, q, c, ovp):
    A, B = orc.d25_blocks(n, q, c)
    Cb = zeros_like_blocks(A)
    orc.d25_summa(n, q, c, ovp, A, B, Cb)
    ref = golden[name]
    for r in range(q * q * c):
        assert rel_frob(Cb[r], ref[r]) <= 10 * n * EPS, (name, r)


@pytest.mark.parametrize("name,n,q", [("summa_n64_q2", 64, 2), ("summa_n96_q3", 96, 3)])
def test_oracle_summa_matches_reference(golden, name, n, q):
    A, B = orc.d25_blocks(n, q, 1)
    Cb = zeros_like_blocks(A)
    orc.summa(n, q, A, B, Cb)
    for r in range(q * q):
        assert rel_frob(Cb[r], golden[name][r]) <= 10 * n * EPS


@pytest.mark.parametrize("ovp", [0, 1])
def test_oracle_dcn_matches_reference(golden, ovp):
    n, x1, x2 = 64, 2, 1
    A, B = orc.dcn_blocks(n, x1, x2)
    Cb = zeros_like_blocks(A)
    orc.bcast_cannon_4d(n, x1, x2, ovp, A, B, Cb)
    for r in range(4):
        assert rel_frob(Cb[r], golden[f"dcn_n64_x2_1_ovp{ovp}"][r]) <= 10 * n * EPS


TRANS = [("summa_n64_q2_TN", "summa", 64, 2, 1, 0, "T", "N"), ("summa_n64_q2_NT", "summa", 64, 2, 1, 0, "N", "T"),
         ("d25_n64_q2_c2_ovp0_TT", "d25", 64, 2, 2, 0, "T", "T"), ("d25_n96_q2_c1_ovp1_TN", "d25", 96, 2, 1, 1, "T", "N"),
         ("dcn_n64_x2_1_ovp0_TN", "dcn", 64, 2, 1, 0, "T", "N"), ("dcn_n64_x2_1_ovp1_NT", "dcn", 64, 2, 1, 1, "N", "T"),
         ("dcn_n64_x2_1_ovp0_TT", "dcn", 64, 2, 1, 0, "T", "T")]


@pytest.mark.parametrize("name,kind,n,q,c,ovp,tA,tB", TRANS)
def test_oracle_trans_flags_match_reference(golden, name, kind, n, q, c, ovp, tA, tB):
    """trans_A / trans_B in the unmodified reference reach the local dgemm only (summa.cxx:97, d25_summa.cxx:185,
    dual_cannon.cxx:163-166): blocks travel as stored, every block product is op(A block) * op(B block)."""
    if kind == "dcn":
        A, B = orc.dcn_blocks(n, q, 1)
        Cb = zeros_like_blocks(A)
        orc.bcast_cannon_4d(n, q, 1, ovp, A, B, Cb, trans_A=tA, trans_B=tB)
    else:
        A, B = orc.d25_blocks(n, q, c)
        Cb = zeros_like_blocks(A)
        if kind == "summa":
            orc.summa(n, q, A, B, Cb, trans_A=tA, trans_B=tB)
        else:
            orc.d25_summa(n, q, c, ovp, A, B, Cb, tA, tB)
    for r in range(len(A)):
        assert rel_frob(Cb[r], golden[name][r]) <= 10 * n * EPS, (name, r)
    # ... which is NOT the product of the assembled transposes: the flagged result differs from the plain one
    plain = golden[{"summa": "summa_n64_q2", "dcn": f"dcn_n64_x2_1_ovp{ovp}"}.get(kind, name[:-3])]
    assert rel_frob(golden[name][0], plain[0]) > 1e-3


@pytest.mark.parametrize("x1,x2,n", [(1, 2, 32), (2, 2, 64), (1, 3, 48)])
def test_oracle_dcn_cannon_level_vs_serial(x1, x2, n):
    """x2_np > 1 cannot be run in the reference (it deadlocks); the expected output is by definition the serial product
    in the dcn_unit layout (test/MM/topo_pdgemm_unit.cxx:139-160)."""
    A, B = orc.dcn_blocks(n, x1, x2)
    Cb = zeros_like_blocks(A)
    orc.bcast_cannon_4d(n, x1, x2, 0, A, B, Cb)
    fullA = orc.unit_block(n, n, 0, 0, n, 0)
    fullB = orc.unit_block(n, n, 0, 0, n, 1)
    full = fullA @ fullB
    b = n // (x1 * x2)
    for r in range(len(A)):
        xa, ya = r % x1, (r // x1) % x1
        xb, yb = (r // (x1 * x1)) % x2, r // (x1 * x1 * x2)
        row0, col0 = (ya * x2 + yb) * b, (xa * x2 + xb) * b
        assert rel_frob(Cb[r], full[row0:row0 + b, col0:col0 + b]) <= 10 * n * EPS


SPC = [("spc_bidir1_p4_m24_k16_n20_N", 1, 2, 2, 20, 24, 16, "N"), ("spc_bidir0_p4_m24_k16_n20_N", 0, 2, 2, 20, 24, 16, "N"),
       ("spc_bidir1_p4_m24_k16_n20_T", 1, 2, 2, 20, 24, 16, "T"), ("spc_bidir1_p9_m16_k12_n8_N", 1, 3, 2, 8, 16, 12, "N"),
       ("spc_bidir0_p9_m16_k12_n8_N", 0, 3, 2, 8, 16, 12, "N"),
       ("spc_bidir1_p16_ndim4_m8_k16_n8_N", 1, 2, 4, 8, 8, 16, "N"),
       ("spc_bidir0_p16_ndim4_m8_k16_n8_N", 0, 2, 4, 8, 8, 16, "N")]


@pytest.mark.parametrize("name,bidir,kary,ndim,n,m,k,tB", SPC)
def test_oracle_spcannon_matches_reference(golden, name, bidir, kary, ndim, n, m, k, tB):
    A, B, Cb, (fA, fB, fC) = orc.spc_blocks(kary, ndim, 3, n, m, k, tB)
    orc.spcannon(bidir, kary, ndim, n, m, k, "N", 1.2, A, tB, 0.8, B, Cb)
    ref = golden[name]
    khalf = kary ** (ndim // 2)
    full = 1.2 * fA @ fB + 0.8 * fC
    for r in range(kary ** ndim):
        assert rel_frob(Cb[r], ref[r]) <= 10 * k * khalf * EPS, (name, r)
    # and the reference test's own criterion: serial product, |diff| <= 1e-6 (test/MM/test_spc.cxx:116-126)
    A, B, Cb, _ = orc.spc_blocks(kary, ndim, 3, n, m, k, tB)
    orc.spcannon(bidir, kary, ndim, n, m, k, "N", 1.2, A, tB, 0.8, B, Cb)
    for rank in range(kary ** ndim):
        px = py = 0
        sc, tr = 1, rank
        for _ in range(ndim // 2):
            px += (tr % kary) * sc; tr //= kary
            py += (tr % kary) * sc; tr //= kary
            sc *= kary
        assert np.abs(Cb[rank] - full[py * m:(py + 1) * m, px * n:(px + 1) * n]).max() <= 1e-6


@pytest.mark.parametrize("tA,tB", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")])
@pytest.mark.parametrize("c", [2, 3])
def test_oracle_d25_ksplit_extension_vs_serial(tA, tB, c):
    n = 48
    A = [np.asfortranarray(orc.unit_block(n, n, 0, 0, n, 0)) for _ in range(c)]
    B = [np.asfortranarray(orc.unit_block(n, n, 0, 0, n, 1)) for _ in range(c)]
    Cb = zeros_like_blocks(A)
    orc.d25_summa(n, 1, c, 0, A, B, Cb, tA, tB)
    want = (A[0].T if tA == "T" else A[0]) @ (B[0].T if tB == "T" else B[0])
    for r in range(c):
        assert rel_frob(Cb[r], want) <= 10 * n * EPS


def test_oracle_upd_A_vs_numpy():
    rng = np.random.default_rng(11)
    b, kb, mbs = 8, 12, [20, 16]
    T = np.asfortranarray(np.eye(b) + 0.01 * np.tril(rng.random((b, b))))
    Y = [np.asfortranarray(rng.random((mb, b))) for mb in mbs]
    A = [np.asfortranarray(rng.random((mb, kb))) for mb in mbs]
    Yf, Af = np.vstack(Y), np.vstack(A)
    want = Af - Yf @ np.linalg.solve(T, Yf.T @ Af)
    orc.upd_A(mbs, kb, b, Y, mbs, A, mbs, T)
    assert rel_frob(np.vstack(A), want) <= 1e-13


UPDA = [("upda_m96_k64_b8_2x2_r00", 96, 64, 8, 2, 2, 0, 0), ("upda_m80_k48_b8_2x3_r12", 80, 48, 8, 2, 3, 1, 2),
        ("upda_m64_k32_b16_1x1", 64, 32, 16, 1, 1, 0, 0), ("upda_m72_k40_b8_4x1_r20", 72, 40, 8, 4, 1, 2, 0)]


@pytest.mark.parametrize("name,m,k,b,nprow,npcol,rrow,rcol", UPDA)
def test_oracle_update_A_matches_reference(golden, name, m, k, b, nprow, npcol, rrow, rcol):
    """the reference's own update_A (alg/QR/qr_2d/qr_2d.cxx:124-177, W == NULL) on a block-cyclic grid with a
    rotated root; the fixture holds every rank's updated trailing block."""
    Y, A = orc.update_A_blocks(nprow, npcol, rrow, rcol, m, k, b)
    orc.update_A(nprow, npcol, rrow, rcol, m, k, b, Y, A)
    for r in range(nprow * npcol):
        ref = golden[f"{name}.r{r}"]
        assert A[r].size == ref.size, (name, r)
        if ref.size:
            assert rel_frob(A[r], ref) <= 10 * m * EPS, (name, r)


UPDW = [("updw_m96_k64_b8_2x2_r00", 96, 64, 8, 2, 2, 0, 0), ("updw_m96_k64_b8_2x2_r11", 96, 64, 8, 2, 2, 1, 1),
        ("updw_m80_k48_b8_2x3_r00", 80, 48, 8, 2, 3, 0, 0), ("updw_m64_k32_b16_1x1", 64, 32, 16, 1, 1, 0, 0),
        ("updw_m72_k40_b8_4x1_r20", 72, 40, 8, 4, 1, 2, 0), ("updw_m48_k72_b8_1x3_r02", 48, 72, 8, 1, 3, 0, 2)]


@pytest.mark.parametrize("name,m,k,b,nprow,npcol,rrow,rcol", UPDW)
def test_oracle_update_A_with_panel_W_matches_reference(golden, name, m, k, b, nprow, npcol, rrow, rcol):
    """SURVEY §8f N1, the form QR_2D itself uses (qr_2d.cxx:325): W is the panel QR's upper-triangular factor on the root rank,
    W_is_T == false, T = lower(-W^-T Y1) by comp_bcast_T_from_W (:179-208).  Fixture: the reference's own update_A."""
    Y, A = orc.update_A_blocks(nprow, npcol, rrow, rcol, m, k, b)
    orc.update_A(nprow, npcol, rrow, rcol, m, k, b, Y, A, orc.panel_W(b), W_is_T=False)
    for r in range(nprow * npcol):
        ref = golden[f"{name}.r{r}"]
        assert A[r].size == ref.size, (name, r)
        if ref.size:
            assert rel_frob(A[r], ref) <= 10 * m * EPS, (name, r)


UPDY = [("updy_m96_k64_b8_2x2_r00", 96, 64, 8, 2, 2, 0, 0), ("updy_m80_k48_b8_2x3_r12", 80, 48, 8, 2, 3, 1, 2),
        ("updy_m64_k32_b16_1x1", 64, 32, 16, 1, 1, 0, 0), ("updy_m72_k40_b8_4x1_r20", 72, 40, 8, 4, 1, 2, 0)]


@pytest.mark.parametrize("name,m,k,b,nprow,npcol,rrow,rcol", UPDY)
def test_oracle_update_Yamamoto_A_matches_reference(golden, name, m, k, b, nprow, npcol, rrow, rcol):
    """the reference's own update_Yamamoto_A (alg/QR/qr_2d/qr_y2d.cxx:68-120, agg == NULL); in the fixture's
    run only the root column holds the panel and T, the other columns must receive both."""
    Qm, A = orc.update_A_blocks(nprow, npcol, rrow, rcol, m, k, b)
    orc.update_Yamamoto_A(nprow, npcol, rrow, rcol, m, k, b, Qm, A, orc.yamamoto_T(b))
    for r in range(nprow * npcol):
        ref = golden[f"{name}.r{r}"]
        assert A[r].size == ref.size, (name, r)
        if ref.size:
            assert rel_frob(A[r], ref) <= 10 * m * EPS, (name, r)


UPDYAGG = [("updyagg_m96_k32_b8_2x2_r00", 96, 32, 8, 2, 2, 0, 0), ("updyagg_m96_k32_b8_2x2_r11", 96, 32, 8, 2, 2, 1, 1),
           ("updyagg_m80_k24_b8_2x3_r12", 80, 24, 8, 2, 3, 1, 2), ("updyagg_m64_k32_b16_1x1", 64, 32, 16, 1, 1, 0, 0),
           ("updyagg_m72_k24_b8_4x1_r20", 72, 24, 8, 4, 1, 2, 0)]


def split_updyagg(flat, mb0, kb0, k):
    """one rank's fixture of ref_dump's `updyagg` mode: A (mb0 x kb0) | aQm (max(mb0, 1) x k) | aT (k x k), column-major"""
    na, nq = mb0 * kb0, max(mb0, 1) * k
    assert flat.size == na + nq + k * k
    A = flat[:na].reshape(kb0, mb0).T if na else np.zeros((mb0, kb0))
    return A, flat[na:na + nq].reshape(k, max(mb0, 1)).T[:mb0], flat[na + nq:].reshape(k, k).T


@pytest.mark.parametrize("name,m,k,b,nprow,npcol,rrow,rcol", UPDYAGG)
def test_oracle_yamamoto_aggregator_matches_reference(golden, name, m, k, b, nprow, npcol, rrow, rcol):
    """SURVEY §8f N1, the aggregated form: the reference's own update_Yamamoto_A with an aggregator (aggregator::append,
    alg/QR/qr_2d/qr_y2d.cxx:38-62) over the k/b panels of a block column, driven as QR_Yamamoto_2D drives it (:171-277):
    the trailing updates, the aggregated panels aQm and the aggregated T on every rank."""
    A, aQm, aT = orc.yamamoto_aggregate(m, k, b)
    for r in range(nprow * npcol):
        myrow, mycol = r % nprow, r // nprow
        Al = orc.cyclic_local(A, b, nprow, npcol, rrow, rcol, myrow, mycol)
        Ql = orc.cyclic_local(aQm, b, nprow, 1, rrow, 0, myrow, 0)   # every grid column holds its grid row's rows of every panel
        gA, gQ, gT = split_updyagg(golden[f"{name}.r{r}"], Al.shape[0], Al.shape[1], k)
        assert np.abs(gA - Al).max() <= 10 * m * EPS and np.abs(gT - aT).max() <= 10 * m * EPS, (name, r)
        assert np.array_equal(gQ, Ql), (name, r)
