"""GPU parity tests of the local kernels, through the C ABI (ctypes), against the CPU oracle.
FP64 GEMM tolerance: relative Frobenius <= 10*k*eps (BASELINE.json north_star: 10*n*eps); pack kernels and the drand48
generator are bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # collected but skipped on the CPU box
    pytest.skip("needs a B200", allow_module_level=True)

import candmc_b200 as cb  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402

EPS = 2.220446049250313e-16


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).cuda()


def host(t, rows, cols):
    return t.cpu().numpy().reshape(cols, -1).T[:rows]


def rel_frob(x, ref):
    x = np.asarray(x).ravel(order="F"); ref = np.asarray(ref).ravel(order="F")
    return np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300)


SHAPES = [(128, 128, 16), (128, 128, 128), (256, 384, 64), (1, 1, 1), (7, 5, 3), (130, 70, 33), (64, 200, 100),
          (257, 129, 17), (512, 96, 40), (31, 1000, 8), (1000, 31, 9), (333, 222, 111), (640, 512, 300),
          (128, 128, 4096), (256, 250, 3000), (130, 5, 2049)]   # the last three take the split-K path


@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
@pytest.mark.parametrize("pad,force_generic", [(0, False), (2, False), (1, False), (0, True)])
def test_dgemm_vs_oracle(ta, tb, pad, force_generic):
    rng = np.random.default_rng(1)
    cb.init()
    before = cb.launch_count()
    for (m, n, k) in SHAPES:
        for alpha, beta in ((1.0, 0.0), (1.2, 0.8), (-1.0, 1.0)):
            ra, ca = (m, k) if ta == "N" else (k, m)
            rb, cbk = (k, n) if tb == "N" else (n, k)
            A = np.zeros((ra + pad, ca), order="F"); A[:ra] = rng.random((ra, ca)) - 0.5
            B = np.zeros((rb + pad, cbk), order="F"); B[:rb] = rng.random((rb, cbk)) - 0.5
            Cm = np.asfortranarray(rng.random((m + pad, n)) - 0.5)
            if beta == 0.0:
                Cm[:] = np.nan  # beta == 0 must not read C
            want = Cm.copy(order="F")
            orc.dgemm(ta, tb, m, n, k, alpha, A, ra + pad, B, rb + pad, beta, want, m + pad)
            dA, dB, dC = dev(A), dev(B), dev(Cm)
            cb.lib().candmc_debug_force_generic_gemm(int(force_generic))
            try:
                cb.cdgemm(ta, tb, m, n, k, alpha, dA, ra + pad, dB, rb + pad, beta, dC, m + pad)
            finally:
                cb.lib().candmc_debug_force_generic_gemm(0)
            torch.cuda.synchronize()
            got = host(dC, m + pad, n)
            assert rel_frob(got[:m], want[:m]) <= 10 * max(k, 1) * EPS, (ta, tb, m, n, k, alpha, beta, pad)
            if pad and beta != 0.0:  # rows beyond m belong to the caller
                assert np.array_equal(got[m:], Cm[m:])
    assert cb.launch_count() > before


def test_dgemm_degenerate_and_errors():
    cb.init()
    Cm = np.asfortranarray(np.arange(12.0).reshape(3, 4))
    dC = dev(Cm)
    A = dev(np.zeros((3, 1), order="F")); B = dev(np.zeros((1, 4), order="F"))
    cb.cdgemm("N", "N", 3, 4, 0, 1.0, A, 3, B, 1, 2.0, dC, 3)   # k = 0 -> C = beta*C
    torch.cuda.synchronize()
    assert np.array_equal(host(dC, 3, 4), 2.0 * Cm)
    cb.cdgemm("N", "N", 0, 4, 5, 1.0, A, 1, B, 5, 0.0, dC, 1)   # m = 0 -> no-op
    with pytest.raises(cb.CandmcError):
        cb.cdgemm("X", "N", 3, 4, 1, 1.0, A, 3, B, 1, 0.0, dC, 3)
    with pytest.raises(cb.CandmcError):
        cb.cdgemm("N", "N", 3, 4, 1, 1.0, A, 2, B, 1, 0.0, dC, 3)  # lda < m


def test_dgemm_host_pointers():
    """The reference's callers own host matrices; the C ABI stages them."""
    rng = np.random.default_rng(2)
    m, n, k = 200, 150, 70
    A = np.asfortranarray(rng.random((m + 3, k))); B = np.asfortranarray(rng.random((k, n)))
    Cm = np.asfortranarray(rng.random((m, n))); want = Cm.copy(order="F")
    orc.dgemm("N", "N", m, n, k, 0.5, A, m + 3, B, k, 2.0, want, m)
    cb.cdgemm("N", "N", m, n, k, 0.5, A, m + 3, B, k, 2.0, Cm, m)
    assert rel_frob(Cm, want) <= 10 * k * EPS


@pytest.mark.parametrize("nrow,ncol,lda,ldb", [(64, 32, 64, 64), (48, 20, 50, 60), (7, 5, 9, 8), (1000, 333, 1024, 1000),
                                               (33, 17, 33, 40)])
def test_pack_kernels_bit_exact(nrow, ncol, lda, ldb):
    rng = np.random.default_rng(4)
    A = np.asfortranarray(rng.random((lda, ncol))); B = np.asfortranarray(rng.random((ldb, ncol)))
    want = B.copy(order="F"); orc.lda_cpy(nrow, ncol, lda, ldb, A, want)
    dB = dev(B); cb.lda_cpy(nrow, ncol, lda, ldb, dev(A), dB); torch.cuda.synchronize()
    assert np.array_equal(host(dB, ldb, ncol), want)
    want = B.copy(order="F"); orc.lda_cpy(nrow, ncol, lda, ldb, A, want, 1.5, -0.25)
    dB = dev(B); cb.lda_cpy(nrow, ncol, lda, ldb, dev(A), dB, 1.5, -0.25); torch.cuda.synchronize()
    # b*B + a*A may contract to an fma on the GPU: identical up to one rounding of the larger product
    got = host(dB, ldb, ncol)
    bound = 4 * EPS * (1.5 * np.abs(A[:nrow]) + 0.25 * np.abs(B[:nrow]))
    assert (np.abs(got[:nrow] - want[:nrow]) <= bound).all() and np.array_equal(got[nrow:], want[nrow:])
    T = np.zeros((ncol + 1, nrow), order="F"); wantT = T.copy(order="F")
    orc.transpose(nrow, ncol, A, lda, wantT, ncol + 1)
    dT = dev(T); cb.transpose(nrow, ncol, dev(A), lda, dT, ncol + 1); torch.cuda.synchronize()
    assert np.array_equal(host(dT, ncol + 1, nrow), wantT)
    # host-side operands go through the same entry points
    Bh = B.copy(order="F"); cb.lda_cpy(nrow, ncol, lda, ldb, A, Bh)
    want = B.copy(order="F"); orc.lda_cpy(nrow, ncol, lda, ldb, A, want)
    assert np.array_equal(Bh, want)


def test_drand48_generator_bit_exact():
    n, b = 96, 48
    for which in (0, 1):
        X = torch.zeros(b * b, dtype=torch.float64, device="cuda")
        cb.fill_drand48(X, b, b, b, 48, 0, n, which); torch.cuda.synchronize()
        assert np.array_equal(host(X, b, b), orc.unit_block(b, b, 48, 0, n, which))


def test_splitk_matches_plain_kernel_and_is_deterministic():
    """Small tile counts are cut along k (partials summed in a fixed order): same result to rounding as the unsplit kernel,
    bit-identical from run to run."""
    m, n, k = 512, 384, 8192
    A = torch.empty(m * k, dtype=torch.float64, device="cuda"); B = torch.empty(k * n, dtype=torch.float64, device="cuda")
    cb.fill_drand48(A, m, k, m, 0, 0, m, 0); cb.fill_drand48(B, k, n, k, 0, 0, k, 1)
    outs = []
    for on in (1, 1, 0):
        cb.lib().candmc_debug_splitk(on)
        Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
        cb.cdgemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, Cm, m)
        torch.cuda.synchronize()
        outs.append(Cm.cpu().numpy())
    cb.lib().candmc_debug_splitk(1)
    assert np.array_equal(outs[0], outs[1])
    assert rel_frob(outs[0], outs[2]) <= 10 * k * EPS


def test_splitk_launches_on_two_streams_do_not_mix_their_scratch():
    """The split-K partial tiles and arrival counters are one buffer per process: a split-K GEMM enqueued on a second stream
    while another one is still in flight must wait for it (runtime.cu: splitk_buffers / splitk_release) — each result is
    bit-identical to the same multiply run alone."""
    m, n, k = 512, 384, 16384
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ops = []
    for seed in (0, 1):
        A = torch.empty(m * k, dtype=torch.float64, device="cuda"); B = torch.empty(k * n, dtype=torch.float64, device="cuda")
        cb.fill_drand48(A, m, k, m, seed * 7, 0, m + 13 * seed, 0); cb.fill_drand48(B, k, n, k, 0, seed * 5, k + seed, 1)
        alone = torch.zeros(m * n, dtype=torch.float64, device="cuda")
        cb.cdgemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, alone, m)
        ops.append((A, B, alone))
    torch.cuda.synchronize()
    for _ in range(5):
        outs = [torch.zeros(m * n, dtype=torch.float64, device="cuda") for _ in ops]
        torch.cuda.synchronize()
        for (A, B, _), out, st in zip(ops, outs, (s1, s2)):
            cb.cdgemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, out, m, stream=st)
        torch.cuda.synchronize()
        for (_, _, alone), out in zip(ops, outs):
            assert torch.equal(alone, out)


def test_frob_diff():
    rng = np.random.default_rng(6)
    X = np.asfortranarray(rng.random((40, 30))); Y = np.asfortranarray(rng.random((44, 30)))
    d2, y2 = cb.frob_diff(dev(X), 40, dev(Y), 44, 40, 30)
    assert abs(d2 - ((X - Y[:40]) ** 2).sum()) < 1e-10 and abs(y2 - (Y[:40] ** 2).sum()) < 1e-10


def test_gemm_linearity_at_scale():
    """Size-independent property at a size the oracle cannot reach: (A1 + A2) B == A1 B + A2 B to rounding, n = 4096."""
    n = 4096
    A1 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A2 = torch.empty_like(A1); B = torch.empty_like(A1)
    cb.fill_drand48(A1, n, n, n, 0, 0, n, 0); cb.fill_drand48(A2, n, n, n, 0, 0, 2 * n, 1); cb.fill_drand48(B, n, n, n, 0, 0, n, 1)
    C12 = torch.empty_like(A1); Cs = torch.empty_like(A1)
    cb.cdgemm("N", "N", n, n, n, 1.0, A1 + A2, n, B, n, 0.0, C12, n)
    cb.cdgemm("N", "N", n, n, n, 1.0, A1, n, B, n, 0.0, Cs, n)
    cb.cdgemm("N", "N", n, n, n, 1.0, A2, n, B, n, 1.0, Cs, n)
    d2, r2 = cb.frob_diff(C12, n, Cs, n, n, n)
    assert np.sqrt(d2 / r2) <= 10 * n * EPS
