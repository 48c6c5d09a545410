"""GPU test of the CAQR seam (SURVEY.md §8f N1 names test/QR/test_qr_2d.cxx as the pin of the trailing update): the reference's
OWN 2D QR test, unmodified, with every trailing update — the GEMM pair, the all-reduce of Y^T A and the triangular solve of
upd_A (alg/QR/qr_2d/qr_2d.cxx:224-282) — running in libcandmc_b200.so through integration/qr_2d_upd_A_gpu.cxx.  The panels
(TSQR, Householder reconstruction) stay the reference's host code over the mini-MPI shim and OpenBLAS (`make -C oracle dropin`;
the binaries travel in oracle/_ref/dropin/).  test_qr_2d_gpu drives QR_2D_pipe as the test is shipped (W_is_T form);
test_qr_2d_2d_gpu is the same test sent through QR_2D_2D (oracle/qr_2d_tap.cxx): T from the panel's factor and T from the
aggregated Y; test_qr_y2d_gpu is test/QR/test_qr_y2d.cxx (QR_Yamamoto_2D_2D) with upd_Yamamoto_A in the library.  Criterion: the reference's own ||A - QR|| <= 1e-9 line — parsed, because the test prints "Test successful."
for a NaN as well.

STATUS: written after round 2's GPU minutes were spent.  Green on the CPU simulator (tests/test_cpusim.py, 1 / 2 / 4 / 9
ranks); no B200 has run it, hence xfail(strict=False): XPASS when right, and it cannot turn the validated suite red on first
contact.  The marker goes away once a round has seen it pass.
"""
import os
import re

import pytest

from pending_util import run_guarded

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def qr_residual(stdout):
    """the reference test's verdict line '||A-QR||_2 = 1.80E-14' (test/QR/test_qr_2d.cxx:338)"""
    m = re.findall(r"\|\|A-QR\|\|_2 = (\S+)", stdout)
    assert m, stdout[-2000:]
    return float(m[-1])


CASES = [  # np, exe, m, k, b, nprow, outer block of the tapped variant
    (1, "test_qr_2d_gpu", 256, 128, 16, 1, None),
    (1, "test_qr_2d_2d_gpu", 256, 128, 16, 1, None),
    (1, "test_qr_2d_2d_gpu", 256, 128, 16, 1, 64),
    (1, "test_qr_2d_gpu", 1024, 512, 64, 1, None),
    (4, "test_qr_2d_gpu", 256, 128, 16, 2, None),
    (4, "test_qr_2d_2d_gpu", 256, 128, 8, 2, 32),
    (4, "test_qr_2d_gpu", 2048, 1024, 64, 2, None),
    # test/QR/test_qr_y2d.cxx (QR_Yamamoto_2D_2D, aggregator on the host): upd_Yamamoto_A in the library; here the last number is
    # the test's own fifth argument, the outer block
    (1, "test_qr_y2d_gpu", 256, 128, 16, 1, 64),
    (4, "test_qr_y2d_gpu", 512, 256, 16, 2, 64),
    # integration/cdgemm_gpu.cxx: the reference's cdgemm itself in the library, nothing else replaced (default size threshold)
    (1, "test_qr_2d_cdgemm_gpu", 1024, 512, 64, 1, None),
    (4, "test_qr_2d_cdgemm_gpu", 2048, 1024, 128, 2, None),
]


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="the reference's QR test over the GPU upd_A seam: first B200 run pending (written after round 2's GPU budget was spent)")
@pytest.mark.parametrize("np_,exe,m,k,b,nprow,b2", CASES)
def test_reference_qr_2d_test_passes_with_gpu_trailing_updates(np_, exe, m, k, b, nprow, b2):
    path = os.path.join(REFDIR, "dropin", exe)
    if not (os.path.exists(path) and os.path.exists(os.path.join(REFDIR, "mpirun"))):
        pytest.skip("drop-in binaries not built (needs /root/reference at build time)")
    if _ngpu() < np_:
        pytest.skip(f"needs {np_} GPUs")
    env = dict(os.environ, CANDMC_SEAM_VERBOSE="1")
    if b2 is not None:
        env["QR_TAP_B2"] = str(b2)
    cmd = [os.path.join(REFDIR, "mpirun"), "-np", str(np_), "-timeout", "200", "-threads", "2", path, str(m), str(k), str(b), str(nprow)]
    if exe == "test_qr_y2d_gpu":
        cmd.append(str(b2))
    rc, so, se = run_guarded("qr_dropin", cmd, 300, ROOT, env=env)
    assert rc == 0, so[-2000:] + se[-2000:]
    res = qr_residual(so)
    assert res == res and res <= 1e-9, so[-1500:]          # the reference's criterion, NaN-proof
    if "cdgemm" in exe:
        lib_calls = [int(x) for x in re.findall(r"cdgemm_gpu: (\d+) products in the library", se)]
        assert len(lib_calls) == np_ and all(c > 0 for c in lib_calls), se[-1500:]
    else:
        assert "qr_2d_upd_A_gpu: upd_" in se               # ... and the updates really went through the library


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="the reference's full -> band reduction over the cdgemm seam: first B200 run pending")
@pytest.mark.parametrize("case,np_", [("f2b_p1_n24_b8_s4", 1), ("f2b_p4_n48_b8_s4", 4)])
def test_reference_full_to_band_with_cdgemm_in_the_library(case, np_):
    """alg/SE/full_to_band.cxx, unmodified, with nothing but cdgemm replaced (integration/cdgemm_gpu.cxx, threshold 0: every
    product in the library) against the all-host outputs in tests/golden/f2b_ref_outputs.npz (tests/f2b_seam_check.py)"""
    import json
    import sys

    if not os.path.exists(os.path.join(REFDIR, "dropin", "ref_f2b_dump_cdgemm_gpu")):
        pytest.skip("drop-in binaries not built (needs /root/reference at build time)")
    if _ngpu() < np_:
        pytest.skip(f"needs {np_} GPUs")
    rc, so, se = run_guarded("qr_dropin", [sys.executable, os.path.join(HERE, "f2b_seam_check.py"), case], 300, ROOT)
    assert rc == 0, so[-2000:] + se[-2000:]
    out = json.loads([line for line in so.splitlines() if line.startswith("{")][-1])
    assert out["ok"] and out["max_rel_diff"] <= 1e-12


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="the reference's 2.5D LU test over the cdgemm seam: first B200 run pending")
def test_reference_lu_test_passes_with_cdgemm_in_the_library():
    """test/LU/lu_25d_pvt_unit_test.cxx (partial pivoting, offload seam on the reference's host fallback) with nothing but
    cdgemm replaced (integration/cdgemm_gpu.cxx, default size threshold): the reference's own 'test passed' line"""
    path = os.path.join(REFDIR, "dropin", "lu_pp_cdgemm_gpu")
    if not os.path.exists(path):
        pytest.skip("drop-in binaries not built (needs /root/reference at build time)")
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    cmd = [os.path.join(REFDIR, "mpirun"), "-np", "4", "-timeout", "200", "-threads", "2", path, "-n", "2048", "-b_sm", "64", "-b_lrg", "512"]
    rc, so, se = run_guarded("qr_dropin", cmd, 300, ROOT, env=dict(os.environ, CANDMC_SEAM_VERBOSE="1"))
    assert rc == 0, so[-2000:] + se[-2000:]
    assert "test passed" in so.lower() and "fail" not in so.lower(), so[-1500:]
    lib_calls = [int(x) for x in re.findall(r"cdgemm_gpu: (\d+) products in the library", se)]
    assert len(lib_calls) == 4 and all(c > 0 for c in lib_calls), se[-1500:]
