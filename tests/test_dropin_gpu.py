"""Drop-in evidence on the GPU: the reference's OWN test mains (test/MM/topo_pdgemm_unit.cxx, test/MM/test_spc.cxx),
compiled UNMODIFIED against include/CANDMC.h + include/candmc_compat/mpi.h and linked with libcandmc_b200.so
(`make -C oracle dropin`, only possible where /root/reference exists; the binaries travel in oracle/_ref/dropin/),
run one rank per GPU under tools/candmc_run and must print the reference's own PASS lines."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
RUN = os.path.join(ROOT, "tools", "candmc_run")


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _run(np_, exe, *args):
    if not (os.path.exists(os.path.join(DROPIN, exe)) and os.path.exists(RUN)):
        pytest.skip("drop-in binaries not built (needs /root/reference at build time)")
    if _ngpu() < np_:
        pytest.skip(f"needs {np_} GPUs")
    p = subprocess.run([RUN, "-np", str(np_), "-timeout", "300", os.path.join(DROPIN, exe), *args], capture_output=True,
                       text=True, timeout=400, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("np_,args", [(1, ["-n", "128"]), (1, ["-n", "96", "-ovp", "0"]), (4, ["-n", "128"]),
                                      (4, ["-n", "256", "-ovp", "0"]), (8, ["-n", "128"]), (8, ["-n", "256", "-ovp", "0"])])
def test_reference_d25_unit_test_passes_on_gpu(np_, args):
    out = _run(np_, "topo_pdgemm_unit", *args)
    assert "D25 UNIT TEST PASSED" in out and "FAILED" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("np_,args", [(1, []), (4, []), (4, ["-bidir", "0", "-m", "64", "-k", "32", "-n", "48"])])
def test_reference_split_cannon_test_passes_on_gpu(np_, args):
    out = _run(np_, "test_spc", *args)
    assert "Test passed." in out and "FAILED" not in out


@pytest.mark.gpu
def test_reference_bench_main_runs_on_gpu():
    out = _run(1, "topo_pdgemm_bench", "-n", "2048", "-niter", "2", "-nwarm", "1")
    assert "Gigaflops:" in out
