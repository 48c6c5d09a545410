"""CPU runs of the GPU test workers on the functional simulator (tests/cpusim): the product's host schedules (stream / event
ordering, staging, NCCL call sequences on real multi-process grids) and its simple kernels (pack, permute, fold, sparse
rows, trsm; every thread executed as a fiber) run for real, the TMA + DMMA GEMM kernel is replaced by plain loops.

What this is NOT: a product path or a fallback (libcandmc_b200.so is not involved and still fails without a B200), nor a
substitute for the `-m gpu` parity tests — it is the cheapest place to catch index arithmetic, ordering and protocol bugs
of the code AROUND the hot kernel before GPU minutes are spent (it found its first one, the DMatrix extent check that
rejected the ragged matrices the reference accepts, the day it was written).  Checks are the workers' own: golden outputs
of the unmodified reference + the oracle, same tolerances as on the GPU.
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SIM = os.path.join(HERE, "cpusim")
FULL = os.environ.get("CANDMC_CPUSIM_FULL") == "1"   # every world size; the default set keeps the CPU suite short


JOBS = {}     # name -> command (+ env); all of them are started by one fixture and run four at a time
RESULTS = {}


def _env(**extra):
    env = dict(os.environ, CANDMC_CPUSIM="1", CPUSIM_TIMEOUT="40", OMP_NUM_THREADS="1")
    env.update(extra)
    return env


def _job(name, cmd, **extra):
    JOBS[name] = (cmd, extra)
    return name


def _torchrun(nproc, port, script, *args):
    if nproc == 1:
        return [sys.executable, script, *args]
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
            "--master-port", str(port), script, *args]


DIST_MAIN = [1, 2, 4, 8] if FULL else [4, 8]
DIST_PENDING = [1, 2, 4, 6, 8, 9] if FULL else [1, 4, 6]
GOLD = np.load(os.path.join(HERE, "golden", "lu_offload_ref_outputs.npz"))
SCRIPTS = sorted(k[: -len("__script")] for k in GOLD.files if k.endswith("__script"))
REDIST = [(24, 36, 2, 2, 3, 0, 0), (36, 24, 3, 3, 2, 1, 1), (30, 30, 5, 3, 3, 0, 2), (16, 16, 4, 1, 4, 0, 3), (512, 256, 32, 2, 2, 1, 0)]

for _n in DIST_MAIN:     # longest first
    _job(f"main{_n}", _torchrun(_n, 29700 + _n, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0")
for _n in DIST_PENDING:
    _job(f"pending{_n}", _torchrun(_n, 29720 + _n, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="1")
for _m in ("mismatch", "stuck"):
    _job(f"proto_{_m}", _torchrun(2, 29755 + len(_m), os.path.join(SIM, "probe_dist.py"), _m), CPUSIM_TIMEOUT="3")
for _s in SCRIPTS:
    for _o in (0, 1):
        _job(f"lu_{_s}_{_o}", [sys.executable, os.path.join(HERE, "off_worker.py"), "script", _s, str(_o)])
_job("lu_trailing", [sys.executable, os.path.join(HERE, "off_worker.py"), "trailing", "384", "48", "1"])
for _c in REDIST:
    _job("redist_" + "_".join(map(str, _c)), [sys.executable, os.path.join(HERE, "redist_worker.py"), *map(str, _c)])
for _m in ("oob", "uninit", "gemm_range"):
    _job(f"probe_{_m}", [sys.executable, os.path.join(SIM, "probe_local.py"), _m])


@pytest.fixture(scope="module", autouse=True)
def sim_runs():
    """builds the simulator library, then runs every job (each a fresh process or process group; most of a job's time is
    `import torch`) four at a time"""
    subprocess.check_call(["make", "-s", "-C", SIM, "-j8"])
    assert os.path.exists(os.path.join(SIM, "_build", "libcandmc_b200_cpusim.so"))

    def run(name):
        cmd, extra = JOBS[name]
        try:
            p = subprocess.run(cmd, cwd=ROOT, env=_env(**extra), capture_output=True, text=True, timeout=600)
            return name, (p.returncode, p.stdout, p.stderr)
        except subprocess.TimeoutExpired as e:
            return name, (-999, str(e.stdout or ""), "TIMEOUT " + str(e.stderr or ""))

    with ThreadPoolExecutor(max_workers=4) as pool:
        for name, res in pool.map(run, list(JOBS)):
            RESULTS[name] = res
    return RESULTS


def _dist(name):
    rc, so, se = RESULTS[name]
    assert rc == 0, so[-3000:] + se[-3000:]
    out = json.loads([line for line in so.splitlines() if line.startswith("{")][-1])
    assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0 and out["launches_rank0"] > 0
    return out


@pytest.mark.parametrize("nproc", DIST_MAIN)
def test_validated_distributed_suite_on_the_simulator(nproc):
    """summa / d25_summa(_ovp) / bcast_cannon_4d / kput,kuni_cannon / upd_A / update_A on 1x1x1, 1x1x2, 2x2x1, 2x2x2 grids:
    host and device operands, padded leading dimensions, tiny k-chunks (the pipelined sweep), all against the reference's
    own outputs and the oracle"""
    _dist(f"main{nproc}")


@pytest.mark.parametrize("nproc", DIST_PENDING)
def test_widening_rows_on_the_simulator(nproc):
    """update_Yamamoto_A, the DMatrix pack operations (bit-exact against the unmodified dmatrix.cxx) and
    candmc_redistribute over the simulated NCCL on 1x1, 2x2, 4x1, 1x4, 2x3, ... grids"""
    _dist(f"pending{nproc}")


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("name", SCRIPTS)
def test_lu_offload_scripts_on_the_simulator(name, overlap):
    """the LU accelerator seam's operation scripts against the outputs of the unmodified lu_offload.cxx"""
    rc, so, se = RESULTS[f"lu_{name}_{overlap}"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["padding_untouched"] and r["max_abs_vs_reference"] <= 1e-12 and r["max_abs_vs_oracle"] <= 1e-12
    assert r["exact_fraction"] > 0.5 and r["launches"] > 0


def test_lu_trailing_update_pattern_on_the_simulator():
    rc, so, se = RESULTS["lu_trailing"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["first_block_exact"] and max(r["rel_frobenius"], r["panel_rel"], r["rows_rel"]) <= r["bound"]


@pytest.mark.parametrize("case", REDIST)
def test_redistribution_kernels_on_the_simulator(case):
    """permute / pack kernels exactly as candmc_redistribute launches them, one process playing every rank"""
    rc, so, se = RESULTS["redist_" + "_".join(map(str, case))]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["to_blocked_exact"] and r["to_blocked_matches_generator"] and r["to_cyclic_exact"] and r["single_rank_identity"]


# ---- the simulator's own detectors must fire (a checker that cannot fail proves nothing) ----------------------------------
def test_simulator_detects_out_of_bounds_kernel_writes():
    rc, so, se = RESULTS["probe_oob"]
    assert rc != 0 and "OUT-OF-BOUNDS WRITE" in se


def test_simulator_poisons_uninitialised_device_memory():
    rc, so, se = RESULTS["probe_uninit"]
    assert rc == 0 and "uninit-ok" in so, se[-2000:]


def test_simulator_checks_gemm_operand_ranges():
    rc, so, se = RESULTS["probe_gemm_range"]
    assert rc != 0 and "runs past the end of its allocation" in se


@pytest.mark.parametrize("mode,needle", [("mismatch", "MISMATCHED OPERATIONS"), ("stuck", "NO PROGRESS")])
def test_simulated_nccl_reports_protocol_errors(mode, needle):
    rc, so, se = RESULTS[f"proto_{mode}"]
    assert rc != 0 and needle in se, se[-3000:]
