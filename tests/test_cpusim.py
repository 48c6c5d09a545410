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


DIST_MAIN = [1, 2, 4, 8] if FULL else [2, 4, 8]
DIST_PENDING = [1, 2, 3, 4, 6, 8, 9, 16] if FULL else [1, 3, 4, 6, 8, 9]
GOLD = np.load(os.path.join(HERE, "golden", "lu_offload_ref_outputs.npz"))
SCRIPTS = sorted(k[: -len("__script")] for k in GOLD.files if k.endswith("__script"))
REDIST = [(24, 36, 2, 2, 3, 0, 0), (36, 24, 3, 3, 2, 1, 1), (30, 30, 5, 3, 3, 0, 2), (16, 16, 4, 1, 4, 0, 3), (512, 256, 32, 2, 2, 1, 0)]

for _n in DIST_MAIN:     # longest first.  On 8 ranks the fused GEMM + depth all-reduce is switched on for the 2x2x2 grid too (opt-in in
    # the product until a B200 has seen it): the kernel's peer-memory epilogue and ipc.cu run for real, over simulated CUDA IPC
    _job(f"main{_n}", _torchrun(_n, 29700 + _n, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0",
         CANDMC_TEST_FUSED_GRIDS="1" if _n == 8 else "0", CANDMC_TEST_UNSEEN="1")   # + the cases no B200 has run yet
_job("unseen1", _torchrun(1, 29745, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_TEST_UNSEEN="only")
# the default data path since round 2 (SUMMA panels by copy engines, transport.h) named explicitly on 2x2 and, together with the
# opt-in fused depth sum, on 2x2x2 — deferred streams, LIFO order; and round 1's path (NCCL kernels, one launch per k-chunk)
_job("nccl4", _torchrun(4, 29747, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_TEST_NCCL_PANELS="1",
     CPUSIM_SCHED="lifo")
# ... and the automatic fallback when peer windows are unavailable: NCCL panels (full-width communicators) under the launch groups
_job("ncclmerge4", _torchrun(4, 29748, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_PANEL_TRANSPORT="0",
     CPUSIM_SCHED="random:31")
# fault injection: one rank cannot map its peers' memory (cudaIpcOpenMemHandle fails there) — every communicator that rank is in
# must agree to stay on NCCL (panels AND the depth sum), the others keep their windows
_job("ipcfail2", _torchrun(2, 29749, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_TEST_KC="8", CANDMC_TEST_QUICK="1", CANDMC_TEST_UNSEEN="1", CPUSIM_IPC_FAIL_RANK="1",
     CPUSIM_SCHED="lifo")
_job("ipcfail4", _torchrun(4, 29750, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_TEST_KC="8", CANDMC_TEST_QUICK="1", CPUSIM_IPC_FAIL_RANK="2",
     CPUSIM_SCHED="lifo")
# (the copy-engine transport and launch groups of mode 2 ARE main4 / main8 since round 2 — and main4@lifo / main8@random below)
# opt-in: the last panel of a sweep multiplied in one launch over its k-chunks (B read chunk-major through one tensor map by the
# product's own kernel on the PTX emulation); the validated 2x2 suite with the switch on
_job("merge4", _torchrun(4, 29743, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_TEST_MERGE_PANELS="1",
     CPUSIM_SCHED="lifo")
# ... and the doubling groups of mode 3 (mode 2 is the default: main4 / main8)
_job("merge4_doubling", _torchrun(4, 29746, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CANDMC_TEST_MERGE_PANELS="3",
     CPUSIM_SCHED="random:22")
# the hot kernel itself on the PTX emulation, and the 4-rank suite with every GEMM going through it
_job("kernel", [sys.executable, os.path.join(SIM, "probe_gemm.py")])
_job("kernel_pack", [sys.executable, os.path.join(SIM, "probe_pack.py")])
_job("kernel_bchunk", [sys.executable, os.path.join(HERE, "bchunk_worker.py")])
_job("kernel_f32", [sys.executable, os.path.join(HERE, "f32_worker.py"), "--fuzz", "17", "12" if not FULL else "80"])
if FULL:
    _job("main4_device_gemm", _torchrun(4, 29719, os.path.join(HERE, "dist_worker.py")), CANDMC_TEST_PENDING="0", CPUSIM_GEMM="device")
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
for _m, _sched in (("race", "sync"), ("race", "lifo"), ("ordered", "lifo"), ("ordered", "random:5")):
    _job(f"probe_{_m}_{_sched}", [sys.executable, os.path.join(SIM, "probe_local.py"), _m], CPUSIM_SCHED=_sched)
# bench.py itself (argument handling, byte accounting of the end-to-end leg, JSON contract) on 1 and 8 simulated ranks; the
# numbers are not measurements.  `--n` travels in CPUSIM_ARGS because torchrun's parser trips over it after the script name.
for _n in (1, 8):
    _job(f"bench{_n}", _torchrun(_n, 29790 + _n, os.path.join(SIM, "run_sim.py"), "bench.py", "--gpus", str(_n), "--steps", "1",
                                 "--warmup", "1", "--no-cpu-baseline"), CPUSIM_ARGS="--n 512", CPUSIM_SCHED="sync" if _n == 1 else "lifo")

_job("bench_watchdog", _torchrun(2, 29795, os.path.join(SIM, "run_sim.py"), "bench.py", "--gpus", "2", "--steps", "1", "--warmup", "1",
                                 "--no-cpu-baseline", "--e2e-watchdog", "0.0005"), CPUSIM_ARGS="--n 512", CPUSIM_SCHED="sync")

# randomised cases (tests/fuzz_worker.py: random grids, sizes, paddings, roots, transposes, host/device operands, knobs)
FUZZ = [(4, 3, "lifo")] + ([(p, sd, pol) for p in (1, 2, 4, 6, 8, 9, 16) for sd, pol in ((11, "sync"), (12, "lifo"), (13, "random:13"))]
                           if FULL else [])
for _p, _sd, _pol in FUZZ:
    _job(f"fuzz{_p}_{_sd}", _torchrun(_p, 29600 + _p * 20 + _sd, os.path.join(HERE, "fuzz_worker.py"), str(_sd), "24"), CPUSIM_SCHED=_pol)

for _sd, _pol in ((1, "lifo"),) + (((2, "random:2"), (3, "sync"), (4, "lifo")) if FULL else ()):
    _job(f"lufuzz{_sd}", [sys.executable, os.path.join(HERE, "off_worker.py"), "fuzz", str(_sd), "1"], CPUSIM_SCHED=_pol)

# the reference's OWN test mains (compiled unmodified against include/, oracle/_ref/dropin — present where /root/reference
# was available at build time): the simulator build is preloaded in front of libcandmc_b200.so, which exports the same ABI
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
PRELOAD = os.path.join(SIM, "_build", "libcandmc_b200_cpusim.so")
DROPIN_CASES = {
    "d25_p1": ("candmc_run", 1, "topo_pdgemm_unit", ["-n", "96", "-ovp", "0"], "D25 UNIT TEST PASSED", "sync"),
    "d25_p4": ("candmc_run", 4, "topo_pdgemm_unit", ["-n", "128"], "D25 UNIT TEST PASSED", "lifo"),
    "d25_p8": ("candmc_run", 8, "topo_pdgemm_unit", ["-n", "128"], "D25 UNIT TEST PASSED", "sync"),
    "spc_p4": ("candmc_run", 4, "test_spc", [], "Test passed.", "lifo"),
    "spc_p4_uni": ("candmc_run", 4, "test_spc", ["-bidir", "0", "-m", "64", "-k", "32", "-n", "48"], "Test passed.", "sync"),
    "lu_pp": ("mpirun", 4, "lu_pp_gpu", ["-n", "256", "-b_sm", "8", "-b_lrg", "32"], "test passed", "sync"),
    "lu_tp": ("mpirun", 4, "lu_tp_gpu", ["-n", "256", "-b_sm", "8", "-b_lrg", "32"], "test passed", "lifo"),
    # the reference's own D25 test with the opt-in peer-memory paths switched on from the environment (an unmodified main
    # cannot call setters): panels by copy engines, depth sum fused into the GEMM epilogue (n = 1024: b = 512 = 2 * 128 * c)
    "d25_p8_peer_paths": ("candmc_run", 8, "topo_pdgemm_unit", ["-n", "1024", "-ovp", "0"], "D25 UNIT TEST PASSED", "lifo"),
    # the reference's own 2D QR test with its trailing updates in the library (integration/qr_2d_upd_A_gpu.cxx): as shipped
    # (QR_2D_pipe, W_is_T form) and sent through QR_2D_2D (oracle/qr_2d_tap.cxx: T from the panel factor / from the aggregated Y)
    "qr_pipe_p4": ("mpirun", 4, "test_qr_2d_gpu", ["96", "48", "8", "2"], "Test successful.", "lifo"),
    "qr_pipe_p1": ("mpirun", 1, "test_qr_2d_gpu", ["64", "32", "4", "1"], "Test successful.", "sync"),
    "qr_2d_p4": ("mpirun", 4, "test_qr_2d_2d_gpu", ["128", "64", "4", "2"], "Test successful.", "lifo"),
    "qr_2d_p9": ("mpirun", 9, "test_qr_2d_2d_gpu", ["144", "72", "4", "3"], "Test successful.", "sync"),
    "qr_2d_p2": ("mpirun", 2, "test_qr_2d_2d_gpu", ["64", "32", "8", "2"], "Test successful.", "lifo"),
    # ... and test/QR/test_qr_y2d.cxx (QR_Yamamoto_2D_2D with its aggregator): upd_Yamamoto_A served by candmc_upd_Yamamoto_A
    "qr_y2d_p4": ("mpirun", 4, "test_qr_y2d_gpu", ["96", "48", "8", "2", "16"], "Test successful.", "lifo"),
    "qr_y2d_p6": ("mpirun", 6, "test_qr_y2d_gpu", ["96", "48", "4", "2", "24"], "Test successful.", "sync"),
    # integration/cdgemm_gpu.cxx: the reference's cdgemm itself over candmc_dgemm, nothing else replaced (threshold 0: every
    # product, however small, goes through the library) — its QR test and its 2.5D LU test (offload seam on the host fallback)
    "cdgemm_qr_p4": ("mpirun", 4, "test_qr_2d_cdgemm_gpu", ["96", "48", "8", "2"], "Test successful.", "lifo"),
    "cdgemm_lu_p4": ("mpirun", 4, "lu_pp_cdgemm_gpu", ["-n", "256", "-b_sm", "8", "-b_lrg", "32"], "test passed", "sync"),
}
DROPIN_ENV = {"d25_p8_peer_paths": dict(CANDMC_PANEL_TRANSPORT="1", CANDMC_FUSED_REDUCE="2", CANDMC_MIN_KCHUNK="64"),
              "qr_pipe_p4": dict(CANDMC_SEAM_VERBOSE="1"), "qr_pipe_p1": dict(CANDMC_SEAM_VERBOSE="1"),
              "qr_2d_p4": dict(CANDMC_SEAM_VERBOSE="1", QR_TAP_B2="32"), "qr_2d_p9": dict(CANDMC_SEAM_VERBOSE="1", QR_TAP_B2="24"),
              "qr_2d_p2": dict(CANDMC_SEAM_VERBOSE="1"), "qr_y2d_p4": dict(CANDMC_SEAM_VERBOSE="1"),
              "qr_y2d_p6": dict(CANDMC_SEAM_VERBOSE="1"),
              "cdgemm_qr_p4": dict(CANDMC_SEAM_VERBOSE="1", CANDMC_CDGEMM_MIN_FLOP="0"),
              "cdgemm_lu_p4": dict(CANDMC_SEAM_VERBOSE="1", CANDMC_CDGEMM_MIN_FLOP="0")}
HAVE_DROPIN = all(os.path.exists(os.path.join(DROPIN, c[2])) for c in DROPIN_CASES.values()) and \
    os.path.exists(os.path.join(ROOT, "tools", "candmc_run")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mpirun"))
if HAVE_DROPIN:
    for _k, (_launcher, _np, _exe, _args, _needle, _sched) in DROPIN_CASES.items():
        _run = os.path.join(ROOT, "tools", "candmc_run") if _launcher == "candmc_run" else os.path.join(ROOT, "oracle", "_ref", "mpirun")
        _job(f"dropin_{_k}", [_run, "-np", str(_np), "-timeout", "200", os.path.join(DROPIN, _exe), *_args], LD_PRELOAD=PRELOAD,
             CPUSIM_SCHED=_sched, **DROPIN_ENV.get(_k, {}))

# the reference's symmetric full -> band reduction with nothing but cdgemm replaced (integration/cdgemm_gpu.cxx), against the
# all-host outputs of the same driver in tests/golden/f2b_ref_outputs.npz
if HAVE_DROPIN and os.path.exists(os.path.join(DROPIN, "ref_f2b_dump_cdgemm_gpu")):
    for _c, _sched in (("f2b_p4_n48_b8_s4", "lifo"), ("f2b_p9_n72_b12_s4", "sync")):
        _job(f"f2bseam_{_c}", [sys.executable, os.path.join(HERE, "f2b_seam_check.py"), _c], LD_PRELOAD=PRELOAD, CPUSIM_SCHED=_sched)

# the same workers with DEFERRED streams: work runs only at host synchronisation points, and among the runnable streams
# the one whose head was enqueued last goes first (lifo) or a random one — a missing event dependency computes garbage
ADVERSARIAL = [("main4", "lifo"), ("pending4", "lifo"), ("main8", "random:7")] + [(f"lu_{_s}_1", "lifo") for _s in SCRIPTS]
if FULL:
    ADVERSARIAL += [(f"main{_n}", _p) for _n in DIST_MAIN for _p in ("lifo", "random:1", "random:2")]
    ADVERSARIAL += [(f"pending{_n}", _p) for _n in DIST_PENDING for _p in ("lifo", "random:3")]
    ADVERSARIAL += [(f"lu_{_s}_1", "random:4") for _s in SCRIPTS] + [("lu_trailing", "lifo"), ("lu_trailing", "random:9")]
    ADVERSARIAL = sorted(set(ADVERSARIAL))
for _i, (_name, _sched) in enumerate(ADVERSARIAL):
    _cmd, _extra = JOBS[_name]
    _cmd = [str(29800 + _i) if (_k > 0 and _cmd[_k - 1] == "--master-port") else _a for _k, _a in enumerate(_cmd)]
    _job(f"{_name}@{_sched}", _cmd, **dict(_extra, CPUSIM_SCHED=_sched))


@pytest.fixture(scope="module", autouse=True)
def sim_runs():
    """builds the simulator library, then runs every job (each a fresh process or process group; most of a job's time is
    `import torch`) four at a time"""
    subprocess.check_call(["make", "-s", "-C", SIM, "-j8"])
    assert os.path.exists(os.path.join(SIM, "_build", "libcandmc_b200_cpusim.so"))

    import time
    times = {}

    def run(name):
        cmd, extra = JOBS[name]
        t0 = time.time()
        try:
            p = subprocess.run(cmd, cwd=ROOT, env=_env(**extra), capture_output=True, text=True, timeout=600)
            return name, (p.returncode, p.stdout, p.stderr)
        except subprocess.TimeoutExpired as e:
            return name, (-999, str(e.stdout or ""), "TIMEOUT " + str(e.stderr or ""))
        finally:
            times[name] = round(time.time() - t0, 1)

    with ThreadPoolExecutor(max_workers=4) as pool:
        for name, res in pool.map(run, list(JOBS)):
            RESULTS[name] = res
    with open(os.path.join(SIM, "_build", "job_times.json"), "w") as f:   # where the suite's time goes (longest first)
        json.dump(dict(sorted(times.items(), key=lambda kv: -kv[1])), f, indent=1)
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
    """update_Yamamoto_A, update_A with the panel QR's factor W (comp_bcast_T_from_W: the form QR_2D uses; 1x1, 1x3, 2x2, 4x1,
    2x3 grids, poisoned W everywhere but on the root), the DMatrix pack operations (bit-exact against the unmodified dmatrix.cxx) and
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


@pytest.mark.parametrize("name,sched", ADVERSARIAL)
def test_stream_dependencies_hold_under_adversarial_scheduling(name, sched):
    """every event / stream-order dependency the schedules rely on is explicit: results do not change when everything
    that is not ordered runs in the opposite (or a random) order"""
    rc, so, se = RESULTS[f"{name}@{sched}"]
    assert rc == 0, so[-3000:] + se[-3000:]
    out = json.loads([line for line in so.splitlines() if line.startswith("{")][-1])
    if name.startswith("lu_trailing"):
        assert out["first_block_exact"] and max(out["rel_frobenius"], out["panel_rel"], out["rows_rel"]) <= out["bound"]
    elif name.startswith("lu_"):
        assert out["padding_untouched"] and out["max_abs_vs_reference"] <= 1e-12 and out["exact_fraction"] > 0.5
        assert out["stats"]["cross_stream_waits"] > 0   # the scoreboard had to order transfers behind in-flight GEMMs
    else:
        assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0


@pytest.mark.parametrize("case", sorted(DROPIN_CASES))
def test_reference_test_mains_pass_on_the_simulator(case):
    """test/MM/topo_pdgemm_unit.cxx, test/MM/test_spc.cxx, the 2.5D LU unit tests (sixteen offload calls served by
    libcandmc_lu_offload.so) and test/QR/test_qr_2d.cxx (upd_A served by candmc_upd_A) — the reference's own sources and PASS
    criteria, our C++ drop-in layer and host schedules"""
    if not HAVE_DROPIN:
        pytest.skip("drop-in binaries not built (needs /root/reference at build time)")
    rc, so, se = RESULTS[f"dropin_{case}"]
    assert rc == 0, so[-2000:] + se[-2000:]
    assert DROPIN_CASES[case][4] in so and "FAILED" not in so and "test failed" not in so.lower()
    if case.startswith("cdgemm_"):
        import re

        counts = [(int(a), int(b)) for a, b in re.findall(r"cdgemm_gpu: (\d+) products in the library, (\d+) on the host BLAS", se)]
        assert len(counts) == DROPIN_CASES[case][1] and all(a > 0 and b == 0 for a, b in counts), se[-1500:]
    if case.startswith("qr_") or case == "cdgemm_qr_p4":
        import re

        res = float(re.findall(r"\|\|A-QR\|\|_2 = (\S+)", so)[-1])
        assert res == res and res <= 1e-9, so[-1500:]
    if case.startswith("qr_"):
        # the reference's QR test (SURVEY 8f N1's pin) prints "Test successful." for a NaN residual too: read the number, and
        # make sure the trailing updates went through the library (integration/qr_2d_upd_A_gpu.cxx reports each one)
        import re

        res = float(re.findall(r"\|\|A-QR\|\|_2 = (\S+)", so)[-1])
        assert res == res and res <= 1e-9, so[-1500:]
        assert "qr_2d_upd_A_gpu: upd_" in se
        if case.startswith("qr_y2d_"):
            assert "upd_Yamamoto_A" in se
        elif case.startswith("qr_2d_"):
            assert "form=T from W" in se and ("QR_TAP_B2" not in DROPIN_ENV[case] or "form=T from Y" in se)
        else:
            assert "form=W is T" in se


@pytest.mark.parametrize("case", ["f2b_p4_n48_b8_s4", "f2b_p9_n72_b12_s4"])
def test_reference_full_to_band_with_cdgemm_in_the_library(case):
    """alg/SE/full_to_band.cxx, unmodified, under the mini-MPI: cdgemm alone comes from integration/cdgemm_gpu.cxx, so the panel
    QR's updates, Y^T A, U V^T and the rank-2b update of every level multiply in the library; every rank's local array after
    each recorded level equals the all-host run's (the fixture) to 1e-12, and every rank really sent products to the library"""
    if f"f2bseam_{case}" not in RESULTS:
        pytest.skip("drop-in binaries not built (needs /root/reference at build time)")
    rc, so, se = RESULTS[f"f2bseam_{case}"]
    assert rc == 0, so[-2000:] + se[-2000:]
    out = json.loads([line for line in so.splitlines() if line.startswith("{")][-1])
    assert out["ok"] and out["arrays_compared"] >= 36 and out["max_rel_diff"] <= 1e-12


@pytest.mark.parametrize("nproc,seed,policy", FUZZ)
def test_randomised_cases_on_the_simulator(nproc, seed, policy):
    rc, so, se = RESULTS[f"fuzz{nproc}_{seed}"]
    assert rc == 0, so[-3000:] + se[-3000:]
    out = json.loads([line for line in so.splitlines() if line.startswith("{")][-1])
    assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0


@pytest.mark.parametrize("seed", [1, 2, 3, 4] if FULL else [1])
def test_random_lu_offload_scripts_on_the_simulator(seed):
    """random op sequences over the three offloaded matrices (tests/off_script.py random_script), two streams, deferred
    scheduling: the scoreboard must order every transfer behind exactly the GEMMs it conflicts with"""
    rc, so, se = RESULTS[f"lufuzz{seed}"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["max_rel_vs_oracle"] <= 1e-11 and r["exact_fraction"] > 0.8


@pytest.mark.parametrize("nproc", [1, 8])
def test_bench_script_logic_on_the_simulator(nproc):
    """bench.py end to end on the simulator: one JSON line with the contract's keys, both legs agree with the cross-check,
    and the end-to-end byte accounting matches what the ranks really upload (2x2x2: a layer's rank uploads only the blocks
    its panels use — one block per rank on average instead of two)"""
    rc, so, se = RESULTS[f"bench{nproc}"]
    assert rc == 0, so[-2000:] + se[-3000:]
    lines = [line for line in so.splitlines() if line.startswith('{"metric"')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["n_gpus"] == nproc and d["warmup"] >= 3 and d["dtype"] == "f64" and d["gpu_launches"] > 0
    assert d["rel_frobenius_vs_cublas_crosscheck"] <= d["tolerance_10_n_eps"]
    # ... and the check that shares no code with the device path: sampled entries against operands regenerated on the host
    assert d["max_rel_err_sampled_entries_vs_host_regenerated_operands"] <= d["tolerance_10_n_eps"]
    assert d["e2e"]["valid"] and d["e2e"]["rel_frobenius_vs_device_path"] <= d["tolerance_10_n_eps"]
    b = d["config"]["block"]
    # two end-to-end passes: the host-operand settings B200s have run, then the library's defaults; `e2e` is one of them
    first, second = d["e2e_passes"]
    assert (first["host_operand_settings"], second["host_operand_settings"]) == ("round1_settings", "library_defaults")
    assert d["e2e"] in (first, second)
    for p_ in (first, second):
        assert p_["valid"] and p_["rel_frobenius_vs_device_path"] <= d["tolerance_10_n_eps"]
        assert p_["d2h_bytes_per_step"] == nproc * b * b * 8
    assert first["h2d_bytes_per_step"] == 2 * nproc * b * b * 8
    assert second["h2d_bytes_per_step"] == (2 if nproc == 1 else 8) * b * b * 8


def test_bench_watchdog_keeps_the_line_when_the_second_pass_does_not_return():
    """--e2e-watchdog: a second end-to-end pass that does not come back in time must not cost the measurement — the line is
    printed with the first pass and the process ends with status 0"""
    rc, so, se = RESULTS["bench_watchdog"]
    assert rc == 0, so[-2000:] + se[-3000:]
    lines = [line for line in so.splitlines() if line.startswith('{"metric"')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["e2e"]["valid"] and d["e2e"]["host_operand_settings"] == "round1_settings" and d["value"] > 0
    assert "watchdog" in d["e2e_passes"][1]["error"]


@pytest.mark.parametrize("nproc", [4, 8])
def test_copy_engine_panel_transport_on_the_simulator(nproc):
    """candmc_set_panel_transport(1): panel chunks DMA-written into the consumers' IPC windows, ready / done flags as 4-byte
    DMAs, cuStreamWaitValue32 on the consumer side — the whole distributed suite incl. repeated multiplies on one grid (window
    halves reused, windows regrown), no ncclBroadcast left on the path"""
    out = _dist(f"main{nproc}")
    assert out["panel_transport_sends_rank0"] > 50


def test_nccl_panels_and_per_chunk_launches_on_the_simulator():
    """round 1's data path, still selectable (candmc_set_panel_transport(0), candmc_set_merge_panels(0)) and the automatic
    fallback when peer windows are unavailable: ncclBroadcast panels on the capped communicators, one launch per k-chunk"""
    out = _dist("nccl4")
    assert out["panel_transport_sends_rank0"] == 0 and out["merged_panel_launches_all_ranks"] == [0, 0]
    out = _dist("ncclmerge4")   # the fallback of the default schedule: launch groups fed by ncclBroadcast
    assert out["panel_transport_sends_rank0"] == 0 and out["merged_panel_launches_all_ranks"][0] > 0


def test_cases_no_b200_has_seen_yet_on_the_simulator():
    """tests/dist_worker.py:unseen_cases — what was written after round 2's GPU minutes were spent (upd_A with T formed from Y
    and with host operands, error recovery of the fused depth sum, the peer-argument check, trans flags on the k-split, in
    bcast_cannon_4d and against the reference's flagged outputs).  On 2, 4 and 8 ranks they ride on the main jobs
    (CANDMC_TEST_UNSEEN=1); this is the single-rank set on its own, the way tests/test_zz_unseen_gpu.py runs it on a B200."""
    out = _dist("unseen1")
    assert out["checks_rank0"] >= 4


def test_ranks_agree_to_leave_peer_windows_when_one_cannot_map_them():
    """a local failure inside the collective window set-up becomes the agreed outcome instead of leaving
    the peers inside an all-gather.  Rank 1 of 2 / rank 2 of 4 cannot open IPC handles: the whole validated suite still passes,
    the 1x1x2 depth sum and every communicator with that rank in it on NCCL, rank 0's row communicator (ranks 0, 1 of 4)
    still on copy engines."""
    out = _dist("ipcfail2")
    assert out["panel_transport_sends_rank0"] == 0
    out = _dist("ipcfail4")
    assert out["panel_transport_sends_rank0"] > 0


@pytest.mark.parametrize("job", ["merge4", "main4", "merge4_doubling", "main8"])
def test_merged_panel_launches_on_the_simulator(job):
    """candmc_set_merge_panels(1 / 2 / 3) under the validated 2x2 and 2x2x2 suites (deferred streams, LIFO / random order): same
    results, and the merged launch with chunk-major B really ran (the 3x3 grid, ragged tile columns, host operands and the
    fallbacks are cases of the pending group: test_widening_rows_on_the_simulator)"""
    out = _dist(job)
    assert out["merged_panel_launches_all_ranks"][0] > 0 and out["merged_panel_launches_all_ranks"][1] > 0


def test_hot_kernel_reads_chunk_major_b_through_one_tensor_map():
    """candmc_dgemm_chunked_b on the PTX emulation: bit for bit the plain-layout launch of the same kernel, within 10 k eps of
    numpy, loud errors for chunk depths that are not multiples of the k-tile"""
    rc, so, se = RESULTS["kernel_bchunk"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["cases"] >= 10 and not r["failures"] and r["launches"] >= 14


def test_pack_kernels_on_the_simulator():
    """candmc_b200/csrc/pack.cu: the tiled lda_cpy / scaled lda_cpy kernels (double2 and scalar paths, ragged tiles) and both
    transpose kernels — TMA load, turn inside swizzled shared memory, TMA store; and the LDG/STG fallback — bit for bit"""
    rc, so, se = RESULTS["kernel_pack"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["cases"] >= 36 and r["launches"] >= 36


def test_hot_gemm_kernel_on_the_ptx_emulation():
    """candmc_b200/csrc/gemm_f64.cu itself — TMA boxes with the 128-byte swizzle and zero fill, the mbarrier ring, the
    permuted fragment loads, DMMA.8x8x4, split-K, the dynamic and the static tile scheduler, both epilogues — against numpy"""
    rc, so, se = RESULTS["kernel"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["cases"] >= 37 and r["max_rel_err"] <= 1e-13   # 25 with 128 x 128 tiles + 12 with the two-CTAs-per-SM shape


def test_fp32_tcgen05_kernel_on_the_emulation():
    """candmc_b200/csrc/gemm_f32.cu itself — K-major TMA boxes, the operand split in shared memory (hi in place, lo behind), UMMA
    shared-memory / instruction descriptors decoded by CUTLASS's bit fields, TF32 truncation, the double-buffered accumulator in
    tensor memory, the 32-lane x 32-column epilogue loads with the warp-quarter rule enforced, the K-major pack of the other
    transpose cases — against a float64 product: 3xTF32 within 4 * 2^-20 of the summed terms, one TF32 product within 2^-9"""
    rc, so, se = RESULTS["kernel_f32"]
    assert rc == 0, so[-2000:] + se[-3000:]
    r = json.loads(so.strip().splitlines()[-1])
    assert r["cases"] >= 36 and r["max_rel_err_3xtf32"] <= r["tol_3xtf32"] and r["max_rel_err_1xtf32"] <= 2.0 ** -9   # 24 fixed + 12 random
    assert r["max_rel_err_1xtf32"] > 100 * r["max_rel_err_3xtf32"]   # the split really buys the accuracy


@pytest.mark.skipif(not FULL, reason="CANDMC_CPUSIM_FULL=1")
def test_distributed_suite_with_every_gemm_on_the_emulated_kernel():
    _dist("main4_device_gemm")


def test_adversarial_scheduler_exposes_a_missing_event():
    assert "copied 0.0" in RESULTS["probe_race_sync"][1]          # enqueue order hides the race
    assert "copied 8.0" in RESULTS["probe_race_lifo"][1]          # ... the deferred scheduler does not
    assert "copied 0.0" in RESULTS["probe_ordered_lifo"][1]       # with the event the answer no longer depends on the policy
    assert "copied 0.0" in RESULTS["probe_ordered_random:5"][1]


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
