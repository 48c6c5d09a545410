"""GPU tests of the block-cyclic <-> blocked redistribution (candmc_redistribute) and of the Yamamoto form
of the CAQR trailing update (N1, candmc_update_Yamamoto_A).

STATUS: first run on B200s in round 1's driver session (1 GPU) and in round 2's 1- and 4-GPU sessions (profiles/r02_*): green.
Every case still runs in its own process group with a timeout (pending_util.run_guarded), so a hang costs minutes.
"""
import json
import os
import sys

import pytest

from pending_util import run_guarded

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(24, 36, 2, 2, 3, 0, 0), (36, 24, 3, 3, 2, 1, 1), (30, 30, 5, 3, 3, 0, 2),
                                  (16, 16, 4, 1, 4, 0, 3), (2048, 1024, 64, 2, 2, 1, 0), (8192, 8192, 128, 2, 2, 0, 0)])
def test_kernels_play_every_rank_on_one_gpu(case):
    """permute + pack kernels exactly as candmc_redistribute launches them, the all-to-all replaced by device copies"""
    rc, out, err = run_guarded("redist", [sys.executable, os.path.join(HERE, "redist_worker.py"), *map(str, case)], 300, ROOT)
    assert rc == 0, out[-2000:] + err[-3000:]
    r = json.loads(out.strip().splitlines()[-1])
    assert r["to_blocked_exact"] and r["to_blocked_matches_generator"] and r["to_cyclic_exact"]
    assert r["single_rank_identity"] and r["launches"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [1, 2, 4, 8])
def test_pending_distributed_cases(nproc):
    """tests/dist_worker.py's pending group: candmc_redistribute over NCCL (every grid shape the world size allows) and
    update_Yamamoto_A against the reference's own outputs"""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, CANDMC_TEST_PENDING="1")
    env.setdefault("NCCL_DEBUG", "WARN")
    worker = os.path.join(HERE, "dist_worker.py")
    cmd = [sys.executable, worker] if nproc == 1 else [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
        "--master-port", str(29600 + nproc), worker]
    rc, so, se = run_guarded("redist", cmd, 400, ROOT, env=env)
    assert rc == 0, so[-3000:] + se[-3000:]
    out = json.loads([l for l in so.splitlines() if l.startswith("{")][-1])
    assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_nccl_panel_paths_and_fused_grid_sum(nproc):
    """the validated distributed suite once more with the data paths that are NOT the default: SUMMA panels and Cannon shifts by
    NCCL kernels on the CTA-capped communicators (round 1's default; since round 2 panels travel by copy engines into CUDA-IPC
    windows, transport.cu), and on 8 GPUs the fused GEMM + depth all-reduce on the 2x2x2 grid (opt-in)."""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, CANDMC_TEST_NCCL_PANELS="1", CANDMC_TEST_FUSED_GRIDS="1" if nproc == 8 else "0")
    env.setdefault("NCCL_DEBUG", "WARN")
    worker = os.path.join(HERE, "dist_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29620 + nproc), worker]
    rc, so, se = run_guarded("peer_paths", cmd, 400, ROOT, env=env)
    assert rc == 0, so[-3000:] + se[-3000:]
    out = json.loads([l for l in so.splitlines() if l.startswith("{")][-1])
    assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0
    assert out["panel_transport_sends_rank0"] == 0
