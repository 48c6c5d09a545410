"""GPU run of tests/dist_worker.py:unseen_cases — everything written after round 2's GPU minutes were spent that lives in the
distributed worker: candmc_upd_A with T formed from Y and with host operands (1 / 2 / 4 ranks), error recovery of the fused
depth sum and the peer-argument check (2 ranks), transposed operands on the 1x1x2 k-split (2 ranks), trans flags in summa /
d25_summa / bcast_cannon_4d against the unmodified reference's flagged outputs (4 / 8 ranks).

STATUS: green on the CPU simulator (tests/test_cpusim.py: main2 / main4 / main8 / unseen1 jobs); no B200 has run them, hence
xfail(strict=False) — XPASS when right, and a first-contact failure cannot turn the validated suite red.  The cases move into
the validated group (dist_worker.main) and the marker goes away once a round has seen them pass.
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
@pytest.mark.xfail(strict=False, reason="cases written after round 2's GPU budget was spent: first B200 run pending")
@pytest.mark.parametrize("nproc", [1, 2, 4, 8])
def test_unseen_cases(nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, CANDMC_TEST_UNSEEN="only")
    env.setdefault("NCCL_DEBUG", "WARN")
    worker = os.path.join(HERE, "dist_worker.py")
    cmd = [sys.executable, worker] if nproc == 1 else [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
        "--master-port", str(29660 + nproc), worker]
    rc, so, se = run_guarded("unseen", cmd, 400, ROOT, env=env)
    assert rc == 0, so[-3000:] + se[-3000:]
    out = json.loads([l for l in so.splitlines() if l.startswith("{")][-1])
    assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0
