"""GPU parity tests of the distributed multiplies: tests/dist_worker.py under torchrun at every world size the box has."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _run(nproc, extra_env=None, timeout=900):
    env = dict(os.environ)
    env.update(extra_env or {})
    env.setdefault("NCCL_DEBUG", "WARN")
    if nproc == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc), os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0, f"dist_worker failed at world size {nproc}"
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["failed_all_ranks"] == 0 and out["checks_rank0"] > 0
    assert out["launches_rank0"] > 0, "no candmc kernels were launched: native path not exercised"
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [1, 2, 4, 8])
def test_distributed_parity(nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    _run(nproc)


@pytest.mark.gpu
def test_full_size_property_single_gpu():
    """n = 8192 on one GPU: d25_summa (1x1x1) vs an independently generated local GEMM and vs cuBLAS."""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    _run(1, {"CANDMC_TEST_BIG_N": "8192"})
