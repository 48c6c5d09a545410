"""CPU tests of bench.py's contract pieces that need no GPU: the reference arm prints exactly ONE JSON line on stdout
with the keys the driver reads, and the product arm refuses to run without a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "topo_pdgemm_bench")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True and d["scaling"] == "strong"


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU box")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0 and p.stdout.strip() == ""   # no number without the CUDA path
