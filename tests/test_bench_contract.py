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


def test_host_generator_of_the_sampled_entries_check_is_the_reference_generator():
    """bench.py's full-size check regenerates operands on the host with its own numpy LCG; it must be the reference unit test's
    per-element generator (test/MM/topo_pdgemm_unit.cxx:250-256) bit for bit — compared with the oracle's blocks"""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle_py as orc

    n = 48
    A, B = orc.d25_blocks(n, 1, 1)
    rows, cols = np.meshgrid(np.arange(n, dtype=np.uint64), np.arange(n, dtype=np.uint64), indexing="ij")
    for which, ref in ((0, A[0]), (1, B[0])):
        got = bench.host_unit_entries(n, rows.ravel(), cols.ravel(), which).reshape(n, n)
        assert np.array_equal(got, np.asarray(ref).reshape(n, n, order="F") if np.asarray(ref).ndim == 1 else np.asarray(ref))
    # the large seeds of the headline size (col * n + row up to 2^30) through the split 64-bit multiply
    big = bench.host_unit_entries(32768, np.array([32767, 5], dtype=np.uint64), np.array([32767, 32000], dtype=np.uint64), 1)
    for v, (r, c) in zip(big, ((32767, 32767), (5, 32000))):
        x = (((c * 32768 + r) & 0xFFFFFFFF) << 16) | 0x330E
        for _ in range(2):
            x = (0x5DEECE66D * x + 0xB) & ((1 << 48) - 1)
        assert v == x / float(1 << 48)
