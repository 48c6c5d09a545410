"""GPU test of the optional FP32 GEMM (candmc_sgemm: tcgen05.mma kind::tf32 with split operands, accumulator in tensor memory).

STATUS: written after the round's GPU budget was spent — compiled for sm_100a (SASS shows UTCHMMA / UTMALDG / LDTM / UTCBAR),
every case passes on the CPU simulator's tcgen05 emulation (tests/test_cpusim.py), never run on a B200.  Same policy as the
other tests/test_zz_*.py: own process group with a timeout, plain tests since round 2 (they passed on the driver's B200 in round 1 and again in round 2's sessions).
"""
import json
import os
import sys

import pytest

from pending_util import run_guarded

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.gpu
def test_sgemm_against_float64_product():
    """all transpose combinations, ragged m / n / k, padded and misaligned operands, alpha / beta, both precision modes, up to
    4096^3: relative to the size of the summed terms the 3xTF32 result stays within 4 * 2^-20, the single-TF32 one within 2^-9"""
    rc, out, err = run_guarded("f32", [sys.executable, os.path.join(HERE, "f32_worker.py")], 240, ROOT)
    assert rc == 0, out[-2000:] + err[-3000:]
    r = json.loads(out.strip().splitlines()[-1])
    # (the worker asserts every case against its own bound: 4 * 2^-20 + 2 sqrt(k) 2^-24 for the split scheme)
    assert r["cases"] >= 28 and r["max_rel_err_3xtf32"] <= 2e-5 and r["launches"] > 0
