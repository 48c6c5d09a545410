"""GPU test of the hot kernel reading B chunk-major through one tensor map (candmc_dgemm_chunked_b, the launch behind
candmc_set_merge_last_panel).

STATUS: written after the round's GPU budget was spent — every case passes on the CPU simulator's PTX emulation
(tests/test_cpusim.py), never run on a B200.  Same policy as the other tests/test_zz_*.py: own process group with a timeout,
plain tests since round 2 (they passed on the driver's B200 in round 1 and again in round 2's sessions).
"""
import json
import os
import sys

import pytest

from pending_util import run_guarded

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.gpu
def test_chunk_major_b_launch_equals_plain_launch():
    """whole and ragged tiles, one to eight chunks, both layouts of A, up to the merged launch of a b = 8192 panel (K = 7168 in
    1024-deep chunks): bit for bit the plain-layout launch of the same kernel and within 10 k eps of a float64 numpy product"""
    rc, out, err = run_guarded("bchunk", [sys.executable, os.path.join(HERE, "bchunk_worker.py")], 300, ROOT)
    assert rc == 0, out[-2000:] + err[-3000:]
    r = json.loads(out.strip().splitlines()[-1])
    assert r["cases"] >= 14 and not r["failures"] and r["launches"] > 0
