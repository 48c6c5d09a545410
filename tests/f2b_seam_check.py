"""The reference's symmetric full -> band reduction (alg/SE/full_to_band.cxx, unmodified, driven by oracle/ref_f2b_dump.cxx)
with nothing but cdgemm replaced by integration/cdgemm_gpu.cxx — every local multiply in libcandmc_b200.so — against the outputs
of the all-host run of the same binary's twin that tests/golden/f2b_ref_outputs.npz holds (tests/golden/make_golden_f2b.py).

    python tests/f2b_seam_check.py <case name in the fixture>      (LD_PRELOAD the CPU simulator to run without a GPU)

Prints one JSON line: levels compared, largest relative difference of any rank's local array after a level's update."""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def main():
    name = sys.argv[1]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "f2b_ref_outputs.npz"))
    P, n, b, bs, levels = (int(x) for x in gold[f"{name}.args"])
    env = dict(os.environ, CANDMC_SEAM_VERBOSE="1", CANDMC_CDGEMM_MIN_FLOP=os.environ.get("CANDMC_CDGEMM_MIN_FLOP", "0"))
    worst, compared = 0.0, 0
    with tempfile.TemporaryDirectory() as td:
        p = subprocess.run([os.path.join(REFDIR, "mpirun"), "-np", str(P), "-timeout", "200", "-threads", "1",
                            os.path.join(REFDIR, "dropin", "ref_f2b_dump_cdgemm_gpu"), str(n), str(b), str(bs), os.path.join(td, "x")],
                           env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stdout[-2000:], p.stderr[-2000:])
            sys.exit(1)
        in_lib = [int(x) for x in re.findall(r"cdgemm_gpu: (\d+) products in the library", p.stderr)]
        for key in gold.files:
            m = re.fullmatch(re.escape(name) + r"\.L(\d+)\.r(\d+)\.(Ain|Y|Aout)", key)
            if not m:
                continue
            got = np.fromfile(os.path.join(td, f"x.L{m.group(1)}.r{m.group(2)}.{m.group(3)}"))
            ref = gold[key]
            assert got.shape == ref.shape, (key, got.shape, ref.shape)
            worst = max(worst, float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)))
            compared += 1
    ok = compared > 0 and worst <= 1e-12 and len(in_lib) == P and all(c > 0 for c in in_lib)
    print(json.dumps({"case": name, "ranks": P, "arrays_compared": compared, "max_rel_diff": worst,
                      "products_in_library_per_rank": in_lib, "ok": ok}))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
