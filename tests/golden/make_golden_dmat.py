"""Generates tests/golden/dmat_ref_outputs.npz: per-rank outputs of the UNMODIFIED reference DMatrix pack operations
(alg/SE/dmatrix.cxx, driven by oracle/_ref/ref_dmat_dump under the mini-MPI) for the cases below.
Run in the build container after `make -C oracle ref`:   python tests/golden/make_golden_dmat.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# name: (ranks, op, nrow, ncol, b, nprow, rrow, rcol, factor, sliced)
CASES = {
    "repv_2x2": (4, "repv", 16, 24, 2, 2, 0, 0, 1, 0),
    "repv_2x2_r11_sliced": (4, "repv", 24, 32, 4, 2, 1, 1, 1, 1),
    "repv_4x2_r30": (8, "repv", 32, 16, 2, 4, 3, 0, 1, 0),
    "reph_2x2": (4, "reph", 16, 24, 2, 2, 0, 0, 1, 0),
    "reph_2x4_r13_sliced": (8, "reph", 12, 48, 2, 2, 1, 3, 1, 1),
    "rsh_2x2_r01": (4, "rsh", 16, 24, 2, 2, 0, 1, 1, 0),
    "rsh_2x4": (8, "rsh", 8, 32, 2, 2, 0, 0, 1, 0),
    "rsh_1x8_r05": (8, "rsh", 6, 48, 3, 1, 0, 5, 1, 0),
    "tpd_2x2": (4, "tpd", 16, 24, 2, 2, 0, 0, 1, 0),
    "tpd_3x3_r12_sliced": (9, "tpd", 24, 30, 2, 3, 1, 2, 1, 1),
    "fc_2x2_f2": (4, "fc", 16, 24, 2, 2, 0, 0, 2, 0),
    "fc_2x2_f3_r10_sliced": (4, "fc", 28, 20, 2, 2, 1, 0, 3, 1),
    "fr_2x2_f2": (4, "fr", 16, 24, 2, 2, 0, 0, 2, 0),
    "fr_2x3_f4_r02_sliced": (6, "fr", 12, 54, 2, 2, 0, 2, 4, 1),
    # one rank: what a single-GPU box can run (the fold kernels are fully exercised; the collectives degenerate to copies)
    "fc_1x1_f2": (1, "fc", 48, 20, 4, 1, 0, 0, 2, 0),
    "fc_1x1_f3_sliced": (1, "fc", 39, 11, 3, 1, 0, 0, 3, 1),
    "fr_1x1_f5_sliced": (1, "fr", 14, 32, 2, 1, 0, 0, 5, 1),
    "rsh_1x1": (1, "rsh", 12, 10, 2, 1, 0, 0, 1, 0),
    "repv_1x1_sliced": (1, "repv", 12, 10, 2, 1, 0, 0, 1, 1),
    "reph_1x1": (1, "reph", 12, 10, 2, 1, 0, 0, 1, 0),
    "tpd_1x1_sliced": (1, "tpd", 12, 10, 2, 1, 0, 0, 1, 1),
}


def main():
    exe = os.path.join(REFDIR, "ref_dmat_dump")
    if not os.path.exists(exe):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (ranks, op, *args) in CASES.items():
            prefix = os.path.join(tmp, name)
            cmd = [os.path.join(REFDIR, "mpirun"), "-np", str(ranks), "-timeout", "60", "-threads", "1", exe, op,
                   *map(str, args), prefix]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
            for r in range(ranks):
                out[f"{name}.r{r}"] = np.fromfile(f"{prefix}.r{r}.f64", dtype="<f8")
            out[f"{name}.args"] = np.array([ranks] + list(args), dtype=np.int64)
            print(name, [out[f"{name}.r{r}"].size for r in range(ranks)])
    path = os.path.join(HERE, "dmat_ref_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
