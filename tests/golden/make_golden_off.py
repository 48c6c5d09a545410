"""Generates tests/golden/lu_offload_ref_outputs.npz: the outputs of the UNMODIFIED reference LU offload seam
(alg/LU/lu_offload.cxx, host fallback, driven by oracle/_ref/ref_off_dump) for the seeded scripts of tests/off_script.py.

Run in the build container (needs /root/reference to have been compiled: `make -C oracle ref`):
    python tests/golden/make_golden_off.py
The npz stores, per script, the script text itself and every output record, so the tests need neither the reference nor
ref_off_dump at run time.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from off_script import make_scripts  # noqa: E402


def read_records(path):
    raw = open(path, "rb").read()
    outs, pos = [], 0
    while pos < len(raw):
        n = int(np.frombuffer(raw, dtype=np.int64, count=1, offset=pos)[0])
        pos += 8
        outs.append(np.frombuffer(raw, dtype=np.float64, count=n, offset=pos).copy())
        pos += 8 * n
    return outs


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_off_dump")
    if not os.path.exists(exe):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    data = {}
    for name, text in make_scripts().items():
        with tempfile.TemporaryDirectory() as d:
            sp, op = os.path.join(d, "s.txt"), os.path.join(d, "o.bin")
            open(sp, "w").write(text)
            subprocess.run([exe, sp, op], check=True)
            outs = read_records(op)
        data[f"{name}__script"] = np.array(text)
        data[f"{name}__nout"] = np.array(len(outs))
        for i, o in enumerate(outs):
            data[f"{name}__out{i}"] = o
        print(f"{name}: {len(text.splitlines())} ops, {len(outs)} output records, {sum(o.size for o in outs)} doubles")
    out = os.path.join(HERE, "lu_offload_ref_outputs.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
