#!/usr/bin/env python
"""Generate tests/golden/canmm_ref_outputs.npz from the UNMODIFIED reference CANMM.

Runs oracle/_ref/ref_dump (our dump driver linked against the reference's own objects, built by `make -C oracle ref`
from /root/reference under the mini-MPI shim and scipy-OpenBLAS) for a handful of small grids and stores every
rank's C block.  Only runnable where /root/reference exists (the build container); the fixture it writes is
committed and is what the CPU and GPU parity tests read.

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# name -> (ranks, ref_dump argv, elements per rank)
CASES = {
    # d25_summa / d25_summa_ovp: the reference test's own configurations (scripts/test_all.sh: P=4 -> 2x2x1, P=8 -> 2x2x2)
    "d25_n96_q2_c1_ovp0": (4, ["d25", "96", "1", "0"], 48 * 48),
    "d25_n96_q2_c1_ovp1": (4, ["d25", "96", "1", "1"], 48 * 48),
    "d25_n64_q2_c2_ovp0": (8, ["d25", "64", "2", "0"], 32 * 32),
    "d25_n64_q2_c2_ovp1": (8, ["d25", "64", "2", "1"], 32 * 32),
    "d25_n40_q1_c1_ovp0": (1, ["d25", "40", "1", "0"], 40 * 40),
    "d25_n96_q4_c2_ovp0": (32, ["d25", "96", "2", "0"], 24 * 24),
    "d25_n90_q3_c1_ovp1": (9, ["d25", "90", "1", "1"], 30 * 30),
    # summa (never called by a live reference main, but its body is complete: summa.cxx:26-101)
    "summa_n64_q2": (4, ["summa", "64"], 32 * 32),
    "summa_n96_q3": (9, ["summa", "96"], 32 * 32),
    # bcast_cannon_4d with x2_np = 1 (the only configuration in which the reference terminates)
    "dcn_n64_x2_1_ovp0": (4, ["dcn", "64", "1", "0"], 32 * 32),
    "dcn_n64_x2_1_ovp1": (4, ["dcn", "64", "1", "1"], 32 * 32),
    # trans_A / trans_B set: the reference hands them to its local dgemm only (summa.cxx:97, d25_summa.cxx:185,
    # dual_cannon.cxx:163-166) — every block product is op(A block) * op(B block), blocks travel as stored
    "summa_n64_q2_TN": (4, ["summat", "64", "T", "N"], 32 * 32),
    "summa_n64_q2_NT": (4, ["summat", "64", "N", "T"], 32 * 32),
    "d25_n64_q2_c2_ovp0_TT": (8, ["d25t", "64", "2", "0", "T", "T"], 32 * 32),
    "d25_n96_q2_c1_ovp1_TN": (4, ["d25t", "96", "1", "1", "T", "N"], 48 * 48),
    "dcn_n64_x2_1_ovp0_TN": (4, ["dcnt", "64", "1", "0", "T", "N"], 32 * 32),
    "dcn_n64_x2_1_ovp1_NT": (4, ["dcnt", "64", "1", "1", "N", "T"], 32 * 32),
    "dcn_n64_x2_1_ovp0_TT": (4, ["dcnt", "64", "1", "0", "T", "T"], 32 * 32),
    # split-dimensional Cannon: spc <bidir> <ndim> <seed> <n> <m> <k> <alpha> <beta> <tB>
    "spc_bidir1_p4_m24_k16_n20_N": (4, ["spc", "1", "2", "3", "20", "24", "16", "1.2", "0.8", "N"], 24 * 20),
    "spc_bidir0_p4_m24_k16_n20_N": (4, ["spc", "0", "2", "3", "20", "24", "16", "1.2", "0.8", "N"], 24 * 20),
    "spc_bidir1_p4_m24_k16_n20_T": (4, ["spc", "1", "2", "3", "20", "24", "16", "1.2", "0.8", "T"], 24 * 20),
    "spc_bidir1_p9_m16_k12_n8_N": (9, ["spc", "1", "2", "3", "8", "16", "12", "1.2", "0.8", "N"], 16 * 8),
    "spc_bidir0_p9_m16_k12_n8_N": (9, ["spc", "0", "2", "3", "8", "16", "12", "1.2", "0.8", "N"], 16 * 8),
    "spc_bidir1_p16_ndim4_m8_k16_n8_N": (16, ["spc", "1", "4", "3", "8", "8", "16", "1.2", "0.8", "N"], 8 * 8),
    "spc_bidir0_p16_ndim4_m8_k16_n8_N": (16, ["spc", "0", "4", "3", "8", "8", "16", "1.2", "0.8", "N"], 8 * 8),
    # update_A (CAQR trailing update), W == NULL: upda <m> <k> <b> <nprow> <rrow> <rcol>; ragged sizes per rank
    "upda_m96_k64_b8_2x2_r00": (4, ["upda", "96", "64", "8", "2", "0", "0"], None),
    "upda_m80_k48_b8_2x3_r12": (6, ["upda", "80", "48", "8", "2", "1", "2"], None),
    "upda_m64_k32_b16_1x1": (1, ["upda", "64", "32", "16", "1", "0", "0"], None),
    "upda_m72_k40_b8_4x1_r20": (4, ["upda", "72", "40", "8", "4", "2", "0"], None),
    # update_A with the panel QR's upper-triangular W, W_is_T == false (the form QR_2D itself uses, qr_2d.cxx:325; T by
    # comp_bcast_T_from_W :179-208).  Roots where the reference's broadcast-root formula (:250) matches its drivers' grid.
    "updw_m96_k64_b8_2x2_r00": (4, ["updw", "96", "64", "8", "2", "0", "0"], None),
    "updw_m96_k64_b8_2x2_r11": (4, ["updw", "96", "64", "8", "2", "1", "1"], None),
    "updw_m80_k48_b8_2x3_r00": (6, ["updw", "80", "48", "8", "2", "0", "0"], None),
    "updw_m64_k32_b16_1x1": (1, ["updw", "64", "32", "16", "1", "0", "0"], None),
    "updw_m72_k40_b8_4x1_r20": (4, ["updw", "72", "40", "8", "4", "2", "0"], None),
    "updw_m48_k72_b8_1x3_r02": (3, ["updw", "48", "72", "8", "1", "0", "2"], None),
    # update_Yamamoto_A (agg == NULL, alg/QR/qr_2d/qr_y2d.cxx:68-120): updy <m> <k> <b> <nprow> <rrow> <rcol>
    "updy_m96_k64_b8_2x2_r00": (4, ["updy", "96", "64", "8", "2", "0", "0"], None),
    "updy_m80_k48_b8_2x3_r12": (6, ["updy", "80", "48", "8", "2", "1", "2"], None),
    "updy_m64_k32_b16_1x1": (1, ["updy", "64", "32", "16", "1", "0", "0"], None),
    "updy_m72_k40_b8_4x1_r20": (4, ["updy", "72", "40", "8", "4", "2", "0"], None),
    # update_Yamamoto_A WITH an aggregator over the k/b panels of an m x k block column, driven as QR_Yamamoto_2D drives it
    # (qr_y2d.cxx:38-62,68-120,171-277): updyagg <m> <k> <b> <nprow> <rrow> <rcol>; per rank A (mb0 x kb0) | aQm (mb0 x k) | aT (k x k).
    # Sizes keep at least one row block on every rank at every step: with mb == 0 the reference sums an uninitialised buffer
    # into aT (:54-56 clear b*b of the b*n doubles the all-reduce then adds).
    "updyagg_m96_k32_b8_2x2_r00": (4, ["updyagg", "96", "32", "8", "2", "0", "0"], None),
    "updyagg_m96_k32_b8_2x2_r11": (4, ["updyagg", "96", "32", "8", "2", "1", "1"], None),
    "updyagg_m80_k24_b8_2x3_r12": (6, ["updyagg", "80", "24", "8", "2", "1", "2"], None),
    "updyagg_m64_k32_b16_1x1": (1, ["updyagg", "64", "32", "16", "1", "0", "0"], None),
    "updyagg_m72_k24_b8_4x1_r20": (4, ["updyagg", "72", "24", "8", "4", "2", "0"], None),
}


def main():
    if not os.path.exists(os.path.join(REFDIR, "ref_dump")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (ranks, argv, elems) in CASES.items():
            prefix = os.path.join(tmp, name)
            cmd = [os.path.join(REFDIR, "mpirun"), "-np", str(ranks), "-timeout", "120", "-threads", "1",
                   os.path.join(REFDIR, "ref_dump")] + argv + [prefix]
            subprocess.check_call(cmd)
            blocks = []
            for r in range(ranks):
                a = np.fromfile(f"{prefix}.r{r}.f64", dtype="<f8")
                assert elems is None or a.size == elems, (name, r, a.size, elems)
                blocks.append(a)
            if elems is None:   # ragged per-rank sizes: one key per rank
                for r, a in enumerate(blocks):
                    out[f"{name}.r{r}"] = a
            else:
                out[name] = np.stack(blocks)
            print(f"{name}: {ranks} ranks x {elems if elems else [a.size for a in blocks]} doubles")
    path = os.path.join(ROOT, "tests", "golden", "canmm_ref_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
