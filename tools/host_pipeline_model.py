#!/usr/bin/env python
"""Timeline model of the one-GPU host-streamed multiply (host_pipelined_gemm_nn, candmc_b200/csrc/mm_algs.cu): three streams
(H2D, compute, D2H), the dependencies of the real schedule, and measured rates — PCIe 45-55 GB/s per direction, 36.1 TFLOP/s,
and 0.5-2 us per row of a 2-D copy (the figure behind upload_chunks' comment in mm_algs.cu).  It evaluates the cuts the
LIBRARY produces (candmc_host_pipeline_cut through the C ABI; host arithmetic, no GPU needed), so what is modelled is what
runs.  A model, not a measurement: bench.py's end-to-end leg reports the measured passes.

    python tools/host_pipeline_model.py [n]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from candmc_b200._lib import check, lib  # noqa: E402

TF = 36.1e12


def cut(n, k, panels):
    w, c = (C.c_int64 * 128)(), (C.c_int64 * 128)()
    nw, nc = C.c_int(), C.c_int()
    check(lib().candmc_host_pipeline_cut(n, k, panels, w, c, 128, C.byref(nw), C.byref(nc)))
    return list(w[: nw.value]), list(c[: nc.value])


def simulate(m, n, k, widths, kchunks, row_cost, bw):
    """Seconds until the last C panel is in host memory."""
    h2d = comp = d2h = 0.0
    for kc in kchunks:   # first panel: A column slab (wide rows) + B row slab (one narrow row per column) per k-chunk
        h2d += m * kc * 8 / bw + widths[0] * row_cost + kc * widths[0] * 8 / bw
        comp = max(comp, h2d) + 2.0 * m * widths[0] * kc / TF
    g_done, c_free = [comp], []
    d2h = max(d2h, comp) + m * widths[0] * 8 / bw
    c_free.append(d2h)
    for j in range(1, len(widths)):
        w = widths[j]
        if j >= 2:
            h2d = max(h2d, g_done[j - 2])       # the B buffer of panel j-2 is free
        h2d += k * w * 8 / bw
        start = max(comp, h2d, c_free[j - 2] if j >= 2 else 0.0)
        comp = start + 2.0 * m * w * k / TF
        g_done.append(comp)
        d2h = max(d2h, comp) + m * w * 8 / bw
        c_free.append(d2h)
    return d2h


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    m = k = n
    ideal = 2.0 * m * n * k / TF
    doubling = ([n // 8] * 7 + [n // 16, n // 32, n // 32], [k // 64, k // 64, k // 32, k // 16, k // 8, k // 4, k // 2])
    cases = [("8 equal panels (measured: 34.9 TFLOP/s)", cut(n, k, 8)), ("16 equal panels", cut(n, k, 16)),
             ("doubling k-chunks + shrinking tail (rejected)", doubling), ("library default (graduated)", cut(n, k, 0))]
    print(f"n = {n}: multiply alone {ideal * 1e3:.0f} ms; exposed ms / end-to-end TFLOP/s at (row cost us, PCIe GB/s)")
    grid = [(0.5e-6, 55e9), (0.5e-6, 45e9), (2e-6, 55e9), (2e-6, 45e9)]
    print(" " * 48 + "".join(f"({rc * 1e6:.1f}, {bw / 1e9:.0f})".rjust(16) for rc, bw in grid))
    for name, (w, c) in cases:
        cells = []
        for rc, bw in grid:
            t = simulate(m, n, k, w, c, rc, bw)
            cells.append(f"{(t - ideal) * 1e3:5.0f} / {2.0 * m * n * k / t / 1e12:5.2f}".rjust(16))
        print(name.ljust(48) + "".join(cells))


if __name__ == "__main__":
    main()
