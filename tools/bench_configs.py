#!/usr/bin/env python
"""Measure BASELINE.json's other configurations (parity-test cases, not bench.py lines) on the GPUs at hand:
  config 2: 2D SUMMA FP64 n=16384 on 4 GPUs (2x2)                         -> summa
  config 4: Cannon FP64 n=24576 on 4 GPUs, shift overlap                   -> bcast_cannon_4d (x1_np=1, x2_np=2), kput_cannon
  config 5: CAQR trailing-update shape m=65536, n=8192, k=512 on 1/2/4 GPUs -> upd_A (TN GEMM, all-reduce, trsm, NN GEMM)
Launch: python tools/bench_configs.py            (1 GPU: config 5 on a 1x1 grid)
        torchrun --nproc-per-node 4 tools/bench_configs.py
Prints one JSON line per configuration on rank 0 (device-timed with CUDA events, max over ranks).  Every line carries a
result check at full size, made outside the timed region: the rank's block of the product against an independent cross-check
(operands regenerated locally with the device generator, multiplied by cuBLAS through torch.matmul — the use of cuBLAS
BASELINE's north_star allows), relative Frobenius error, max over ranks, against the tolerance 10 * k * eps of north_star
(k = the contracted dimension).
`--pending` adds the SURVEY §8f widening rows that have not been measured yet: upd_Yamamoto_A at the config-5 shape, the
redistribution's permute kernels (GB/s against the HBM copy peak) and candmc_redistribute over NCCL, and the LU seam's
trailing-update step (GEMM + the panel download queued behind it)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import candmc_b200 as cb  # noqa: E402

PEAK = 148 * 4 * 16 * 2 * 1.965e9 / 1e12
# `--shrink S` divides every extent by S: a dry run of this script's own logic on the CPU simulator
# (python tests/cpusim/run_sim.py tools/bench_configs.py --pending --shrink 32); numbers of such a run mean nothing
SHRINK = int(sys.argv[sys.argv.index("--shrink") + 1]) if "--shrink" in sys.argv else 1


def sz(x):
    return max(x // SHRINK, 2)


def main():
    rank = int(os.environ.get("RANK", 0)); ws = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = cb.init_world(rank, ws, local)

    def timed(fn, steps=3, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timeline(name, fn):
        """`--timeline`: one more call with every DMMA launch bracketed by events — rank 0's launches with the gaps in front of
        them (a shift or a panel that was not hidden under the previous multiply shows up as a gap)"""
        if "--timeline" not in sys.argv:
            return
        import ctypes as C
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        cb.lib().candmc_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        cap = 512
        ts, te, cnt = (C.c_double * cap)(), (C.c_double * cap)(), C.c_int64()
        cb.lib().candmc_profile_gemm_timeline(ts, te, cap, C.byref(cnt))
        cb.lib().candmc_profile_enable(0)
        if rank == 0:
            busy = sum(te[i] - ts[i] for i in range(cnt.value))
            sys.stderr.write(f"timeline {name}: call {e0.elapsed_time(e1):.3f} ms, {cnt.value} DMMA launches busy {busy:.3f} ms; idx start end dur gap_before\n")
            for i in range(cnt.value):
                sys.stderr.write(f"  {i:3d} {ts[i]:9.3f} {te[i]:9.3f} {te[i] - ts[i]:8.3f} {(ts[i] - te[i - 1]) if i else 0.0:8.3f}\n")

    def report(name, flops, ms, extra=None):
        if rank == 0:
            tf = flops / (ms * 1e-3) / 1e12
            line = {"config": name, "n_gpus": ws, "ms": ms, "tflops": tf, "pct_of_fp64_tensor_peak": 100 * tf / (PEAK * ws)}
            line.update(extra or {})
            print(json.dumps(line), flush=True)

    EPS = 2.220446049250313e-16

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gen(rows, cols, row0, col0, n, which):
        """rows x cols block at (row0, col0) of the generated n x n matrix, as a torch matrix M[i, j] (column-major storage)"""
        X = torch.empty(rows * cols, dtype=torch.float64, device="cuda")
        cb.fill_drand48(X, rows, cols, rows, row0, col0, n, which)
        return X.view(cols, rows).T

    def check_block(Cm, b, ref, kdim):
        """relative Frobenius error of my b x b column-major block against the torch matrix `ref`, max over ranks"""
        got = Cm.view(b, b).T
        err = float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref))
        err = maxr(err)
        tol = 10 * kdim * EPS
        return {"rel_frobenius_vs_cublas_crosscheck": err, "tolerance_10_k_eps": tol, "check_passed": bool(err <= tol)}

    def blocks(b, row0, col0, n):
        A = torch.empty(b * b, dtype=torch.float64, device="cuda"); B = torch.empty_like(A); Cm = torch.empty_like(A)
        cb.fill_drand48(A, b, b, b, row0, col0, n, 0); cb.fill_drand48(B, b, b, b, row0, col0, n, 1)
        return A, B, Cm

    if ws == 4:
        # ---- config 2: SUMMA n = 16384, 2x2 ----
        n = sz(16384); g = cb.d25_grid(world, 1); b = n // 2
        A, B, Cm = blocks(b, g["row"] * b, g["col"] * b, n)
        args = cb.ctb_args_t(n=n, lda_A=b, lda_B=b, lda_C=b, buffer_size=4 * b * b * 8)
        ms = timed(lambda: cb.summa(args, A, B, Cm, None, g["cdt_row"], g["cdt_col"]))
        chk = check_block(Cm, b, gen(b, n, g["row"] * b, 0, n, 0) @ gen(n, b, 0, g["col"] * b, n, 1), n)
        report("config2: summa n=16384 2x2", 2.0 * n ** 3, ms, chk)
        timeline("config2 summa", lambda: cb.summa(args, A, B, Cm, None, g["cdt_row"], g["cdt_col"]))
        del A, B, Cm
        # ---- config 4a: bcast_cannon_4d as pure Cannon, n = 24576 ----
        n = sz(24576); d = cb.dcn_grid(world, 2); b = n // 2
        A, B, Cm = blocks(b, (d["y1"] * 2 + d["y2"]) * b, (d["x1"] * 2 + d["x2"]) * b, n)
        args = cb.ctb_args_t(n=n, lda_A=b, lda_B=b, lda_C=b, buffer_size=5 * b * b * 8, ovp=1)
        ms = timed(lambda: cb.bcast_cannon_4d(args, A, B, Cm, None, d["cdt_x1"], d["cdt_y1"], d["cdt_x2"], d["cdt_y2"]))
        row0, col0 = (d["y1"] * 2 + d["y2"]) * b, (d["x1"] * 2 + d["x2"]) * b
        chk = check_block(Cm, b, gen(b, n, row0, 0, n, 0) @ gen(n, b, 0, col0, n, 1), n)
        report("config4: bcast_cannon_4d (Cannon level) n=24576 2x2", 2.0 * n ** 3, ms, chk)
        timeline("config4 bcast_cannon_4d", lambda: cb.bcast_cannon_4d(args, A, B, Cm, None, d["cdt_x1"], d["cdt_y1"], d["cdt_x2"], d["cdt_y2"]))
        # ---- config 4b: split-dim Cannon kput, 12288^3 blocks ----
        px, py = rank % 2, rank // 2   # block column / block row of A, B and C alike (test/MM/test_spc.cxx:78-102)
        cb.fill_drand48(A, b, b, b, py * b, px * b, n, 0); cb.fill_drand48(B, b, b, b, py * b, px * b, n, 1)
        # (the reference permutes A and B in place, spcannon.cxx:76-77; this implementation works on internal copies)
        ms = timed(lambda: cb.kput_cannon(rank, 2, 2, world, b, b, b, "N", 1.0, A, "T", 0.0, B, Cm))
        # with transp_B = 'T' the local B buffer is the transpose of the rank's block of B: block (k, px) of B is G(k, px)^T
        ref = sum(gen(b, b, py * b, kk * b, n, 0) @ gen(b, b, kk * b, px * b, n, 1).T for kk in range(2))
        chk = check_block(Cm, b, ref, n)
        report("config4: kput_cannon 2-ary 2-cube, 12288^3 blocks", 2.0 * n ** 3, ms, chk)
        timeline("config4 kput_cannon", lambda: cb.kput_cannon(rank, 2, 2, world, b, b, b, "N", 1.0, A, "T", 0.0, B, Cm))
        del A, B, Cm
    # ---- config 5: CAQR trailing update ----
    m, ncol, k = sz(65536), sz(8192), sz(512)
    nprow, npcol = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[ws]
    prow, pcol = rank % nprow, rank // nprow
    ccol = cb.setup_sub_comm(world, prow, pcol, nprow) if ws > 1 else None   # ranks sharing my grid column
    mb, kb = m // nprow, ncol // npcol
    Y = torch.empty(mb * k, dtype=torch.float64, device="cuda"); Am = torch.empty(mb * kb, dtype=torch.float64, device="cuda")
    cb.fill_drand48(Y, mb, k, mb, prow * mb, 0, m, 0); cb.fill_drand48(Am, mb, kb, mb, prow * mb, pcol * kb, m, 1)
    Y.mul_(1.0 / 256.0)   # keep the update well scaled
    T = (torch.eye(k, dtype=torch.float64, device="cuda") + 0.01 * torch.tril(torch.rand(k, k, dtype=torch.float64, device="cuda"))).T.contiguous()
    ms = timed(lambda: cb.upd_A(Y, mb, Am, mb, mb, kb, k, T, ccol))

    def check_upd():
        """one application on a fresh A against W = T^-1 (Y^T A) and A - Y W computed by torch from regenerated operands"""
        cb.fill_drand48(Am, mb, kb, mb, prow * mb, pcol * kb, m, 1)
        cb.upd_A(Y, mb, Am, mb, mb, kb, k, T, ccol)
        torch.cuda.synchronize()
        Yf = gen(m, k, 0, 0, m, 0) * (1.0 / 256.0)
        W = Yf.T @ gen(m, kb, 0, pcol * kb, m, 1)
        W = torch.linalg.solve_triangular(T.view(k, k).T, W, upper=False)   # T is stored column-major, lower triangular
        ref = gen(mb, kb, prow * mb, pcol * kb, m, 1) - Yf[prow * mb:(prow + 1) * mb] @ W
        got = Am.view(kb, mb).T
        err = maxr(float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref)))
        tol = 10 * m * EPS
        return {"rel_frobenius_vs_cublas_crosscheck": err, "tolerance_10_k_eps": tol, "check_passed": bool(err <= tol)}

    chk5 = check_upd()
    timeline("config5 upd_A", lambda: cb.upd_A(Y, mb, Am, mb, mb, kb, k, T, ccol))
    report(f"config5: upd_A m=65536 n=8192 k=512 on {nprow}x{npcol}", 2 * 2.0 * m * ncol * k, ms,
           dict(chk5, note="flops = the two GEMMs (SURVEY §8d); all-reduce of W and the triangular solve are inside the time"))
    # the same with round 1's triangular solve (a block barrier per row of T instead of one warp per right-hand side)
    cb.lib().candmc_set_trsm_variant(0)
    ms = timed(lambda: cb.upd_A(Y, mb, Am, mb, mb, kb, k, T, ccol))
    chk5 = check_upd()
    cb.lib().candmc_set_trsm_variant(1)
    report(f"config5: upd_A m=65536 n=8192 k=512 on {nprow}x{npcol}, trsm variant 0", 2 * 2.0 * m * ncol * k, ms,
           dict(chk5, note="candmc_set_trsm_variant(0): round 1's kernel"))
    # 1-GPU local GEMM roofline at the Cannon block size (config 4, second half)
    if ws == 1:
        for n in (sz(12288), sz(16384)):
            A, B, Cm = blocks(n, 0, 0, n)
            ms = timed(lambda: cb.cdgemm("N", "N", n, n, n, 1.0, A, n, B, n, 0.0, Cm, n))
            report(f"config4: 1-GPU local GEMM n={n}", 2.0 * n ** 3, ms)
            del A, B, Cm
    if "--pending" in sys.argv:
        pending(world, rank, ws, timed, report, ccol, Y, Am, mb, kb, k)
    world.free()
    if ws > 1:
        dist.destroy_process_group()


def pending(world, rank, ws, timed, report, ccol, Y, Am, mb, kb, k):
    import numpy as np
    HBM = 6456.2  # MEASURED_PEAKS.json hbm_gbs
    # ---- N1: Yamamoto form at the config-5 shape (three GEMMs: 2*mb*kb*k twice + 2*k*k*kb) ----
    T = torch.rand(k * k, dtype=torch.float64, device="cuda") * (1.0 / k)
    ms = timed(lambda: cb.upd_Yamamoto_A(Y, mb, Am, mb, mb, kb, k, T, ccol))
    report("N1: upd_Yamamoto_A m=65536 n=8192 k=512", ws * (2 * 2.0 * mb * kb * k + 2.0 * k * k * kb), ms)
    # ---- N3: permute kernels alone (one GPU plays rank 1 of a 4-rank axis), 8192 x 8192 local piece, nb = 64 ----
    rows = cols = sz(8192); nb = sz(64)
    X = torch.rand(rows * cols, dtype=torch.float64, device="cuda"); S = torch.empty_like(X)
    st = torch.cuda.current_stream().cuda_stream
    for rows_axis in (1, 0):
        for gather in (1, 0):
            fn = lambda: cb.lib().candmc_debug_redist_permute(4, 1, 0, rows // nb, nb, rows_axis, gather, X.data_ptr(), rows,  # noqa: E731
                                                              S.data_ptr(), rows, cols, st)
            ms = timed(fn, steps=10)
            if rank == 0:
                gbs = 16.0 * rows * cols / (ms * 1e-3) / 1e9
                print(json.dumps({"config": f"N3: permute_blocks_kernel rows_axis={rows_axis} gather={gather} 8192x8192 nb=64",
                                  "ms": ms, "GBps": gbs, "pct_of_hbm_copy_peak": 100 * gbs / HBM}), flush=True)
    del X, S
    # ---- N3: the whole redistribution over NCCL (n = 16384 * sqrt-ish grid) ----
    nprow = {1: 1, 2: 2, 4: 2, 8: 4}[ws]; npcol = ws // nprow
    myrow, mycol = rank % nprow, rank // nprow
    crow = cb.setup_sub_comm(world, mycol, myrow, npcol); ccol2 = cb.setup_sub_comm(world, myrow, mycol, nprow)
    pv = cb.pview(0, 0, crow, ccol2, world)
    m = rows * nprow; n = cols * npcol
    src = torch.rand(rows * cols, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
    ms = timed(lambda: cb.cyclic_to_blocked(m, n, nb, src, rows, dst, rows, pv))
    if rank == 0:
        print(json.dumps({"config": f"N3: cyclic_to_blocked {m}x{n} nb=64 on {nprow}x{npcol}", "ms": ms,
                          "GBps_per_gpu_algorithmic": 16.0 * rows * cols / (ms * 1e-3) / 1e9}), flush=True)
    crow.free(); ccol2.free()
    del src, dst
    # ---- N4: sym_full2band trailing update, first level of n = 16384 * q with band 512 (2 GEMMs: 2 * 2 * mb * kb * b flop) ----
    pr = {1: 1, 4: 2}.get(ws)
    if pr:
        n4, b4, bs4 = sz(16384) * pr, sz(512), sz(128)
        myrow, mycol = rank % pr, rank // pr
        crow = cb.setup_sub_comm(world, mycol, myrow, pr); ccol3 = cb.setup_sub_comm(world, myrow, mycol, pr)
        import ctypes
        if myrow == mycol:
            cdiag = cb.setup_sub_comm(world, myrow, 0, pr)
        else:
            h = ctypes.c_void_p(); cb._lib.check(cb.lib().candmc_comm_split(world.cm, 1, rank, ctypes.byref(h)))
            cdiag = cb.CommData_t(cm=h.value, np=ws - pr, rank=-1)
        ro, co, mb4, kb4 = cb.sym_full2band_extents(n4, b4, bs4, pr, myrow, mycol, 0, 0)
        A4 = torch.rand((n4 // pr) ** 2, dtype=torch.float64, device="cuda")
        Y4 = torch.rand(mb4 * b4, dtype=torch.float64, device="cuda") - 0.5
        pv4 = cb.pview(0, 0, crow, ccol3, world)
        ms = timed(lambda: cb.sym_full2band_update(A4, n4 // pr, n4, b4, bs4, pv4, cdiag if myrow == mycol else None, Y4, mb4))
        report(f"N4: sym_full2band_update n={n4} b={b4} on {pr}x{pr}", ws * 2 * 2.0 * mb4 * kb4 * b4, ms,
               {"note": "flops = W = Y^T A and U V'; invT, Z, the triangular solve, three broadcasts, the partner exchange and the "
                        "HBM-bound rank-2b update (32 B per trailing element) are inside the time"})
        crow.free(); ccol3.free(); cdiag.free()
        del A4, Y4
    # ---- N2: LU seam, one trailing-update step on a 16384^2 local matrix, panel width 512 ----
    from candmc_b200 import lu_offload as lo
    nloc, kp = sz(16384), sz(512)
    lo.alloc_A(nloc * nloc, None); lo.alloc_L(nloc * kp); lo.alloc_U(kp * nloc); lo.alloc_transfer(kp * nloc)
    hA = np.random.rand(nloc * kp) - 0.5
    lo.upload_lda_cpy(nloc, kp, nloc, nloc, hA, 0, lo.OFF_L); lo.upload_lda_cpy(kp, nloc, kp, kp, hA, 0, lo.OFF_U)
    for c0 in range(0, nloc, kp):
        lo.upload_lda_cpy(nloc, kp, nloc, nloc, hA, c0 * nloc, lo.OFF_A)
    panel = np.empty((nloc - kp) * kp)
    import time
    for overlap in (1, 0):
        lo.set_overlap(overlap)
        best = 1e9
        for _ in range(3):
            lo.sync(); t0 = time.perf_counter()
            lo.offload_gemm_A("N", "N", nloc - kp, nloc - kp, kp, -1.0, 0, lo.OFF_L, nloc, 0, lo.OFF_U, kp, 1.0, kp * nloc + kp, lo.OFF_A, nloc)
            t_issue = time.perf_counter() - t0
            lo.download_lda_cpy(nloc - kp, kp, nloc, nloc - kp, kp * nloc + kp, panel, lo.OFF_A)   # waits for the GEMM by itself
            best = min(best, time.perf_counter() - t0)
        if rank == 0:
            fl = 2.0 * (nloc - kp) ** 2 * kp
            print(json.dumps({"config": f"N2: LU seam step, local 16384^2, panel 512, overlap={overlap}", "ms_gemm_plus_panel_download": best * 1e3,
                              "tflops_incl_download": fl / best / 1e12, "ms_host_blocked_issuing_gemm": t_issue * 1e3,
                              "timing": "host wall clock around the seam calls (the download is synchronous by contract)"}), flush=True)
    lo.free_offload_A(); lo.free_offload_L(); lo.free_offload_U(); lo.free_offload_transfer()


if __name__ == "__main__":
    main()
