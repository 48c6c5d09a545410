"""Randomised parity cases on top of tests/dist_worker.py's case functions (same checkers: oracle + tolerances).
Runs on GPUs (`torchrun --nproc-per-node N tests/fuzz_worker.py <seed> <count>`) or, with CANDMC_CPUSIM=1, on the CPU
functional simulator (tests/cpusim).  Every rank draws the same parameters from the seed.  Prints one JSON line on rank 0."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dist_worker as dw  # noqa: E402  (installs the simulator hooks when CANDMC_CPUSIM=1)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

cb = dw.cb


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rank = int(os.environ.get("RANK", 0)); P = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if P > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = cb.init_world(rank, P, local)
    if os.environ.get("CANDMC_TEST_PANEL_TRANSPORT") == "1":
        cb.lib().candmc_set_panel_transport(1)
    if os.environ.get("CANDMC_TEST_FUSED_GRIDS") == "1":
        cb.lib().candmc_set_fused_reduce(2)
    golden = {}
    rng = random.Random(seed * 1000 + P)
    log = []
    for it in range(count):
        kinds = ["d25", "upd_A", "update_A", "redist", "yamamoto"]
        if int(round(P ** 0.5)) ** 2 == P:
            kinds += ["summa", "spc", "f2b"]
        if P in (4, 16):
            kinds += ["dcn"]
        kind = rng.choice(kinds)
        cb.set_min_kchunk(rng.choice([1024, 8, 2, 16]))
        cb.lib().candmc_set_merge_panels(rng.choice([0, 1, 2, 3]))
        cb.lib().candmc_set_panel_transport(rng.choice([1, 1, 0]))   # copy engines (the default) or the NCCL fallback; same on every rank
        pad = rng.choice([0, 0, 1, 2, 3])
        tag = f"fz{seed}.{it}.{kind}"
        if os.environ.get("FUZZ_TRACE") == "1":
            print(f"[rank {rank}] {tag} (previous: {log[-1] if log else None})", file=sys.stderr, flush=True)
        if kind == "d25":
            cs = [c for c in (1, 2, 4) if P % c == 0 and int(round((P // c) ** 0.5)) ** 2 == P // c
                  and (int(round((P // c) ** 0.5)) % c == 0 or P // c == 1)]
            if not cs:
                continue
            c = rng.choice(cs)
            q = int(round((P // c) ** 0.5))
            b = rng.choice([4, 6, 10, 16, 24, 32, 48, 64, 128, 256]) * (c if q == 1 else 1)
            host = rng.random() < 0.4
            if host:
                cb.lib().candmc_set_host_pipeline_min(rng.choice([16, 2048]))
                cb.lib().candmc_set_early_c_download(rng.choice([0, 1]))
                cb.lib().candmc_set_skip_unused_uploads(rng.choice([0, 1]))
                cb.lib().candmc_set_b_first_chunk_early(rng.choice([0, 1]))
                cb.lib().candmc_set_host_gather(rng.choice([0, 1, 1]))
                if rng.random() < 0.5:
                    host = "pinned"   # page-locked blocks: B gathered chunk-wise out of host memory (where the chunking allows)
            # trans flags: on grids they reach the local multiply only (blocks as stored, the oracle follows the reference there);
            # on 1 x 1 x c the stored transposes are generated so that op(stored) is the same operand (even k-slices)
            tr = rng.choice([("N", "N"), ("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")]) if (q > 1 or (c > 1 and b % (2 * c) == 0)) else ("N", "N")
            log.append((tag, dict(n=b * q, c=c, pad=pad, host=host, trans=tr)))
            dw.case_d25(world, golden, tag, b * q, c, rng.choice([0, 1]), lda_pad=pad, use_host=host, check_golden=False, trans=tr)
            cb.lib().candmc_set_host_pipeline_min(2048); cb.lib().candmc_set_early_c_download(1); cb.lib().candmc_set_skip_unused_uploads(1)
            cb.lib().candmc_set_host_gather(1)
            cb.lib().candmc_set_b_first_chunk_early(0)
        elif kind == "summa":
            q = int(round(P ** 0.5))
            b = rng.choice([3, 4, 8, 10, 16, 32, 64])
            log.append((tag, dict(n=b * q, pad=pad)))
            dw.case_summa(world, golden, tag, b * q, lda_pad=pad, trans=rng.choice([("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")]))
        elif kind == "dcn":
            x2 = rng.choice([1, 2] if P == 4 else [1, 2, 4])
            x1 = int(round(P ** 0.5)) // x2
            b = rng.choice([4, 8, 12, 32])
            log.append((tag, dict(n=b * x1 * x2, x2=x2)))
            dw.case_dcn(world, golden, tag, b * x1 * x2, x2, rng.choice([0, 1]), lda_pad=pad,
                        trans=rng.choice([("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")]))
        elif kind == "spc":
            ndim = rng.choice([2, 4]) if P == 16 else 2
            kary = int(round(P ** (1.0 / ndim)))
            if kary ** ndim != P:
                continue
            n, m, k = rng.choice([4, 8, 20, 32]), rng.choice([4, 12, 24, 40]), rng.choice([4, 8, 16, 36])   # k % ndim == 0 (spcannon.cxx)
            log.append((tag, dict(kary=kary, ndim=ndim, n=n, m=m, k=k)))
            dw.case_spc(world, golden, tag, rng.choice([0, 1]), kary, ndim, n, m, k, rng.choice(["N", "T"]), use_host=rng.random() < 0.3)
        elif kind == "upd_A":
            mb, kb, b = rng.choice([8, 24, 64, 96]), rng.choice([4, 16, 40, 80]), rng.choice([2, 4, 8, 16, 32])
            form = dict(use_host=rng.random() < 0.4, t_from_y=rng.random() < 0.4)   # host operands staged inside; T formed from Y
            if form["use_host"]:
                form["lda_pad"] = pad
            log.append((tag, dict(mb=mb, kb=kb, b=b, **form)))
            dw.case_upd_A(world, tag, mb, kb, b, **form)
        elif kind in ("update_A", "yamamoto", "redist"):
            nprow = rng.choice([d for d in range(1, P + 1) if P % d == 0])
            npcol = P // nprow
            b = rng.choice([2, 4, 8, 16])
            rrow, rcol = rng.randrange(nprow), rng.randrange(npcol)
            if kind == "redist":
                m, n = b * nprow * rng.choice([1, 2, 3, 5]), b * npcol * rng.choice([1, 2, 4])
                log.append((tag, dict(m=m, n=n, nb=b, grid=(nprow, npcol), roots=(rrow, rcol), pad=pad)))
                dw.case_redistribute(world, tag, m, n, b, nprow, rrow, rcol, pad=pad)
            else:
                m, k = b * rng.choice([2, 3, 5, 8, 11]), b * rng.choice([1, 2, 4, 7])
                log.append((tag, dict(m=m, k=k, b=b, grid=(nprow, npcol), roots=(rrow, rcol))))
                if kind == "update_A":
                    dw.case_update_A(world, golden, tag, m, k, b, nprow, rrow, rcol, with_T=rng.random() < 0.3, use_host=rng.random() < 0.4)
                else:
                    dw.case_update_Yamamoto_A(world, golden, tag, m, k, b, nprow, rrow, rcol, use_host=rng.random() < 0.4)
        elif kind == "f2b":
            pr = int(round(P ** 0.5))
            bs = rng.choice([2, 4, 8, 16]); b = bs * pr * rng.choice([1, 2, 4]); n = b + bs * pr * rng.choice([1, 2, 3, 6])
            log.append((tag, dict(n=n, b=b, b_sub=bs)))
            dw.case_f2b_big(world, tag, n, b, bs)
    cb.set_min_kchunk(1024)
    fails = [r for r in dw.RESULTS if not r[1]]
    flag = torch.tensor([len(fails)], dtype=torch.int64, device="cuda")
    if P > 1:
        dist.all_reduce(flag)
    for name, ok, err, tol in fails:
        print(f"[rank {rank}] FAIL {name}: err={err:.3e} tol={tol:.3e}", flush=True)
    if rank == 0:
        if flag.item():
            print("cases:", log, flush=True)
        print(json.dumps({"world_size": P, "seed": seed, "cases": len(log), "checks_rank0": len(dw.RESULTS),
                          "failed_all_ranks": int(flag.item()), "max_err_rank0": max((r[2] for r in dw.RESULTS), default=0.0),
                          "panel_transport_sends_rank0": int(cb.lib().candmc_panel_transport_sends())}), flush=True)
    world.free()
    if P > 1:
        dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
