"""GPU worker for tests/test_zz_redist_gpu.py (own process).  One GPU plays every rank of the grid: the permute kernels
(candmc_debug_redist_permute) and the contiguous packs (candmc_lda_cpy) run on the device exactly as candmc_redistribute
launches them; only the NCCL all-to-all is replaced by device copies between the per-rank buffers.  Checked bit-exactly
against the oracle.  Prints one JSON line.

    python tests/redist_worker.py <m> <n> <nb> <nprow> <npcol> <rrow> <rcol>
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

if os.environ.get("CANDMC_CPUSIM") == "1":   # CPU suite: the same worker on the functional simulator (tests/cpusim)
    sys.path.insert(0, os.path.join(HERE, "cpusim"))
    import simtorch
    simtorch.install()

import torch  # noqa: E402

import candmc_b200 as cb  # noqa: E402
from test_redist import axis_plan, blocked_pieces, cyclic_pieces, oracle_redistribute  # noqa: E402


def axis_exchange_gpu(rows_axis, to_blocked, P, root, nb, mats, rows, cols):
    """mats: list of P device tensors holding rows x cols column-major pieces (ld = rows)"""
    L = cb.lib()
    if P == 1:
        return [mats[0].clone()]
    other = cols if rows_axis else rows
    K = (rows if rows_axis else cols) // nb
    per_block = nb * other
    plans = [axis_plan(P, me, root, K, nb) for me in range(P)]
    pre = lambda v: np.concatenate([[0], np.cumsum(v)[:-1]]).astype(np.int64)  # noqa: E731
    co = [pre(pl[1]) for pl in plans]
    so = [pre(pl[3]) for pl in plans]
    sb = [torch.full((rows * cols,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(P)]
    rb = [torch.full((rows * cols,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(P)]
    outs = [torch.full((rows * cols,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(P)]
    st = torch.cuda.current_stream().cuda_stream

    def permute(me, X, seg, gather):
        rc = L.candmc_debug_redist_permute(P, me, root, K, nb, int(rows_axis), int(gather), X.data_ptr(), rows,
                                           seg.data_ptr(), rows, cols, st)
        assert rc == 0, cb.last_error()

    def lda(nrow, ncol, lda_a, lda_b, a_ptr, b_ptr):
        assert L.candmc_lda_cpy(nrow, ncol, lda_a, lda_b, a_ptr, b_ptr, st) == 0, cb.last_error()

    if to_blocked:
        for me in range(P):
            lo, ccnt = plans[me][0], plans[me][1]
            for p in range(P):
                if ccnt[p] == 0:
                    continue
                ext, at = ccnt[p] * nb, lo[p] * nb
                dstp = sb[me].data_ptr() + 8 * int(co[me][p]) * per_block
                if rows_axis:
                    lda(ext, cols, rows, ext, mats[me].data_ptr() + 8 * at, dstp)
                else:
                    lda(rows, ext, rows, rows, mats[me].data_ptr() + 8 * at * rows, dstp)
        for me in range(P):
            for p in range(P):
                cnt = plans[p][1][me] * per_block
                s0, r0 = int(co[p][me]) * per_block, int(so[me][p]) * per_block
                rb[me][r0:r0 + cnt] = sb[p][s0:s0 + cnt]
        for me in range(P):
            permute(me, outs[me], rb[me], gather=False)
        return outs
    for me in range(P):
        permute(me, mats[me], sb[me], gather=True)
    for me in range(P):
        for p in range(P):
            cnt = plans[p][3][me] * per_block
            s0, r0 = int(so[p][me]) * per_block, int(co[me][p]) * per_block
            rb[me][r0:r0 + cnt] = sb[p][s0:s0 + cnt]
    for me in range(P):
        lo, ccnt = plans[me][0], plans[me][1]
        for p in range(P):
            if ccnt[p] == 0:
                continue
            ext, at = ccnt[p] * nb, lo[p] * nb
            srcp = rb[me].data_ptr() + 8 * int(co[me][p]) * per_block
            if rows_axis:
                lda(ext, cols, ext, rows, srcp, outs[me].data_ptr() + 8 * at)
            else:
                lda(rows, ext, rows, rows, srcp, outs[me].data_ptr() + 8 * at * rows)
    return outs


def redistribute_gpu(to_cyclic, m, n, nb, nprow, npcol, rrow, rcol, pieces):
    rows, cols = m // nprow, n // npcol
    mats = {(i, j): torch.from_numpy(pieces[i + j * nprow]).cuda() for i in range(nprow) for j in range(npcol)}

    def rows_phase(tb):
        for j in range(npcol):
            outs = axis_exchange_gpu(True, tb, nprow, rrow, nb, [mats[(i, j)] for i in range(nprow)], rows, cols)
            for i in range(nprow):
                mats[(i, j)] = outs[i]

    def cols_phase(tb):
        for i in range(nprow):
            outs = axis_exchange_gpu(False, tb, npcol, rcol, nb, [mats[(i, j)] for j in range(npcol)], rows, cols)
            for j in range(npcol):
                mats[(i, j)] = outs[j]

    if not to_cyclic:
        rows_phase(True); cols_phase(True)
    else:
        cols_phase(False); rows_phase(False)
    torch.cuda.synchronize()
    return [mats[(i, j)].cpu().numpy() for j in range(npcol) for i in range(nprow)]


def main():
    m, n, nb, nprow, npcol, rrow, rcol = map(int, sys.argv[1:8])
    cb.init()
    before = cb.launch_count()
    rng = np.random.RandomState(m + 3 * n + nb)
    G = rng.rand(m, n)
    cyc = cyclic_pieces(G, nb, nprow, npcol, rrow, rcol)
    blk = blocked_pieces(G, nprow, npcol)
    got_blk = redistribute_gpu(0, m, n, nb, nprow, npcol, rrow, rcol, cyc)
    got_cyc = redistribute_gpu(1, m, n, nb, nprow, npcol, rrow, rcol, blk)
    want_blk = oracle_redistribute(0, m, n, nb, nprow, npcol, rrow, rcol, cyc) if m * n <= 1 << 24 else blk
    res = {"case": [m, n, nb, nprow, npcol, rrow, rcol],
           "to_blocked_exact": all(np.array_equal(a, b) for a, b in zip(got_blk, want_blk)),
           "to_blocked_matches_generator": all(np.array_equal(a, b) for a, b in zip(got_blk, blk)),
           "to_cyclic_exact": all(np.array_equal(a, b) for a, b in zip(got_cyc, cyc)),
           "launches": int(cb.launch_count() - before)}
    # the public entry point on the one grid a single GPU has: 1 x 1 (both exchanges degenerate to copies)
    world = cb.init_world(0, 1, 0)
    pv = cb.pview(0, 0, world, world, world)
    src = torch.from_numpy(G.reshape(-1, order="F").copy()).cuda()
    dst = torch.zeros_like(src)
    cb.cyclic_to_blocked(m, n, nb, src, m, dst, m, pv)
    torch.cuda.synchronize()
    res["single_rank_identity"] = bool(torch.equal(src, dst))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
