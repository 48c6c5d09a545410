"""CPU tests of the block-cyclic <-> blocked redistribution.

 * the oracle's layout maps against the reference's own generators: block-cyclic = test/QR/test_qr_2d.cxx:87-94,
   blocked = test/MM/topo_pdgemm_unit.cxx:250-256 (restated here independently in numpy);
 * the index plan the CUDA path uses (candmc_redist_axis_plan / candmc_redist_strided_index are pure host code of the
   product library, the very functions the kernels call) against brute force;
 * the whole two-exchange algorithm of candmc_b200/csrc/redist.cu, emulated rank by rank in numpy ON TOP OF those plan
   functions (segments, offsets, contiguous ranges, strided scatter), against the oracle — so everything but the kernel
   launches and the NCCL calls is exercised without a GPU.
"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle():
    L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    L.oracle_redistribute.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    return L


def oracle_redistribute(to_cyclic, m, n, nb, nprow, npcol, rrow, rcol, pieces):
    L = _oracle()
    P = nprow * npcol
    outs = [np.full_like(p, np.nan) for p in pieces]
    pin = (C.c_void_p * P)(*[p.ctypes.data for p in pieces])
    pout = (C.c_void_p * P)(*[o.ctypes.data for o in outs])
    assert L.oracle_redistribute(to_cyclic, m, n, nb, nprow, npcol, rrow, rcol, pin, pout) == 0
    return outs


def cyclic_pieces(G, nb, nprow, npcol, rrow=0, rcol=0):
    """the reference's block-cyclic generator (test_qr_2d.cxx:87-94) applied to a global matrix; rank = myrow + mycol*nprow"""
    m, n = G.shape
    rows, cols = m // nprow, n // npcol
    out = []
    for mycol in range(npcol):
        for myrow in range(nprow):
            r = np.arange(rows)
            c = np.arange(cols)
            gr = ((myrow - rrow) % nprow) * nb + r % nb + (r // nb) * nb * nprow
            gc = ((mycol - rcol) % npcol) * nb + c % nb + (c // nb) * nb * npcol
            out.append(np.asfortranarray(G[np.ix_(gr, gc)]).reshape(-1, order="F").copy())
    return out


def blocked_pieces(G, nprow, npcol):
    m, n = G.shape
    rows, cols = m // nprow, n // npcol
    return [np.asfortranarray(G[i * rows:(i + 1) * rows, j * cols:(j + 1) * cols]).reshape(-1, order="F").copy()
            for j in range(npcol) for i in range(nprow)]


CASES = [(24, 24, 2, 2, 2, 0, 0), (24, 36, 2, 2, 3, 0, 0), (36, 24, 3, 3, 2, 1, 1), (48, 40, 4, 3, 2, 2, 0),
         (30, 30, 5, 3, 3, 0, 2), (16, 16, 4, 1, 4, 0, 3), (64, 64, 2, 4, 4, 3, 1), (12, 12, 1, 2, 2, 1, 0)]


@pytest.mark.parametrize("m,n,nb,nprow,npcol,rrow,rcol", CASES)
def test_oracle_layout_maps(m, n, nb, nprow, npcol, rrow, rcol):
    G = np.arange(m * n, dtype=np.float64).reshape(m, n) + 0.25
    cyc = cyclic_pieces(G, nb, nprow, npcol, rrow, rcol)
    blk = blocked_pieces(G, nprow, npcol)
    got_blk = oracle_redistribute(0, m, n, nb, nprow, npcol, rrow, rcol, cyc)
    got_cyc = oracle_redistribute(1, m, n, nb, nprow, npcol, rrow, rcol, blk)
    for a, b in zip(got_blk, blk):
        assert np.array_equal(a, b)
    for a, b in zip(got_cyc, cyc):
        assert np.array_equal(a, b)


# ---- the product library's plan ------------------------------------------------------------------------------------
def axis_plan(P, me, root, K, nb):
    from candmc_b200 import lib
    arr = [(C.c_int * P)() for _ in range(4)]
    assert lib().candmc_redist_axis_plan(P, me, root, K, nb, *arr) == 0
    return [list(a) for a in arr]  # lo, ccnt, first, scnt


def strided_index(P, me, root, K, nb, rows_axis, blk, w, o, other):
    from candmc_b200 import lib
    idx, peer = C.c_int64(), C.c_int()
    assert lib().candmc_redist_strided_index(P, me, root, K, nb, int(rows_axis), blk, w, o, other, C.byref(idx),
                                             C.byref(peer)) == 0
    return idx.value, peer.value


def test_axis_plan_against_brute_force():
    for P in range(1, 7):
        for K in range(0, 14):
            for root in range(P):
                plans = [axis_plan(P, me, root, K, 2) for me in range(P)]
                for me in range(P):
                    lo, ccnt, first, scnt = plans[me]
                    mc = (me - root) % P
                    # brute force: global block I -> cyclic owner (I + root) % P at local I // P; blocked owner I // K at I % K
                    for p in range(P):
                        mine_cyc = [I // P for I in range(P * K) if (I + root) % P == me and I // K == p]
                        assert mine_cyc == list(range(lo[p], lo[p] + ccnt[p])), (P, K, root, me, p)
                        mine_blk = [I % K for I in range(P * K) if I // K == me and (I + root) % P == p]
                        assert mine_blk == [first[p] + t * P for t in range(scnt[p])], (P, K, root, me, p)
                        # both ends of an exchange agree on its size and on the identity of every block in it
                        assert ccnt[p] == plans[p][3][me]
                        for t in range(ccnt[p]):
                            I_sender = (lo[p] + t) * P + mc
                            I_receiver = p * K + plans[p][2][me] + t * P
                            assert I_sender == I_receiver


def test_strided_index_is_a_bijection_onto_the_segments():
    for (P, me, root, K, nb, other) in [(2, 1, 0, 4, 2, 6), (3, 0, 2, 5, 3, 4), (4, 2, 1, 4, 1, 3), (3, 1, 0, 2, 2, 5)]:
        lo, ccnt, first, scnt = axis_plan(P, me, root, K, nb)
        for rows_axis in (True, False):
            seen = {}
            for blk in range(K):
                for w in range(nb):
                    for o in range(other):
                        idx, peer = strided_index(P, me, root, K, nb, rows_axis, blk, w, o, other)
                        assert idx not in seen
                        seen[idx] = peer
            assert sorted(seen) == list(range(K * nb * other))
            # segments are laid out in peer order, each of its advertised size
            start = 0
            for p in range(P):
                size = scnt[p] * nb * other
                assert all(seen[i] == p for i in range(start, start + size))
                start += size


# ---- emulation of redist.cu on top of the plan ----------------------------------------------------------------------
def _axis_exchange(rows_axis, to_blocked, P, root, nb, mats, rows, cols):
    """mats[me]: rows x cols numpy (F order) pieces of the P ranks on this axis -> list of output pieces.  Mirrors
    axis_exchange() in candmc_b200/csrc/redist.cu: segment buffers, offsets in doubles, plan-driven copies."""
    if P == 1:
        return [mats[0].copy()]
    other = cols if rows_axis else rows
    K = (rows if rows_axis else cols) // nb
    per_block = nb * other
    plans = [axis_plan(P, me, root, K, nb) for me in range(P)]

    def coff(pl):  # exclusive prefix sums in blocks
        return np.concatenate([[0], np.cumsum(pl[1])[:-1]]), np.concatenate([[0], np.cumsum(pl[3])[:-1]])

    def strided_scatter_or_gather(me, X, seg, gather):
        for c in range(cols):
            for r in range(rows):
                if rows_axis:
                    idx, _ = strided_index(P, me, root, K, nb, True, r // nb, r % nb, c, cols)
                else:
                    idx, _ = strided_index(P, me, root, K, nb, False, c // nb, c % nb, r, rows)
                if gather:
                    seg[idx] = X[r, c]
                else:
                    X[r, c] = seg[idx]

    outs = [np.full((rows, cols), np.nan, order="F") for _ in range(P)]
    sbufs = [np.full(rows * cols, np.nan) for _ in range(P)]
    rbufs = [np.full(rows * cols, np.nan) for _ in range(P)]
    if to_blocked:
        for me in range(P):  # pack contiguous ranges of the cyclic input
            lo, ccnt, _, _ = plans[me]
            co, _ = coff(plans[me])
            for p in range(P):
                ext, at = ccnt[p] * nb, lo[p] * nb
                blk = mats[me][at:at + ext, :] if rows_axis else mats[me][:, at:at + ext]
                sbufs[me][co[p] * per_block: co[p] * per_block + blk.size] = blk.reshape(-1, order="F")
        for me in range(P):  # all-to-all: segment `me` of p's send buffer -> segment p of my receive buffer
            _, so = coff(plans[me])
            for p in range(P):
                cop, _ = coff(plans[p])
                cnt = plans[p][1][me] * per_block
                assert cnt == plans[me][3][p] * per_block
                rbufs[me][so[p] * per_block: so[p] * per_block + cnt] = sbufs[p][cop[me] * per_block: cop[me] * per_block + cnt]
        for me in range(P):
            strided_scatter_or_gather(me, outs[me], rbufs[me], gather=False)
        return outs
    for me in range(P):
        strided_scatter_or_gather(me, mats[me], sbufs[me], gather=True)
    for me in range(P):
        co, _ = coff(plans[me])
        for p in range(P):
            _, sop = coff(plans[p])
            cnt = plans[p][3][me] * per_block
            assert cnt == plans[me][1][p] * per_block
            rbufs[me][co[p] * per_block: co[p] * per_block + cnt] = sbufs[p][sop[me] * per_block: sop[me] * per_block + cnt]
    for me in range(P):
        lo, ccnt, _, _ = plans[me]
        co, _ = coff(plans[me])
        for p in range(P):
            ext, at = ccnt[p] * nb, lo[p] * nb
            seg = rbufs[me][co[p] * per_block: co[p] * per_block + ext * other]
            if rows_axis:
                outs[me][at:at + ext, :] = seg.reshape(ext, cols, order="F")
            else:
                outs[me][:, at:at + ext] = seg.reshape(rows, ext, order="F")
    return outs


def emulate_redistribute(to_cyclic, m, n, nb, nprow, npcol, rrow, rcol, pieces):
    rows, cols = m // nprow, n // npcol
    mats = {(i, j): pieces[i + j * nprow].reshape(rows, cols, order="F").copy() for i in range(nprow) for j in range(npcol)}

    def rows_phase(to_blocked):
        for j in range(npcol):  # over ccol: the ranks of grid column j
            outs = _axis_exchange(True, to_blocked, nprow, rrow, nb, [mats[(i, j)] for i in range(nprow)], rows, cols)
            for i in range(nprow):
                mats[(i, j)] = outs[i]

    def cols_phase(to_blocked):
        for i in range(nprow):  # over crow: the ranks of grid row i
            outs = _axis_exchange(False, to_blocked, npcol, rcol, nb, [mats[(i, j)] for j in range(npcol)], rows, cols)
            for j in range(npcol):
                mats[(i, j)] = outs[j]

    if not to_cyclic:
        rows_phase(True)
        cols_phase(True)
    else:
        cols_phase(False)
        rows_phase(False)
    return [mats[(i, j)].reshape(-1, order="F") for j in range(npcol) for i in range(nprow)]


@pytest.mark.parametrize("m,n,nb,nprow,npcol,rrow,rcol", CASES[:6])
def test_two_phase_algorithm_matches_oracle(m, n, nb, nprow, npcol, rrow, rcol):
    rng = np.random.RandomState(m * 7 + n)
    G = rng.rand(m, n)
    cyc = cyclic_pieces(G, nb, nprow, npcol, rrow, rcol)
    blk = blocked_pieces(G, nprow, npcol)
    got_blk = emulate_redistribute(0, m, n, nb, nprow, npcol, rrow, rcol, cyc)
    want_blk = oracle_redistribute(0, m, n, nb, nprow, npcol, rrow, rcol, cyc)
    for a, b, c in zip(got_blk, want_blk, blk):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    got_cyc = emulate_redistribute(1, m, n, nb, nprow, npcol, rrow, rcol, blk)
    for a, b in zip(got_cyc, cyc):
        assert np.array_equal(a, b)


def test_redistribute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU box")
    import candmc_b200 as cb
    from candmc_b200._lib import PView

    pv = PView(0, 0, None, None, None)
    assert cb.lib().candmc_redistribute(0, 8, 8, 2, 0, 4, 0, 4, C.byref(pv), None) != 0
