"""CPU tests of the host-side grid logic (the reference's communicator macros, alg/shared/comm.h:143-195) and of the
N > 1 bootstrap plumbing under gloo (world size 2)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split(ranks_ck):
    """MPI_Comm_split semantics: {color: [world ranks ordered by (key, rank)]}."""
    groups = {}
    for r, (color, key) in ranks_ck.items():
        groups.setdefault(color, []).append((key, r))
    return {c: [r for _, r in sorted(v)] for c, v in groups.items()}


@pytest.mark.parametrize("P,c", [(1, 1), (4, 1), (8, 2), (9, 1), (16, 1), (32, 2), (2, 2)])
def test_d25_grid_matches_reference_rank_order(P, c):
    from candmc_b200 import grid

    q, cc = grid.grid_shape_for(P, c)
    assert cc == c and q * q * c == P
    # RSETUP_KDIR_COMM: colour = r % (P/c), key = r / (P/c)
    kdir = split({r: grid.kdir_color_key(r, P, c) for r in range(P)})
    for color, members in kdir.items():
        assert members == [color + l * (P // c) for l in range(c)]
    # RSETUP_LAYER_COMM: world rank r = layer*q*q + row*q + col
    for r in range(P):
        _, layer = grid.kdir_color_key(r, P, c)
        intra = r % (P // c)
        row, col = grid.layer_coords(intra, q)
        assert r == layer * q * q + row * q + col


def test_default_replication_matches_reference_driver():
    """bench/MM/topo_pdgemm_bench.cxx:448-456: c_rep = 2 only for non-square counts >= 8."""
    from candmc_b200 import grid

    assert grid.grid_shape_for(4) == (2, 1)
    assert grid.grid_shape_for(8) == (2, 2)
    assert grid.grid_shape_for(16) == (4, 1)
    assert grid.grid_shape_for(32) == (4, 2)
    assert grid.grid_shape_for(1) == (1, 1)
    assert grid.grid_shape_for(2) == (1, 2)   # this implementation's k-split extension
    from candmc_b200 import CandmcError

    with pytest.raises(CandmcError):
        grid.grid_shape_for(6)


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from candmc_b200 import grid
dist.init_process_group("gloo")
rank = dist.get_rank()
payload = bytes(range(128)) if rank == 0 else None
got = grid._torch_exchange(payload)
assert got == bytes(range(128)), got
# every rank derives the same grid from (rank, world size)
q, c = grid.grid_shape_for(dist.get_world_size())
color, key = grid.kdir_color_key(rank, dist.get_world_size(), c)
out = [None, None]
dist.all_gather_object(out, (rank, q, c, color, key))
assert out == [(0, 1, 2, 0, 0), (1, 1, 2, 0, 1)], out
# bench.py's timing reduction: max over ranks
import torch
t = torch.tensor([1.0 + rank])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 2.0
dist.destroy_process_group()
print("OK", rank)
"""


def test_bootstrap_exchange_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29431", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("OK") == 2


def _cut(n, k, panels):
    import ctypes as C

    from candmc_b200._lib import check, lib

    w, c = (C.c_int64 * 128)(), (C.c_int64 * 128)()
    nw, nc = C.c_int(), C.c_int()
    check(lib().candmc_host_pipeline_cut(n, k, panels, w, c, 128, C.byref(nw), C.byref(nc)))
    return list(w[: nw.value]), list(c[: nc.value])


def test_host_pipeline_cut_of_the_headline_size():
    """bench.py's one-GPU end-to-end leg (n = k = 32768): the automatic cut is the graduated one — first panel twice as wide as
    the others, its k-chunks starting at k/16 and growing by a tenth (chunk t+1's upload must fit under chunk t's multiply: fast
    growth makes the panel upload-bound, tools/host_pipeline_model.py), 1/32 of C behind the last multiply"""
    w, c = _cut(32768, 32768, 0)
    assert w == [8192] + [4096] * 5 + [2048, 1024, 1024]
    assert c[0] == 2048 and c[:4] == [2048, 2256, 2480, 2736] and len(c) == 10
    assert all(c[i + 1] <= 1.13 * c[i] + 16 for i in range(len(c) - 2))   # (the last chunk also takes the remainder)
    assert _cut(32768, 32768, -1) == (w, c)
    # uniform: what B200s have measured (8), and any explicit count
    assert _cut(32768, 32768, 8) == ([4096] * 8, [4096] * 8)
    assert _cut(32768, 32768, 16) == ([2048] * 16, [2048] * 16)
    # below 8192 the automatic choice stays 8 equal panels (the cut the validated GPU tests run)
    assert _cut(320, 320, 0) == ([128, 128, 64], [48] * 6 + [32])


def test_host_pipeline_model_prefers_the_default_cut():
    """tools/host_pipeline_model.py (three-stream timeline with the measured rates) on the library's own cuts: the default is
    ahead of the measured 8 equal panels over the whole range of PCIe rates and per-row DMA costs considered, 16 equal panels
    and doubling chunks are behind — a model, kept honest by bench.py measuring both passes"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("host_pipeline_model", os.path.join(ROOT, "tools", "host_pipeline_model.py"))
    hm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hm)
    n = 32768
    for rc in (0.5e-6, 2e-6):
        for bw in (45e9, 55e9):
            t = {p: hm.simulate(n, n, n, *hm.cut(n, n, p), rc, bw) for p in (0, 8, 16)}
            assert t[0] < t[8] < t[16]


@pytest.mark.parametrize("n,k", [(1, 1), (127, 5), (128, 16), (200, 200), (1100, 77), (2304, 2304), (8192, 8192), (12345, 9999),
                                 (65536, 512), (40000, 70000)])
@pytest.mark.parametrize("panels", [-1, 0, 1, 3, 8, 16, 64])
def test_host_pipeline_cut_covers_the_product_exactly(n, k, panels):
    """every column and every k exactly once, no empty piece, every panel but the last a multiple of the 128-wide CTA tile and
    every k-chunk but the last a multiple of the 16-deep k-tile (the GEMM's TMA boxes), any size"""
    w, c = _cut(n, k, panels)
    assert sum(w) == n and sum(c) == k and min(w) > 0 and min(c) > 0
    assert all(x % 128 == 0 for x in w[:-1]) and all(x % 16 == 0 for x in c[:-1])
    assert len(w) <= 70 and len(c) <= 70
    if panels < 0 and n >= 2048:   # graduated: the last panel is at most an eighth of the first (+ the ragged rest)
        assert w[-1] <= max(w) // 4 + 255 and c[0] <= k // 16 + 16


def _groups(nchunks, mode, first, last, host_ops, all_dma, fused, nn=True):
    import ctypes as C

    from candmc_b200._lib import lib

    out = (C.c_int * nchunks)()
    assert lib().candmc_debug_launch_groups(nchunks, mode, int(first), int(last), int(host_ops), int(all_dma), int(fused), int(nn), out) == 0
    hi, t, groups = list(out), 0, []
    while t < nchunks:
        assert t < hi[t] <= nchunks, (t, hi)
        groups.append((t, hi[t]))
        t = hi[t]
    return groups


def test_launch_groups_partition_the_chunks_and_keep_the_fused_chunk_apart():
    """summa_sweep's launch plan for every combination of its inputs: the groups partition [0, nchunks) in order,
    the chunk whose launch carries the fused depth sum is always alone, and the shapes the defaults ship are the documented ones"""
    import itertools

    for nchunks, mode, first, last, host_ops, all_dma, fused, nn in itertools.product(
            (1, 2, 3, 4, 8), (0, 1, 2, 3), (0, 1), (0, 1), (0, 1), (0, 1), (0, 1), (0, 1)):
        g = _groups(nchunks, mode, first, last, host_ops, all_dma, fused, nn)
        assert g[0][0] == 0 and g[-1][1] == nchunks and all(a[1] == b[0] for a, b in zip(g, g[1:]))
        if fused and last:
            assert g[-1] == (nchunks - 1, nchunks)
        if mode == 0 or not nn or nchunks <= 2:
            assert g == [(t, t + 1) for t in range(nchunks)]
    # the default (mode 2) on device operands by copy engines: chunk 0 + the rest on the first panel, one launch per later panel
    assert _groups(8, 2, True, False, False, True, False) == [(0, 1), (1, 8)]
    assert _groups(8, 2, False, True, False, True, False) == [(0, 8)]
    assert _groups(8, 2, True, True, False, True, True) == [(0, 1), (1, 7), (7, 8)]      # 2x2x2 with the fused depth sum
    # ... on the NCCL path the chunks' buffer slots come back launch by launch: chunk 0 + the rest for every panel
    assert _groups(8, 2, False, True, False, False, False) == [(0, 1), (1, 8)]
    # ... and while operands are still coming up from host memory: chunk 0, chunks 1-2, the rest
    assert _groups(8, 2, True, True, True, True, False) == [(0, 1), (1, 3), (3, 8)]
    assert _groups(4, 2, True, True, True, True, False) == [(0, 1), (1, 3), (3, 4)]
    assert _groups(8, 3, True, True, False, True, False) == [(0, 1), (1, 2), (2, 4), (4, 8)]   # doubling groups
    assert _groups(8, 1, False, True, False, True, False) == [(0, 7), (7, 8)]                  # last panel only


def test_early_c_slabs_cover_the_block():
    """the column slabs a host C block is finalised in: graduated b/2, b/4, b/8, b/8 where b allows, equal slabs otherwise"""
    import ctypes as C

    from candmc_b200._lib import lib

    for b, fin, want in ((16384, 8, [8192, 4096, 2048, 2048]), (1024, 8, [512, 256, 128, 128]), (512, 4, [128] * 4),
                         (768, 2, [384, 384]), (2048, 0, []), (2048, 1, [])):
        w, cnt = (C.c_int64 * 16)(), C.c_int()
        assert lib().candmc_debug_fin_slab_widths(b, fin, w, 16, C.byref(cnt)) == 0
        got = list(w)[:cnt.value]
        assert got == want and (not got or sum(got) == b), (b, fin, got)
