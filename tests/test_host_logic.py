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
