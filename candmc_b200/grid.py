"""Processor-grid construction: the reference's communicator macros (alg/shared/comm.h:117-201) on NCCL.

`init_world` plays INIT_COMM: one process per GPU (launched by torchrun or any launcher that sets RANK / WORLD_SIZE /
LOCAL_RANK); rank 0 creates the NCCL unique id through the C ABI and ships it with torch.distributed (plumbing only).
The split helpers use the same colour/key arithmetic as the macros, so world rank r = layer*q*q + row*q + col.
"""
from __future__ import annotations

import ctypes as C
import math
import os

from ._lib import check, lib, CandmcError
from .mm import CommData_t


def kdir_color_key(my_rank: int, p: int, c: int):
    """RSETUP_KDIR_COMM (comm.h:171-180): commrank (= layer) = r / (p/c), colour (= rank inside layer) = r % (p/c)."""
    return my_rank % (p // c), my_rank // (p // c)  # (color, key)


def layer_coords(intra_layer_rank: int, pesdim: int):
    """RSETUP_LAYER_COMM (comm.h:183-195): row = colour / pesdim, col = colour % pesdim."""
    return intra_layer_rank // pesdim, intra_layer_rank % pesdim


def grid_shape_for(num_pes: int, c_rep: int | None = None):
    """(q, c) the reference drivers pick for `num_pes` ranks (bench/MM/topo_pdgemm_bench.cxx:448-456): c = 1 for square
    counts, 2 for 8/32/...; plus this implementation's 1 x 1 x 2 k-split for two ranks."""
    if c_rep is None:
        c_rep = 1
        if math.isqrt(num_pes) ** 2 != num_pes:
            if num_pes >= 8 and num_pes % 2 == 0:
                c_rep = 2
            elif num_pes == 2:
                c_rep = 2
    q = math.isqrt(num_pes // c_rep)
    if q * q * c_rep != num_pes:
        raise CandmcError(1, f"processor grid mismatch: {num_pes} ranks cannot form q x q x {c_rep}")
    return q, c_rep


def _wrap(handle, size=None, rank=None):
    r, s = C.c_int(), C.c_int()
    check(lib().candmc_comm_rank(handle, C.byref(r)))
    check(lib().candmc_comm_size(handle, C.byref(s)))
    return CommData_t(cm=handle.value if hasattr(handle, "value") else handle, np=s.value, rank=r.value)


def init_world(rank: int | None = None, world_size: int | None = None, device: int | None = None,
               exchange=None) -> CommData_t:
    """INIT_COMM (comm.h:125-136).  `exchange(bytes_or_None) -> bytes` broadcasts rank 0's 128-byte id; by default it
    uses torch.distributed (which must already be initialised when world_size > 1)."""
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    check(lib().candmc_init(device))
    uid = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib().candmc_get_unique_id(uid))
    if world_size > 1:
        if exchange is None:
            exchange = _torch_exchange
        data = exchange(bytes(uid) if rank == 0 else None)
        uid = (C.c_ubyte * 128).from_buffer_copy(data)
    handle = C.c_void_p()
    check(lib().candmc_comm_init_rank(uid, world_size, rank, C.byref(handle)))
    return CommData_t(cm=handle.value, np=world_size, rank=rank)


def _torch_exchange(payload):
    import torch.distributed as dist

    box = [payload]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def setup_sub_comm(master: CommData_t, commrank: int, bcolor: int, p: int) -> CommData_t:
    """SETUP_SUB_COMM (comm.h:143-152): MPI_Comm_split(master, colour, key=commrank)."""
    handle = C.c_void_p()
    check(lib().candmc_comm_split(master.cm, bcolor, commrank, C.byref(handle)))
    out = _wrap(handle)
    out.color = bcolor
    if out.np != p or out.rank != commrank:
        raise CandmcError(1, f"sub-communicator came out as rank {out.rank}/{out.np}, expected {commrank}/{p}")
    return out


def rsetup_kdir_comm(world: CommData_t, c: int):
    """RSETUP_KDIR_COMM (comm.h:171-180) -> (cdt_kdir, layerRank, intraLayerRank)."""
    color, key = kdir_color_key(world.rank, world.np, c)
    cdt = setup_sub_comm(world, key, color, c)
    return cdt, key, color


def rsetup_layer_comm(world: CommData_t, pesdim: int, layer_rank: int, intra_layer_rank: int):
    """RSETUP_LAYER_COMM (comm.h:183-195) -> (cdt_row, cdt_col, myRow, myCol); the intra-layer communicator is
    created and released inside, as in the macro."""
    handle = C.c_void_p()
    check(lib().candmc_comm_split(world.cm, layer_rank, intra_layer_rank, C.byref(handle)))
    layer = _wrap(handle)
    my_row, my_col = layer_coords(intra_layer_rank, pesdim)
    cdt_row = setup_sub_comm(layer, my_col, my_row, pesdim)   # split(myRow, key myCol): ranks of my grid row
    cdt_col = setup_sub_comm(layer, my_row, my_col, pesdim)   # split(myCol, key myRow): ranks of my grid column
    layer.free()
    return cdt_row, cdt_col, my_row, my_col


def d25_grid(world: CommData_t, c_rep: int | None = None):
    """The grid d25_unit / d25_bench build (test/MM/topo_pdgemm_unit.cxx:197-230): returns a dict with q, c, the three
    communicators and (layer, myRow, myCol)."""
    q, c = grid_shape_for(world.np, c_rep)
    cdt_kdir, layer, intra = rsetup_kdir_comm(world, c)
    cdt_row, cdt_col, my_row, my_col = rsetup_layer_comm(world, q, layer, intra)
    return dict(q=q, c=c, cdt_row=cdt_row, cdt_col=cdt_col, cdt_kdir=cdt_kdir, layer=layer, row=my_row, col=my_col)


def dcn_grid(world: CommData_t, x2_np: int):
    """The 4-D grid of dcn_unit (test/MM/topo_pdgemm_unit.cxx:30-72): r = x1 + x1_np*(y1 + x1_np*(x2 + x2_np*y2))."""
    p, r = world.np, world.rank
    x1_np = math.isqrt(p // (x2_np * x2_np))
    if x1_np * x1_np * x2_np * x2_np != p:
        raise CandmcError(1, "processor grid mismatch for the 4-D Cannon/SUMMA grid")
    x1, y1 = r % x1_np, (r // x1_np) % x1_np
    x2, y2 = (r // (x1_np * x1_np)) % x2_np, r // (x1_np * x1_np * x2_np)
    cdt_y2 = setup_sub_comm(world, y2, r % (x1_np * x1_np * x2_np), x2_np)
    cdt_x2 = setup_sub_comm(world, x2, (r % (x1_np * x1_np)) * x2_np + y2, x2_np)
    cdt_y1 = setup_sub_comm(world, y1, (r // (x1_np * x1_np)) * x1_np + (r % x1_np), x1_np)
    cdt_x1 = setup_sub_comm(world, x1, r // x1_np, x1_np)
    return dict(x1_np=x1_np, x2_np=x2_np, x1=x1, y1=y1, x2=x2, y2=y2, cdt_x1=cdt_x1, cdt_y1=cdt_y1, cdt_x2=cdt_x2,
                cdt_y2=cdt_y2)
