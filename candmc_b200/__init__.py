"""candmc_b200 — B200-native CANMM hot path (distributed FP64 matrix multiply) behind the reference's interface.

Python side of the drop-in boundary: a ctypes binding of `libcandmc_b200.so` (C ABI in include/candmc_b200.h) plus a
host-side mirror of the reference's operator interface (`ctb_args_t`, `CommData_t`, `summa`, `d25_summa`,
`d25_summa_ovp`, `bcast_cannon_4d`, `kput_cannon`, `kuni_cannon`, `cdgemm`, `lda_cpy`, the grid-setup macros) so tests
and benchmarks read like the reference's own drivers.  PyTorch is used for device memory, streams and the
torch.distributed bootstrap only.  There is no CPU fallback: every compute call raises CandmcError without a B200.
"""
from ._lib import CandmcError, lib, build_native, launch_count, init, last_error  # noqa: F401
from .mm import (  # noqa: F401
    CommData_t,
    ctb_args_t,
    cdgemm,
    csgemm,
    lda_cpy,
    transpose,
    summa,
    d25_summa,
    d25_summa_ovp,
    bcast_cannon_4d,
    kput_cannon,
    kuni_cannon,
    upd_A,
    update_A,
    pview,
    fill_drand48,
    frob_diff,
    set_min_kchunk,
    cdgemm_chunked_b,
    upd_Yamamoto_A,
    update_Yamamoto_A,
    aggregator,
    cyclic_to_blocked,
    blocked_to_cyclic,
    sym_full2band_update,
    sym_full2band_extents,
)
from .grid import (  # noqa: F401
    init_world,
    rsetup_kdir_comm,
    rsetup_layer_comm,
    setup_sub_comm,
    d25_grid,
    dcn_grid,
    grid_shape_for,
)

from .dmatrix import DMatrix  # noqa: F401,E402

__version__ = "0.1.0"
