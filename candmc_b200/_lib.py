"""ctypes loader for libcandmc_b200.so (the C ABI declared in include/candmc_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcandmc_b200.so")
_lib = None

i64 = C.c_int64
pd = C.c_void_p  # device or host pointers travel as integers
comm_p = C.c_void_p


class CandmcError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference would assert/ABORT, alg/shared/util.h:127-138)."""

    def __init__(self, code, msg):
        super().__init__(f"candmc_b200 error {code}: {msg}")
        self.code = code


class CtbArgs(C.Structure):
    """candmc_ctb_args_t == ctb_args_t (alg/MM/topo_pdgemm/topo_pdgemm_algs.h:6-15)."""
    _fields_ = [("trans_A", C.c_char), ("trans_B", C.c_char), ("n", i64), ("lda_A", i64), ("lda_B", i64),
                ("lda_C", i64), ("buffer_size", i64), ("ovp", C.c_int)]


class PView(C.Structure):
    """candmc_pview_t == pview (alg/shared/comm.h:66-84)."""
    _fields_ = [("rrow", C.c_int), ("rcol", C.c_int), ("crow", C.c_void_p), ("ccol", C.c_void_p), ("cworld", C.c_void_p)]


class Aggregator(C.Structure):
    """candmc_aggregator_t: the reference's aggregator (alg/QR/qr_2d/qr_y2d.h:4-46) with device arrays."""
    _fields_ = [("lda_aQm", i64), ("lda_aT", i64), ("shift", i64), ("n", i64), ("aQm", C.c_void_p), ("aT", C.c_void_p),
                ("scratch", C.c_void_p)]


class DMat(C.Structure):
    """candmc_dmat_t: DMatrix (alg/SE/dmatrix.h:7-33) without the ScaLAPACK descriptor."""
    _fields_ = [("nrow", i64), ("ncol", i64), ("b", i64), ("lda", i64), ("data", C.c_void_p), ("pv", PView)]


# name -> (restype, argtypes); every symbol include/candmc_b200.h declares
SIGNATURES = {
    "candmc_version": (C.c_int, []),
    "candmc_last_error": (C.c_char_p, []),
    "candmc_init": (C.c_int, [C.c_int]),
    "candmc_finalize": (C.c_int, []),
    "candmc_device_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "candmc_launch_count": (C.c_ulonglong, []),
    "candmc_debug_force_generic_gemm": (C.c_int, [C.c_int]),
    "candmc_debug_static_schedule": (C.c_int, [C.c_int]),
    "candmc_debug_splitk": (C.c_int, [C.c_int]),
    "candmc_debug_prefetch_c": (C.c_int, [C.c_int]),
    "candmc_debug_transpose_tma": (C.c_int, [C.c_int]),
    "candmc_debug_gemm_tile": (C.c_int, [C.c_int]),
    "candmc_debug_gemm_reserve_sms": (C.c_int, [C.c_int]),
    "candmc_set_fused_reduce": (C.c_int, [C.c_int]),
    "candmc_set_skip_unused_uploads": (C.c_int, [C.c_int]),
    "candmc_set_check_peer_args": (C.c_int, [C.c_int]),
    "candmc_set_early_c_download": (C.c_int, [C.c_int]),
    "candmc_set_panel_transport": (C.c_int, [C.c_int]),
    "candmc_set_host_gather": (C.c_int, [C.c_int]),
    "candmc_debug_launch_groups": (C.c_int, [C.c_int] * 8 + [C.POINTER(C.c_int)]),
    "candmc_debug_fin_slab_widths": (C.c_int, [i64, C.c_int, C.POINTER(i64), C.c_int, C.POINTER(C.c_int)]),
    "candmc_panel_transport_sends": (C.c_ulonglong, []),
    "candmc_merged_panel_launches": (C.c_ulonglong, [C.c_int]),
    "candmc_set_b_first_chunk_early": (C.c_int, [C.c_int]),
    "candmc_profile_enable": (C.c_int, [C.c_int]),
    "candmc_profile_gemm_timeline": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), i64, C.POINTER(i64)]),
    "candmc_set_background_ctas": (C.c_int, [C.c_int]),
    "candmc_profile_gemm_stats": (C.c_int, [C.POINTER(i64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "candmc_dgemm_chunked_b": (C.c_int, [C.c_char, i64, i64, i64, i64, C.c_double, pd, i64, pd, C.c_double, pd, i64, C.c_void_p]),
    "candmc_dgemm": (C.c_int, [C.c_char, C.c_char, i64, i64, i64, C.c_double, pd, i64, pd, i64, C.c_double, pd, i64,
                               C.c_void_p]),
    "candmc_set_trsm_variant": (C.c_int, [C.c_int]),
    "candmc_host_pipeline_cut": (C.c_int, [i64, i64, C.c_int, C.POINTER(i64), C.POINTER(i64), C.c_int, C.POINTER(C.c_int),
                                           C.POINTER(C.c_int)]),
    "candmc_sgemm": (C.c_int, [C.c_char, C.c_char, i64, i64, i64, C.c_float, C.c_void_p, i64, C.c_void_p, i64, C.c_float,
                               C.c_void_p, i64, C.c_void_p]),
    "candmc_set_f32_mode": (C.c_int, [C.c_int]),
    "candmc_lda_cpy": (C.c_int, [i64, i64, i64, i64, pd, pd, C.c_void_p]),
    "candmc_lda_cpy_scaled": (C.c_int, [i64, i64, i64, i64, pd, pd, C.c_double, C.c_double, C.c_void_p]),
    "candmc_transpose": (C.c_int, [i64, i64, pd, i64, pd, i64, C.c_void_p]),
    "candmc_fill_drand48": (C.c_int, [pd, i64, i64, i64, i64, i64, i64, C.c_int, C.c_void_p]),
    "candmc_frob_diff": (C.c_int, [pd, i64, pd, i64, i64, i64, C.POINTER(C.c_double), C.c_void_p]),
    "candmc_get_unique_id": (C.c_int, [C.c_void_p]),
    "candmc_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(comm_p)]),
    "candmc_comm_split": (C.c_int, [comm_p, C.c_int, C.c_int, C.POINTER(comm_p)]),
    "candmc_comm_free": (C.c_int, [comm_p]),
    "candmc_comm_rank": (C.c_int, [comm_p, C.POINTER(C.c_int)]),
    "candmc_comm_size": (C.c_int, [comm_p, C.POINTER(C.c_int)]),
    "candmc_comm_barrier": (C.c_int, [comm_p]),
    "candmc_comm_bcast": (C.c_int, [comm_p, pd, i64, C.c_int, C.c_void_p]),
    "candmc_comm_allreduce_sum": (C.c_int, [comm_p, pd, pd, i64, C.c_void_p]),
    "candmc_summa": (C.c_int, [C.POINTER(CtbArgs), pd, pd, pd, pd, comm_p, comm_p, C.c_void_p]),
    "candmc_d25_summa": (C.c_int, [C.POINTER(CtbArgs), pd, pd, pd, pd, comm_p, comm_p, comm_p, C.c_int, C.c_void_p]),
    "candmc_bcast_cannon_4d": (C.c_int, [C.POINTER(CtbArgs), pd, pd, pd, pd, comm_p, comm_p, comm_p, comm_p,
                                         C.c_void_p]),
    "candmc_spcannon": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, comm_p, C.c_int, C.c_int, C.c_int, C.c_char,
                                  C.c_double, pd, C.c_char, C.c_double, pd, pd, C.c_void_p]),
    "candmc_upd_A": (C.c_int, [pd, i64, pd, i64, i64, i64, i64, pd, comm_p, C.c_void_p]),
    "candmc_update_A": (C.c_int, [pd, i64, pd, i64, i64, i64, i64, pd, C.POINTER(PView), pd, i64, C.c_int, C.c_void_p]),
    "candmc_upd_Yamamoto_A": (C.c_int, [pd, i64, pd, i64, i64, i64, i64, pd, comm_p, C.c_void_p]),
    "candmc_update_Yamamoto_A": (C.c_int, [pd, i64, pd, i64, i64, i64, i64, pd, C.POINTER(PView), C.c_void_p]),
    "candmc_update_Yamamoto_A_agg": (C.c_int, [pd, i64, pd, i64, i64, i64, i64, pd, C.POINTER(PView), C.POINTER(Aggregator), C.c_int,
                                               C.c_void_p]),
    "candmc_aggregator_create": (C.c_int, [i64, i64, C.POINTER(Aggregator)]),
    "candmc_aggregator_reset": (C.c_int, [C.POINTER(Aggregator)]),
    "candmc_aggregator_shift_down": (C.c_int, [C.POINTER(Aggregator), i64]),
    "candmc_aggregator_free": (C.c_int, [C.POINTER(Aggregator)]),
    "candmc_sym_full2band_update": (C.c_int, [pd, i64, i64, i64, i64, C.POINTER(PView), comm_p, pd, i64, C.c_void_p]),
    "candmc_sym_full2band_extents": (C.c_int, [i64, i64, i64] + [C.c_int] * 5 + [C.POINTER(i64)] * 4),
    "candmc_set_min_kchunk": (C.c_int, [i64]),
    "candmc_set_merge_last_panel": (C.c_int, [C.c_int]),
    "candmc_set_merge_panels": (C.c_int, [C.c_int]),
    "candmc_set_host_pipeline_min": (C.c_int, [i64]),
    "candmc_set_host_pipeline_panels": (C.c_int, [C.c_int]),
    "candmc_redistribute": (C.c_int, [C.c_int, i64, i64, i64, pd, i64, pd, i64, C.POINTER(PView), C.c_void_p]),
    "candmc_redist_axis_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, i64, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "candmc_redist_strided_index": (C.c_int, [C.c_int, C.c_int, C.c_int, i64, C.c_int, C.c_int, i64, C.c_int, i64, i64,
                                              C.POINTER(i64), C.POINTER(C.c_int)]),
    "candmc_dmat_local_extents": (C.c_int, [C.POINTER(DMat), C.POINTER(i64), C.POINTER(i64)]),
    "candmc_dmat_slice": (C.c_int, [C.POINTER(DMat), i64, i64, i64, i64, C.POINTER(DMat)]),
    "candmc_dmat_get_contig": (C.c_int, [C.POINTER(DMat), pd, C.c_void_p]),
    "candmc_dmat_replicate_vertical": (C.c_int, [C.POINTER(DMat), pd, C.c_void_p]),
    "candmc_dmat_replicate_horizontal": (C.c_int, [C.POINTER(DMat), pd, C.c_void_p]),
    "candmc_dmat_reduce_scatter_horizontal": (C.c_int, [C.POINTER(DMat), pd, C.c_void_p]),
    "candmc_dmat_transpose_data": (C.c_int, [C.POINTER(DMat), pd, C.c_void_p]),
    "candmc_dmat_foldcols": (C.c_int, [C.POINTER(DMat), i64, pd, C.c_void_p]),
    "candmc_dmat_foldrows": (C.c_int, [C.POINTER(DMat), i64, pd, C.c_void_p]),
    "candmc_debug_fold_src_index": (C.c_int, [C.c_int] + [i64] * 7 + [C.POINTER(i64)]),
    "candmc_debug_redist_permute": (C.c_int, [C.c_int, C.c_int, C.c_int, i64, C.c_int, C.c_int, C.c_int, pd, i64, pd, i64,
                                              i64, C.c_void_p]),
    # accelerator seam of the 2.5D LU (alg/LU/lu_offload.h)
    "candmc_off_set_device": (C.c_int, [C.c_int]),
    "candmc_off_alloc": (C.c_int, [C.c_int, i64, pd]),
    "candmc_off_free": (C.c_int, [C.c_int]),
    "candmc_off_alloc_transfer": (C.c_int, [i64]),
    "candmc_off_free_transfer": (C.c_int, []),
    "candmc_off_device_ptr": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(i64)]),
    "candmc_off_size": (C.c_int, [C.c_int, C.POINTER(i64)]),
    "candmc_off_host_mirror": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "candmc_off_gemm": (C.c_int, [C.c_char, C.c_char, i64, i64, i64, C.c_double, i64, C.c_int, i64, i64, C.c_int, i64,
                                  C.c_double, i64, C.c_int, i64]),
    "candmc_off_wait_gemm": (C.c_int, []),
    "candmc_off_upload": (C.c_int, [i64, i64, i64, i64, pd, i64, C.c_int]),
    "candmc_off_download": (C.c_int, [i64, i64, i64, i64, i64, pd, C.c_int]),
    "candmc_off_sparse_rw": (C.c_int, [i64, i64, i64, pd, i64, C.POINTER(C.c_int), C.c_int, C.c_char]),
    "candmc_off_sync": (C.c_int, []),
    "candmc_off_set_overlap": (C.c_int, [C.c_int]),
    "candmc_off_stats": (C.c_int, [C.POINTER(i64)]),
    "candmc_off_blocks_overlap": (C.c_int, [i64] * 8),
}


def build_native(verbose: bool = False) -> str:
    """Compile libcandmc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return _SO


def lib() -> C.CDLL:
    """The loaded shared library.  Fails loudly when it is missing — there is no pure-Python or CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise CandmcError(-1, f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(or `make -C candmc_b200/csrc`). candmc_b200 has no fallback implementation.")
        L = C.CDLL(_SO, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().candmc_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc != 0:
        raise CandmcError(rc, last_error())


def init(device: int = -1):
    check(lib().candmc_init(device))


def launch_count() -> int:
    return int(lib().candmc_launch_count())
