"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/candmc_b200.h declares,
fails loudly without a GPU, and the product never touches oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "candmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(candmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import candmc_b200 as cb
    from candmc_b200 import _lib

    L = cb.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/candmc_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in candmc_b200/_lib.py"
    assert L.candmc_version() == 100


def test_ctb_args_layout_matches_reference_struct():
    """ctb_args_t (alg/MM/topo_pdgemm/topo_pdgemm_algs.h:6-15): two chars, five int64, one int -> 56 bytes on LP64."""
    from candmc_b200._lib import CtbArgs

    assert ctypes.sizeof(CtbArgs) == 56
    assert CtbArgs.n.offset == 8 and CtbArgs.buffer_size.offset == 40 and CtbArgs.ovp.offset == 48


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU box")
    import candmc_b200 as cb

    with pytest.raises(cb.CandmcError) as e:
        cb.cdgemm("N", "N", 2, 2, 2, 1.0, 0, 2, 0, 2, 0.0, 0, 2)
    assert e.value.code == 5  # CANDMC_ERR_NODEVICE
    with pytest.raises(cb.CandmcError):
        cb.init_world(0, 1, 0)


def test_product_never_references_the_oracle():
    bad = []
    for base in ("candmc_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cxx", ".cpp", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"oracle[/_.]|liboracle|mpi_shim", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference oracle/: {bad}"
