"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/candmc_b200.h declares,
fails loudly without a GPU, and the product never touches oracle/ or tests/cpusim."""
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
    with pytest.raises(cb.CandmcError) as e:
        cb.csgemm("T", "N", 2, 2, 2, 1.0, 0, 2, 0, 2, 0.0, 0, 2)
    assert e.value.code == 5
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
                    if re.search(r"oracle[/_.]|liboracle|mpi_shim|cpusim|CPUSIM", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference oracle/ or the test-only simulator: {bad}"


def test_product_library_does_not_contain_the_simulator():
    """tests/cpusim builds its own library out of the product sources; the shipped library must not know about it"""
    import subprocess

    so = os.path.join(ROOT, "candmc_b200", "libcandmc_b200.so")
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True, check=True).stdout
    assert "cpusim" not in out
    assert " T cudaMalloc" not in out   # the real (static) CUDA runtime is linked, nothing re-exports a fake one


def test_lu_offload_library_exports_the_reference_entry_points():
    """libcandmc_lu_offload.so must define, with C++ linkage, exactly what alg/LU/lu_offload.h:19-101 declares, so that the
    reference's LU objects link against it instead of alg/LU/lu_offload.cxx."""
    import subprocess

    so = os.path.join(ROOT, "candmc_b200", "libcandmc_lu_offload.so")
    assert os.path.exists(so), "run __graft_entry__.build()"
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", so], capture_output=True, text=True, check=True).stdout
    want = ["get_mat_handle(OFF_MAT)", "set_mic_rank(int)", "wait_gemm()",
            "offload_gemm_A(char, char, int, int, int, double, int, OFF_MAT, int, int, OFF_MAT, int, double, int, OFF_MAT, int)",
            "download_lda_cpy(int, int, int, int, int, double*, OFF_MAT)",
            "upload_lda_cpy(int, int, int, int, double const*, int, OFF_MAT)",
            "offload_sparse_rw(int, int, int, double*, int, int*, OFF_MAT, char)", "alloc_A(long, double*)", "alloc_L(long)",
            "alloc_U(long)", "alloc_transfer(long)", "free_offload_A()", "free_offload_L()", "free_offload_U()",
            "free_offload_transfer()"]
    for w in want:
        assert w in out, f"{w} not exported by libcandmc_lu_offload.so"
    # and the header that mirrors the reference interface declares the same names
    hdr = open(os.path.join(ROOT, "include", "candmc", "lu_offload.h")).read()
    for w in want:
        assert w.split("(")[0] in hdr


def test_lu_offload_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU box")
    import candmc_b200 as cb
    from candmc_b200 import lu_offload as lo

    with pytest.raises(cb.CandmcError) as e:
        lo.alloc_L(16)
    assert e.value.code == 5


def test_cxx_headers_compile_and_cover_the_widening_entry_points(tmp_path):
    """include/CANDMC.h is what a reference-style C++ caller includes; the inline wrappers of the widening rows must compile
    against the C ABI they forward to (g++ only, nothing is run)"""
    import subprocess

    src = tmp_path / "t.cxx"
    src.write_text("""
#include "CANDMC.h"
void use(pview* pv, double* A, double* Y, double* B) {
  sym_full2band_update(A, 64, 128, 32, 8, pv, Y, 48);
  cyclic_to_blocked(64, 64, 8, A, 32, B, 32, pv);
  blocked_to_cyclic(64, 64, 8, B, 32, A, 32, pv);
  // CAQR trailing updates under the reference's names and argument lists (alg/QR/qr_2d/qr_2d.h:74-108, qr_y2d.h:59-93)
  update_A(Y, 64, A, 64, 128, 96, 16, B, pv, NULL, 0);              // W = the panel QR's factor (QR_2D's own call, qr_2d.cxx:325)
  update_A(Y, 64, A, 64, 128, 96, 16, NULL, pv, B, 64);             // T from Y, panel saved into aggreg_Y
  update_A(Y, 64, A, 64, 128, 96, 16, B, pv, NULL, 0, true);        // W is T
  upd_A(Y, 64, A, 64, 64, 48, 16, B, pv, true);
  upd_A(Y, 64, A, 64, 64, 48, 16, NULL, pv);                        // W == NULL: T from Y on the device (QR_2D_2D's call, qr_2d.cxx:873)
  update_Yamamoto_A(Y, 64, A, 64, 128, 96, 16, B, pv, NULL);
  upd_Yamamoto_A(Y, 64, A, 64, 64, 48, 16, B, pv);
  // ... and with the reference's aggregator (qr_y2d.h:4-46), as QR_Yamamoto_2D_2D uses it (qr_y2d.cxx:333-377)
  aggregator agg(64, 32);
  update_Yamamoto_A(Y, 64, A, 64, 128, 32, 16, B, pv, &agg);
  agg.shift_down(16);
  append_last_Yamamoto_panel(Y, 64, 112, 16, B, pv, &agg);
  upd_Yamamoto_A(agg.aQm, agg.lda_aQm, A, 64, 64, 48, agg.n, agg.aT, pv);
  agg.reset();
}
""")
    inc = os.path.join(ROOT, "include")
    p = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-I", inc, "-I", os.path.join(inc, "candmc_compat"), str(src)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]


def test_integration_seams_compile(tmp_path):
    """integration/*.cxx are what a maintainer compiles inside the REFERENCE's build.  cdgemm_gpu.cxx needs nothing but the C ABI
    header; qr_2d_upd_A_gpu.cxx needs the reference's own headers, so it is compiled where /root/reference exists (with the
    flags of oracle/Makefile, against the mini-MPI shim's mpi.h).  g++ only, nothing is run."""
    import subprocess

    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "cdgemm_gpu.cxx"), "-o", str(tmp_path / "cdgemm_gpu.o")])
    syms = subprocess.check_output(["nm", "-C", "--defined-only", str(tmp_path / "cdgemm_gpu.o")], text=True)
    assert " T cdgemm(char, char, int, int, int, double, double const*, int, double const*, int, double, double*, int)" in syms
    ref = "/root/reference"
    if not os.path.isdir(ref):
        return
    subprocess.check_call(["g++", "-c", "-D_POSIX_C_SOURCE=200112L", "-D__STDC_LIMIT_MACROS", "-Drestrict=__restrict__", "-w",
                           "-I", os.path.join(ROOT, "oracle", "mpi_shim"), "-I", os.path.join(ref, "include"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "qr_2d_upd_A_gpu.cxx"), "-o", str(tmp_path / "qr_seam.o")])
    syms = subprocess.check_output(["nm", "-C", "--defined-only", str(tmp_path / "qr_seam.o")], text=True)
    assert " T upd_A(" in syms and " T upd_Yamamoto_A(" in syms
    # neither seam touches test infrastructure
    for f in ("cdgemm_gpu.cxx", "qr_2d_upd_A_gpu.cxx"):
        assert "oracle" not in open(os.path.join(ROOT, "integration", f)).read().replace("oracle/qr_2d_tap", "")
