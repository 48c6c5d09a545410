"""The tcgen05 descriptors of the FP32 kernel against CUTLASS/CuTe's own.

candmc_b200/csrc/gemm_f32.cu hand-encodes the UMMA shared-memory descriptor of its K-major, 128-byte-swizzled operand tiles
and the kind::tf32 instruction descriptor (common.cuh).  The CPU simulator decodes them by the bit-field definitions, but
whether a hand-encoded descriptor means what the hardware takes it to mean can only be checked against an encoder that has run
on the hardware: this test compiles tests/native/cutlass_desc_check.cu (host-only) against the CuTe headers vendored in the
image and compares bit for bit.  Skipped where nvcc or the headers are missing."""
import glob
import json
import os
import shutil
import subprocess
import sysconfig

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _cutlass_include():
    site = sysconfig.get_paths()["purelib"]
    for pat in ("flashinfer/data/cutlass/include", "tilelang/3rdparty/cutlass/include", "vllm/third_party/deep_gemm/include"):
        for d in glob.glob(os.path.join(site, pat)):
            if os.path.exists(os.path.join(d, "cute", "arch", "mma_sm100_desc.hpp")) and os.path.exists(os.path.join(d, "cutlass")):
                return d
    return None


def test_umma_descriptors_match_cute(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = _cutlass_include()
    if not os.path.exists(nvcc) or inc is None:
        pytest.skip("needs nvcc and the vendored CuTe headers")
    exe = str(tmp_path / "cutlass_desc_check")
    p = subprocess.run([nvcc, "-std=c++17", "-O0", "-w", "-I", inc, "-I", os.path.join(ROOT, "candmc_b200", "csrc"),
                        "--expt-relaxed-constexpr", os.path.join(HERE, "native", "cutlass_desc_check.cu"), "-o", exe],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
    lines = [line for line in out.splitlines() if line.startswith("{")]
    assert json.loads(lines[1])["swizzle_mismatches"] == 0     # CuTe's Swizzle<3,4,3> == the XOR of the simulator's TMA / UMMA emulation
    r = json.loads(lines[0])
    assert r["candmc_smem_desc"] == r["cutlass_smem_desc"]     # SBO 1024 B, LBO, version 1, SWIZZLE_128B (start address 0 on the host)
    assert r["candmc_idesc"] == r["cutlass_idesc"]             # kind::tf32, F32 accumulate, K-major x K-major, M = N = 128
    # the address arithmetic of the kernel: rows 128 B apart, 8-row groups 1024 B apart (= SBO), K = 8 steps 32 B apart
    assert (r["elem_offset_row1"], r["elem_offset_row8"], r["elem_offset_k8"], r["elem_offset_k24"]) == (32, 256, 8, 24)
