"""The product's hot kernel (candmc_b200/csrc/gemm_f64.cu: TMA + mbarrier ring + swizzled fragment loads + DMMA.8x8x4, split-K,
dynamic tile scheduler) executed on the simulator's PTX emulation (CPUSIM_GEMM=device) against numpy.  TEST INFRASTRUCTURE.
Prints one JSON line."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
os.environ["CPUSIM_GEMM"] = "device"
import simtorch; simtorch.install()  # noqa: E402,E702
import numpy as np  # noqa: E402
import torch  # noqa: E402
import candmc_b200 as cb  # noqa: E402
from candmc_b200._lib import lib, check  # noqa: E402

check(lib().candmc_init(0))
EPS = 2.220446049250313e-16
worst, cases = 0.0, 0


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def run(ta, tb, m, n, k, alpha, beta, pad=0, c_nan=False, offset=0, seed=0):
    """one cdgemm against numpy; returns the device result (m x n)"""
    global worst, cases
    rng = np.random.RandomState(seed + m * 7 + n * 3 + k)
    ra, ca = (m, k) if ta == "N" else (k, m)
    rb, cbn = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = max(ra, 1) + pad, max(rb, 1) + pad, max(m, 1) + pad
    A = np.zeros((lda, max(ca, 1)), order="F"); A[:ra, :ca] = rng.rand(ra, ca) - 0.5
    B = np.zeros((ldb, max(cbn, 1)), order="F"); B[:rb, :cbn] = rng.rand(rb, cbn) - 0.5
    C = np.zeros((ldc, n), order="F"); C[:m] = rng.rand(m, n) - 0.5
    if c_nan:
        C[:] = np.nan
    flatA = np.concatenate([np.zeros(offset), A.reshape(-1, order="F")])   # offset = 1: an 8-byte-aligned base (generic kernel)
    dA = torch.from_numpy(flatA).cuda()
    dB, dC = dev(B), dev(C)
    cb.cdgemm(ta, tb, m, n, k, alpha, dA.data_ptr() + 8 * offset, lda, dB, ldb, beta, dC, ldc)
    torch.cuda.synchronize()
    got = dC.cpu().numpy().reshape(n, ldc).T[:m].copy()
    opA = A[:ra, :ca] if ta == "N" else A[:ra, :ca].T
    opB = B[:rb, :cbn] if tb == "N" else B[:rb, :cbn].T
    ref = alpha * (opA @ opB) + (beta * C[:m] if beta != 0.0 else 0.0)
    err = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
    assert err <= 10 * max(k, 1) * EPS, (ta, tb, m, n, k, alpha, beta, pad, err)
    if pad:   # rows below m are never written
        tail = dC.cpu().numpy().reshape(n, ldc).T[m:]
        assert np.array_equal(tail, C[m:]) or (c_nan and np.isnan(tail).all())
    worst, cases = max(worst, err), cases + 1
    return got


# every transpose combination, whole tiles and ragged edges in m, n and k (TMA zero-fills out-of-range rows / columns / k)
for ta in "NT":
    for tb in "NT":
        run(ta, tb, 128, 128, 32, 1.0, 0.0)
        run(ta, tb, 150, 70, 36, -0.5, 1.0)
        run(ta, tb, 9, 300, 18, 2.0, -1.0, pad=2)
run("N", "N", 1, 1, 1, 1.0, 0.0, pad=1)
run("N", "N", 256, 128, 16, 1.0, 0.0)                 # several tiles, one k-tile
run("N", "N", 128, 128, 16 * 7 + 3, 1.0, 1.0)         # more k-tiles than ring stages, ragged last one
run("N", "N", 130, 130, 20, 1.0, 0.0, c_nan=True)     # beta == 0 never reads C
run("T", "N", 64, 64, 16 * 40, 1.0, 1.0)              # few tiles, long k: the split-K path
run("N", "N", 96, 40, 0, 1.0, 0.5)                    # k == 0: C = beta C
run("N", "N", 96, 40, 24, 0.0, 0.0, c_nan=True)       # alpha == 0, beta == 0: zeros without reading anything
# unaligned operands leave the TMA path for the CUDA-core kernel: odd leading dimension, 8-byte-aligned base
run("N", "N", 65, 33, 17, 1.0, 1.0, pad=0)
run("N", "T", 64, 64, 16, 1.0, 0.0, offset=1)
# split-K sums its partial tiles in a fixed order: bit-identical run to run; the static schedule computes the same bits
a = run("T", "N", 128, 128, 16 * 48, 1.0, 0.0, seed=5)
b = run("T", "N", 128, 128, 16 * 48, 1.0, 0.0, seed=5)
assert np.array_equal(a, b)
check(lib().candmc_debug_static_schedule(1))
c = run("T", "N", 128, 128, 16 * 48, 1.0, 0.0, seed=5)
check(lib().candmc_debug_static_schedule(0))
assert np.array_equal(a, c)
check(lib().candmc_debug_splitk(0))
d = run("T", "N", 128, 128, 16 * 48, 1.0, 0.0, seed=5)
check(lib().candmc_debug_splitk(1))
assert np.abs(a - d).max() <= 10 * 768 * EPS
# the second CTA shape (128 x 64 tiles, two CTAs per SM, 4-stage ring, 4 consumer warps): every transpose combination with ragged
# edges, more k-tiles than stages, an L2 prefetch of the C tile (beta != 0), and the automatic choice between the shapes
check(lib().candmc_debug_gemm_tile(64))
for ta in "NT":
    for tb in "NT":
        run(ta, tb, 150, 70, 36, -0.5, 1.0)
        run(ta, tb, 128, 200, 16 * 6 + 3, 2.0, 0.0, pad=2, c_nan=True)
run("N", "N", 300, 64, 16, 1.0, 1.0)
e = run("N", "N", 256, 192, 80, 1.0, 0.5, seed=9)
check(lib().candmc_debug_gemm_tile(128))
f = run("N", "N", 256, 192, 80, 1.0, 0.5, seed=9)
check(lib().candmc_debug_gemm_tile(0))
g = run("N", "N", 256, 192, 80, 1.0, 0.5, seed=9)
assert np.array_equal(e, f) and np.array_equal(e, g)   # same sums in the same order, whichever CTA shape multiplies
print(json.dumps({"cases": cases, "max_rel_err": worst, "launches": int(cb.launch_count())}))
