"""The FP32 kernel (candmc_b200/csrc/gemm_f32.cu: TMA -> operand split in shared memory -> tcgen05.mma kind::tf32 into tensor
memory -> tcgen05.ld epilogue) through the C ABI (candmc_sgemm) against a float64 numpy product of the same float32 data.

    python tests/f32_worker.py                    on a B200 (tests/test_zz_f32_gpu.py)
    CANDMC_CPUSIM=1 python tests/f32_worker.py    on the CPU simulator's tcgen05 emulation (tests/test_cpusim.py)

The emulation decodes the UMMA descriptors by CUTLASS's bit-field definitions, truncates operands to TF32 as the tensor core
does and accumulates in FP32, so on the simulator the accuracy figures are those of the 3xTF32 scheme itself.  Error measure:
max over elements of |C - C_ref| / (|alpha| |op(A)| |op(B)| + |beta| |C|), i.e. relative to the size of the terms summed.
Prints one JSON line."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
SIM = os.environ.get("CANDMC_CPUSIM") == "1"
if SIM:
    sys.path.insert(0, os.path.join(HERE, "cpusim"))
    import simtorch
    simtorch.install()
import numpy as np  # noqa: E402
import torch  # noqa: E402
import candmc_b200 as cb  # noqa: E402
from candmc_b200._lib import lib, check  # noqa: E402

check(lib().candmc_init(0))
if "--one" in sys.argv:   # a single large launch for a profiler (tools/r02_session.sh f32): no checks, no output worth reading
    nb = int(sys.argv[sys.argv.index("--one") + 1])
    x = torch.rand(nb * nb, dtype=torch.float32, device="cuda") - 0.5
    z = torch.empty(nb * nb, dtype=torch.float32, device="cuda")
    cb.csgemm("T", "N", nb, nb, nb, 1.0, x, nb, x, nb, 0.0, z, nb)
    torch.cuda.synchronize()
    print(json.dumps({"one": nb}))
    sys.exit(0)
TOL3 = 4 * 2.0 ** -20      # documented bound of the split scheme (3 * 2^-20 per product, rounded up)
TOL1 = 2.0 ** -9           # one TF32 product: 2 * 2^-11 per product, truncation
worst = {1: 0.0, 3: 0.0}
cases = 0


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def run(ta, tb, m, n, k, alpha, beta, pad=0, c_nan=False, mode=3, seed=0, offset=0):
    global cases
    rng = np.random.RandomState(seed + m * 7 + n * 3 + k)
    ra, ca = (m, k) if ta == "N" else (k, m)
    rb, cbn = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = max(ra, 1) + pad, max(rb, 1) + pad, max(m, 1) + pad
    A = np.zeros((lda, max(ca, 1)), order="F", dtype=np.float32); A[:ra, :ca] = rng.rand(ra, ca) - 0.5
    B = np.zeros((ldb, max(cbn, 1)), order="F", dtype=np.float32); B[:rb, :cbn] = rng.rand(rb, cbn) - 0.5
    C = np.zeros((ldc, n), order="F", dtype=np.float32); C[:m] = rng.rand(m, n) - 0.5
    if c_nan:
        C[:] = np.nan
    flatA = np.concatenate([np.zeros(offset, dtype=np.float32), A.reshape(-1, order="F")])   # offset = 1: a 4-byte-aligned base
    dA = torch.from_numpy(flatA).cuda()
    dB, dC = dev(B), dev(C)
    check(lib().candmc_set_f32_mode(mode))
    cb.csgemm(ta, tb, m, n, k, alpha, dA.data_ptr() + 4 * offset, lda, dB, ldb, beta, dC, ldc)
    torch.cuda.synchronize()
    check(lib().candmc_set_f32_mode(3))
    got = dC.cpu().numpy().reshape(n, ldc).T[:m].astype(np.float64)
    opA = (A[:ra, :ca] if ta == "N" else A[:ra, :ca].T).astype(np.float64)
    opB = (B[:rb, :cbn] if tb == "N" else B[:rb, :cbn].T).astype(np.float64)
    ref = alpha * (opA @ opB) + (beta * C[:m].astype(np.float64) if beta != 0.0 else 0.0)
    mag = abs(alpha) * (np.abs(opA) @ np.abs(opB)) + (abs(beta) * np.abs(C[:m]) if beta != 0.0 else 0.0)
    err = float((np.abs(got - ref) / np.maximum(mag, 1e-30)).max()) if m and n and (k or beta) else float(np.abs(got - ref).max())
    # the split scheme's bound (per product, K-independent) + the FP32 accumulation every single-precision GEMM has: the
    # tensor core adds K/8 partial sums per product term in FP32 (rounding unspecified), ~ sqrt(K) * 2^-24 of the summed terms
    tol = (TOL3 if mode == 3 else TOL1) + 2.0 * max(k, 1) ** 0.5 * 2.0 ** -24
    assert err <= tol, (ta, tb, m, n, k, alpha, beta, pad, mode, err, tol)
    if pad:   # rows below m are never written
        tail = dC.cpu().numpy().reshape(n, ldc).T[m:]
        assert np.array_equal(tail, C[m:]) or (c_nan and np.isnan(tail).all())
    worst[mode] = max(worst[mode], err)
    cases += 1
    return got


# every transpose combination (T,N is read in place by TMA, the others go through the K-major pack), whole tiles and ragged
# edges in m, n and k
for ta in "NT":
    for tb in "NT":
        run(ta, tb, 128, 128, 32, 1.0, 0.0)
        run(ta, tb, 150, 70, 36, -0.5, 1.0)
        run(ta, tb, 9, 300, 18, 2.0, -1.0, pad=2)
run("T", "N", 1, 1, 1, 1.0, 0.0, pad=3)
run("T", "N", 256, 128, 32, 1.0, 0.0)                 # two tiles on one CTA row: both accumulator buffers
run("T", "N", 128, 384, 32 * 7 + 5, 1.0, 1.0)         # more k-blocks than ring stages, ragged last one, three tiles
run("T", "N", 130, 130, 40, 1.0, 0.0, c_nan=True)     # beta == 0 never reads C
run("N", "N", 96, 40, 0, 1.0, 0.5)                    # k == 0: C = beta C
run("N", "N", 96, 40, 24, 0.0, 0.0, c_nan=True)       # alpha == 0, beta == 0
run("T", "N", 64, 64, 36, 1.0, 1.0, offset=1)         # 4-byte-aligned base: packed although already K-major
run("T", "N", 64, 64, 33, 1.0, 1.0, pad=1)            # pitch not a multiple of 4 floats: packed
run("T", "N", 128, 128, 32 * 9, 1.0, 0.0, mode=1)     # one TF32 product per FP32 product (no split warps, six stages)
run("N", "T", 150, 70, 36, -0.5, 1.0, mode=1)
# more tiles than CTAs (the simulator has 3 SMs): the persistent loop, both accumulator buffers, barrier phases wrapping round
a = run("T", "N", 128 * 5, 128 * 2, 64, 1.0, 0.0, seed=3)
b = run("T", "N", 128 * 5, 128 * 2, 64, 1.0, 0.0, seed=3)
assert np.array_equal(a, b)
if not SIM:   # the same on 148 SMs: several waves of tiles, long k, ragged everything, all four transpose combinations
    run("T", "N", 4096, 4096, 4096, 1.0, 0.0, seed=4)
    run("N", "N", 4096 + 77, 2048 + 13, 1024 + 5, 0.75, 1.0, seed=5, pad=3)
    run("N", "T", 3000, 5000, 777, -1.0, 0.5, seed=6)
    run("T", "T", 2048, 2048, 2048, 1.0, 0.0, seed=7, mode=1)
# randomised shapes / transposes / scalars / paddings / alignments (`--fuzz SEED COUNT`)
if "--fuzz" in sys.argv:
    i = sys.argv.index("--fuzz")
    frng = np.random.RandomState(int(sys.argv[i + 1]))
    for _ in range(int(sys.argv[i + 2])):
        big = 300 if SIM else 1500
        m, n, k = (int(frng.randint(1, big)) for _ in range(3))
        run("NT"[frng.randint(2)], "NT"[frng.randint(2)], m, n, k, float(frng.choice([1.0, -0.5, 2.0])),
            float(frng.choice([0.0, 1.0, -1.5])), pad=int(frng.randint(0, 4)), mode=int(frng.choice([3, 3, 1])),
            seed=int(frng.randint(1000)), offset=int(frng.randint(0, 2)))
speed = None
if not SIM and "--bench" in sys.argv:   # first numbers for the next GPU session (tools/r02_session.sh f32): device events, 5 launches
    speed = {}
    for nb in (4096, 8192, 16384):
        x = torch.rand(nb * nb, dtype=torch.float32, device="cuda") - 0.5
        y = torch.rand(nb * nb, dtype=torch.float32, device="cuda") - 0.5
        z = torch.empty(nb * nb, dtype=torch.float32, device="cuda")
        for mode in (3, 1):
            check(lib().candmc_set_f32_mode(mode))
            for ta, tb in (("T", "N"), ("N", "N")):
                cb.csgemm(ta, tb, nb, nb, nb, 1.0, x, nb, y, nb, 0.0, z, nb)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    cb.csgemm(ta, tb, nb, nb, nb, 1.0, x, nb, y, nb, 0.0, z, nb)
                e1.record()
                torch.cuda.synchronize()
                speed[f"n{nb}_{ta}{tb}_{mode}xtf32_tflops"] = 2.0 * nb ** 3 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        check(lib().candmc_set_f32_mode(3))
        torch.backends.cuda.matmul.allow_tf32 = False   # cross-check only: cuBLAS SGEMM on the same shapes
        xv, yv = x.view(nb, nb), y.view(nb, nb)
        torch.matmul(xv, yv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            torch.matmul(xv, yv)
        e1.record()
        torch.cuda.synchronize()
        speed[f"n{nb}_cublas_sgemm_crosscheck_tflops"] = 2.0 * nb ** 3 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del x, y, z
print(json.dumps({"cases": cases, "speed": speed, "max_rel_err_3xtf32": worst[3], "max_rel_err_1xtf32": worst[1], "tol_3xtf32": TOL3,
                  "launches": int(cb.launch_count())}))
