"""The hot kernel reading B in the SUMMA pipeline's chunk-major layout through one tensor map (candmc_dgemm_chunked_b,
candmc_b200/csrc/gemm_f64.cu: producer coordinate (k0 mod kc, n0 + (k0 div kc) * n)) against the SAME kernel on the plain k x n
matrix (candmc_dgemm: identical tiles, k order and split-K decisions, so the two results must agree BIT FOR BIT) and against a
numpy float64 product (relative Frobenius error <= 10 k eps, BASELINE's tolerance).

    python tests/bchunk_worker.py                    on a B200 (tests/test_zz_redist_gpu.py), larger cases and speeds included
    CANDMC_CPUSIM=1 python tests/bchunk_worker.py    on the CPU simulator's PTX emulation (tests/test_cpusim.py)

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
    os.environ.setdefault("CPUSIM_GEMM", "device")   # candmc_dgemm on the emulated kernel too: the bit-for-bit partner
import numpy as np  # noqa: E402
import torch  # noqa: E402
import candmc_b200 as cb  # noqa: E402
from candmc_b200._lib import lib, check  # noqa: E402

check(lib().candmc_init(0))
EPS = 2.220446049250313e-16
cases, worst, mismatches = 0, 0.0, []


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def chunk_major(B, kc):
    """k x n column-major -> the chunks one behind the other, each kc x n column-major with ld = kc"""
    k, n = B.shape
    return np.concatenate([np.asfortranarray(B[t * kc:(t + 1) * kc]).reshape(-1, order="F") for t in range(k // kc)])


def run(ta, m, n, k, kc, alpha, beta, pad=0, seed=0):
    global cases, worst
    rng = np.random.RandomState(seed + 13 * m + 7 * n + k + kc)
    ra, ca = (m, k) if ta == "N" else (k, m)
    lda, ldc = ra + pad, m + pad
    A = np.zeros((lda, ca), order="F"); A[:ra] = rng.rand(ra, ca) - 0.5
    B = np.asfortranarray(rng.rand(k, n) - 0.5)
    C0 = np.full((ldc, n), np.nan, order="F"); C0[:m] = rng.rand(m, n) - 0.5
    dA, dB, dBc = dev(A), dev(B), torch.from_numpy(chunk_major(B, kc)).cuda()
    dC1, dC2 = dev(C0), dev(C0)
    if beta == 0.0:
        dC1.fill_(float("nan")); dC2.fill_(float("nan"))   # beta = 0 must not read C
    cb.cdgemm(ta, "N", m, n, k, alpha, dA, lda, dB, k, beta, dC1, ldc)
    cb.cdgemm_chunked_b(ta, m, n, k, kc, alpha, dA, lda, dBc, beta, dC2, ldc)
    torch.cuda.synchronize()
    g1 = dC1.cpu().numpy().reshape(n, ldc).T[:m]
    g2 = dC2.cpu().numpy().reshape(n, ldc).T[:m]
    opA = A[:ra].T if ta == "T" else A[:ra]
    ref = alpha * (opA @ B) + (beta * C0[:m] if beta != 0.0 else 0.0)
    err = float(np.linalg.norm(g2 - ref) / max(np.linalg.norm(ref), 1e-300))
    worst = max(worst, err)
    tag = f"{ta}N m={m} n={n} k={k} kc={kc} alpha={alpha} beta={beta} pad={pad}"
    if not np.array_equal(g1, g2):
        mismatches.append(tag + f": differs from the plain-layout launch (max {np.nanmax(np.abs(g1 - g2)):.2e})")
    if not (err <= 10 * max(k, 1) * EPS):
        mismatches.append(tag + f": rel. Frobenius error {err:.2e}")
    cases += 1


# whole and ragged tiles in m and n (n ragged: the box of the last tile column reaches into the next chunk's columns — those
# results belong to columns that are never stored), one chunk, many chunks, chunks of one k-tile, both layouts of A, beta 0 / 1 / other
run("N", 128, 128, 64, 16, 1.0, 0.0)
run("N", 128, 128, 64, 64, 1.0, 1.0)
run("N", 200, 72, 96, 32, 1.0, 1.0, pad=2)
run("T", 130, 250, 96, 48, -0.5, 0.75)
run("N", 64, 300, 128, 16, 2.0, 0.0, pad=4)
run("T", 256, 192, 48, 16, 1.0, 1.0)
run("N", 384, 128, 32, 32, 1.0, 1.0)          # few tiles, short k: the split-K decision is the same for both launches
if not SIM:
    run("N", 2048, 2048, 2048, 256, 1.0, 1.0)
    run("N", 4096, 1000, 4096, 512, 1.0, 0.0)
    run("T", 1024, 4096, 8192, 1024, 1.0, 1.0)
    run("N", 8192, 8192, 7168, 1024, 1.0, 1.0)   # the merged launch of a b = 8192 panel: seven of eight 1024-deep chunks

# loud errors, not a wrong layout read silently
for bad in (dict(k=96, kc=24), dict(k=100, kc=32), dict(k=64, kc=16, lda_odd=True)):
    m = n = 128
    k, kc = bad["k"], bad["kc"]
    lda = m + (1 if bad.get("lda_odd") else 0)
    x = torch.zeros(max(lda, k) * max(k, n) + 16, dtype=torch.float64, device="cuda")
    rc = lib().candmc_dgemm_chunked_b(b"N", m, n, k, kc, 1.0, x.data_ptr(), lda, x.data_ptr(), 0.0, x.data_ptr(), m, None)
    if rc == 0:
        mismatches.append(f"accepted k={k} kc={kc} lda={lda}")
    cases += 1

speeds = []
if not SIM and "--bench" in sys.argv:
    for b, kc, nch in ((8192, 1024, 7), (16384, 2048, 7)):
        k = kc * nch
        A = torch.rand(b * k, dtype=torch.float64, device="cuda")
        B = torch.rand(k * b, dtype=torch.float64, device="cuda")
        Cm = torch.zeros(b * b, dtype=torch.float64, device="cuda")
        def timed(fn, reps=3):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        t_merged = timed(lambda: cb.cdgemm_chunked_b("N", b, b, k, kc, 1.0, A, b, B, 1.0, Cm, b))
        def per_chunk():
            for t in range(nch):   # what the sweep launches today: one multiply per chunk (chunk t of B: kc x b, ld = kc)
                cb.cdgemm("N", "N", b, b, kc, 1.0, A[t * kc * b:], b, B[t * kc * b:], kc, 1.0, Cm, b)
        t_chunks = timed(per_chunk)
        fl = 2.0 * b * b * k
        speeds.append({"b": b, "kc": kc, "chunks": nch, "one_launch_ms": t_merged, "per_chunk_launches_ms": t_chunks,
                       "one_launch_tflops": fl / t_merged / 1e9, "per_chunk_tflops": fl / t_chunks / 1e9})

print(json.dumps({"cases": cases, "max_rel_frobenius": worst, "failures": mismatches, "launches": cb.launch_count(),
                  "speeds": speeds}))
sys.exit(1 if mismatches else 0)
