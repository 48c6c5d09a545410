"""The product's pack kernels executed on the simulator against numpy, bit for bit: the tiled lda_cpy / scaled lda_cpy kernels
(double2 and scalar paths, tall and short columns, ragged tiles) and both transpose kernels — the TMA load -> turn in shared
memory -> TMA store kernel (128-byte swizzle of the landed boxes, clipped edge tiles) and the LDG/STG kernel it falls back to
for unaligned operands.  TEST INFRASTRUCTURE.  Prints one JSON line."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import simtorch; simtorch.install()  # noqa: E402,E702
import numpy as np  # noqa: E402
import torch  # noqa: E402
import candmc_b200 as cb  # noqa: E402
from candmc_b200._lib import lib, check  # noqa: E402

check(lib().candmc_init(0))
rng = np.random.RandomState(7)
cases = 0


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def back(t, ld, cols):
    return t.cpu().numpy().reshape(cols, ld).T


for (nrow, ncol, lda, ldb, off) in [(64, 32, 64, 64, 0), (48, 20, 50, 60, 0), (7, 5, 9, 8, 0), (2050, 3, 2050, 2052, 0),
                                    (2, 700, 4, 2, 0), (1, 1, 1, 1, 0), (130, 9, 131, 130, 0), (16, 300, 18, 16, 1),
                                    (4100, 2, 4100, 4100, 0)]:
    A = np.asfortranarray(rng.rand(lda, ncol)); B = np.asfortranarray(rng.rand(ldb, ncol))
    fa = torch.from_numpy(np.concatenate([np.zeros(off), A.reshape(-1, order="F")])).cuda()
    dB = dev(B)
    cb.lda_cpy(nrow, ncol, lda, ldb, fa.data_ptr() + 8 * off, dB); torch.cuda.synchronize()
    want = B.copy(); want[:nrow] = A[:nrow]
    assert np.array_equal(back(dB, ldb, ncol), want), ("copy", nrow, ncol, lda, ldb, off)
    dB = dev(B)
    cb.lda_cpy(nrow, ncol, lda, ldb, fa.data_ptr() + 8 * off, dB, a=0.5, b=0.25); torch.cuda.synchronize()
    want = B.copy(); want[:nrow] = B[:nrow] * 0.25 + A[:nrow] * 0.5
    assert np.abs(back(dB, ldb, ncol) - want).max() <= 1e-15, ("axpby", nrow, ncol, lda, ldb, off)
    cases += 2

for tma in (1, 0):
    check(lib().candmc_debug_transpose_tma(tma))
    for (rows, cols, lda, ldb) in [(64, 64, 64, 64), (128, 192, 128, 192), (70, 130, 72, 130), (5, 300, 6, 300), (257, 3, 258, 4),
                                   (1, 1, 2, 2), (200, 100, 201, 100), (96, 96, 96, 97), (330, 77, 330, 78)]:
        A = np.asfortranarray(rng.rand(lda, cols)); B = np.asfortranarray(rng.rand(ldb, rows))
        dA, dB = dev(A), dev(B)
        before = cb.launch_count()
        cb.transpose(rows, cols, dA, lda, dB, ldb); torch.cuda.synchronize()
        assert cb.launch_count() == before + 1
        want = B.copy(); want[:cols] = A[:rows].T
        assert np.array_equal(back(dB, ldb, rows), want), ("transpose", tma, rows, cols, lda, ldb)
        cases += 1
check(lib().candmc_debug_transpose_tma(1))
print(json.dumps({"cases": cases, "launches": int(cb.launch_count())}))
