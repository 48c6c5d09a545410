"""Probe for tests/test_cpusim.py: makes the simulator's detectors fire on purpose.  TEST INFRASTRUCTURE."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import ctypes as C
import simtorch; simtorch.install()
import torch, candmc_b200 as cb
from candmc_b200._lib import lib, check
check(lib().candmc_init(0))
mode = sys.argv[1]
if mode == "oob":            # 8 x 8 copy into a 7 x 8 destination: the kernel writes past the allocation
    a = torch.zeros(64, dtype=torch.float64, device="cuda"); b = torch.zeros(56, dtype=torch.float64, device="cuda")
    check(lib().candmc_lda_cpy(8, 8, 8, 8, a.data_ptr(), b.data_ptr(), None)); torch.cuda.synchronize()
elif mode == "uninit":       # beta = 0 must not read C; beta = 1 on never-written device memory must poison the result
    a = torch.ones(16, dtype=torch.float64, device="cuda"); c = torch.empty(16, dtype=torch.float64, device="cuda")
    cb.cdgemm("N", "N", 4, 4, 4, 1.0, a, 4, a, 4, 0.0, c, 4); assert float(c.sum()) == 64.0
    ws = lib().cpusim_malloc; ws.restype = C.c_void_p; ws.argtypes = [C.c_size_t]
    raw = ws(128)
    cb.cdgemm("N", "N", 4, 4, 4, 1.0, a, 4, a, 4, 1.0, raw, 4)
    out = (C.c_double * 16).from_address(raw); assert all(x != x for x in out), list(out)
    print("uninit-ok")
elif mode == "gemm_range":   # a GEMM operand that runs past its allocation
    a = torch.ones(16, dtype=torch.float64, device="cuda"); c = torch.empty(16, dtype=torch.float64, device="cuda")
    cb.cdgemm("N", "N", 4, 4, 5, 1.0, a, 4, a, 5, 0.0, c, 4)
