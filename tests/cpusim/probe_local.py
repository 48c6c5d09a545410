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
elif mode in ("race", "ordered"):
    # stream 1 clears a buffer, stream 2 copies it: without an event between them the copy may run first — the lifo
    # scheduler makes sure it does; with cudaEventRecord / cudaStreamWaitEvent the result is the same under every policy
    sim = simtorch.sim()
    x = torch.ones(8, dtype=torch.float64, device="cuda"); y = torch.zeros(8, dtype=torch.float64, device="cuda")
    s1, s2, ev = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert sim.cudaStreamCreateWithFlags(C.byref(s1), 1) == 0 and sim.cudaStreamCreateWithFlags(C.byref(s2), 1) == 0
    assert sim.cudaEventCreateWithFlags(C.byref(ev), 2) == 0
    assert sim.cudaMemsetAsync(C.c_void_p(x.data_ptr()), 0, C.c_size_t(64), s1) == 0
    if mode == "ordered":
        assert sim.cudaEventRecord(ev, s1) == 0 and sim.cudaStreamWaitEvent(s2, ev, 0) == 0
    assert sim.cudaMemcpyAsync(C.c_void_p(y.data_ptr()), C.c_void_p(x.data_ptr()), C.c_size_t(64), 3, s2) == 0
    assert sim.cudaDeviceSynchronize() == 0
    print("copied", float(y.sum()))
