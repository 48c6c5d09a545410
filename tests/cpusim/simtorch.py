"""Test-side glue of the CPU functional simulator (tests/cpusim/sim_device.h) — TEST INFRASTRUCTURE ONLY.

`install()` makes the GPU test workers (tests/dist_worker.py, off_worker.py, redist_worker.py) runnable in the CPU suite,
unchanged:
  * candmc_b200's ctypes loader is pointed at tests/cpusim/_build/libcandmc_b200_cpusim.so (the product's host code and
    simple kernels compiled for the simulator; the product itself has no such switch — this is a monkeypatch from tests/);
  * `device="cuda"` tensors become views of simulator "device" memory (cpusim_malloc), so the library's host/device
    pointer classification and staging paths behave as on a GPU;
  * torch.distributed "nccl" becomes "gloo" (bootstrap of the unique id and the final pass/fail reduction only).
Nothing here is reachable from the product package; a worker only calls install() when CANDMC_CPUSIM=1 is set by a test.
"""
import ctypes as C
import os
import subprocess
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libcandmc_b200_cpusim.so")
_sim = None


def build(quiet=True):
    subprocess.check_call(["make", "-C", HERE, "-j8"] + (["-s"] if quiet else []))
    return SO


def sim():
    global _sim
    if _sim is None:
        if not os.path.exists(SO):
            build()
        _sim = C.CDLL(SO, mode=C.RTLD_GLOBAL)
        _sim.cpusim_malloc.restype = C.c_void_p
        _sim.cpusim_malloc.argtypes = [C.c_size_t]
        _sim.cpusim_is_device.restype = C.c_int
        _sim.cpusim_is_device.argtypes = [C.c_void_p]
        _sim.cpusim_check.restype = None
    return _sim


def _to_sim(t):
    """CPU tensor -> tensor of the same shape/dtype whose storage is simulator device memory."""
    import torch

    t = t.detach().contiguous()
    nbytes = max(t.numel() * t.element_size(), 8)
    ptr = sim().cpusim_malloc(nbytes)
    buf = (C.c_char * nbytes).from_address(ptr)
    out = torch.frombuffer(buf, dtype=t.dtype, count=t.numel()).reshape(t.shape)
    out.copy_(t)
    return out


def _to_pinned(t):
    """CPU tensor -> tensor in simulator PINNED host memory (cudaMallocHost): copies to and from it are truly asynchronous
    under the deferred stream scheduler, as on a GPU, while pageable memory makes them host-synchronous."""
    import torch

    t = t.detach().contiguous()
    nbytes = max(t.numel() * t.element_size(), 8)
    ptr = C.c_void_p()
    assert sim().cudaMallocHost(C.byref(ptr), C.c_size_t(nbytes)) == 0
    buf = (C.c_char * nbytes).from_address(ptr.value)
    out = torch.frombuffer(buf, dtype=t.dtype, count=t.numel()).reshape(t.shape)
    out.copy_(t)
    return out


def _is_cuda_device(d):
    return d is not None and str(d).startswith("cuda")


def install():
    import torch
    import torch.distributed as dist

    import candmc_b200._lib as L

    if getattr(torch, "_cpusim_installed", False):
        return
    sim()
    L._SO = SO          # candmc_b200.lib() now loads the simulator build (same C ABI, same symbols)
    L._lib = None

    def factory(fn):
        def wrapped(*a, **k):
            dev = k.pop("device", None)
            out = fn(*a, **k)
            return _to_sim(out) if _is_cuda_device(dev) else out
        return wrapped

    for name in ("full", "zeros", "empty", "ones", "tensor", "arange", "rand", "randn"):
        setattr(torch, name, factory(getattr(torch, name)))

    def like(fn):
        def wrapped(src, *a, **k):
            out = fn(src, *a, **k)
            return _to_sim(out) if sim().cpusim_is_device(src.data_ptr()) else out
        return wrapped

    for name in ("zeros_like", "empty_like", "ones_like", "full_like"):
        setattr(torch, name, like(getattr(torch, name)))

    torch.Tensor.cuda = lambda self, *a, **k: _to_sim(self)
    torch.cuda.synchronize = lambda *a, **k: sim().cpusim_check()
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: 1
    torch.cuda.get_device_properties = lambda *a, **k: types.SimpleNamespace(   # a PCI function that does not exist: no binding
        name="cpusim", pci_domain_id=0xFFFF, pci_bus_id=0xFF, pci_device_id=0x1F, multi_processor_count=148)
    torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0)

    import time as _time

    class SimEvent:
        """torch.cuda.Event on the simulator: host wall clock at record() (the simulator has no device timeline)."""
        def __init__(self, enable_timing=False, **_):
            self.t = None

        def record(self, stream=None):
            sim().cpusim_check()
            self.t = _time.perf_counter()

        def synchronize(self):
            pass

        def query(self):
            return True

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    torch.cuda.Event = SimEvent
    torch.Tensor.pin_memory = lambda self, *a, **k: _to_pinned(self)

    def pinned(fn):
        def wrapped(*a, **k):
            pin = k.pop("pin_memory", None)
            out = fn(*a, **k)
            return _to_pinned(out) if pin else out
        return wrapped

    for name in ("full", "zeros", "empty", "ones"):
        setattr(torch, name, pinned(getattr(torch, name)))

    # results of torch operations on simulated-device tensors live on the simulated device too (bench.py's cuBLAS cross-check
    # `fb.view(..) @ fa.view(..)`, reductions, clones): a function mode moves fresh outputs over
    from torch.overrides import TorchFunctionMode

    def _flat(xs):
        for x in xs:
            if isinstance(x, torch.Tensor):
                yield x
            elif isinstance(x, (list, tuple)):
                yield from _flat(x)

    _METADATA = {"data_ptr", "numel", "size", "dim", "stride", "element_size", "is_contiguous", "__get__", "storage_offset",
                 "__len__", "nelement", "is_floating_point"}

    class SimDeviceMode(TorchFunctionMode):
        busy = False

        def __torch_function__(self, func, types, args=(), kwargs=None):
            kwargs = dict(kwargs or {})
            if _is_cuda_device(kwargs.get("device")):        # any factory: torch.eye(..., device="cuda"), x.to(device="cuda")
                kwargs.pop("device")
                kwargs.pop("pin_memory", None)
                out = func(*args, **kwargs)
                if isinstance(out, torch.Tensor) and not SimDeviceMode.busy:
                    SimDeviceMode.busy = True
                    try:
                        return _to_sim(out)
                    finally:
                        SimDeviceMode.busy = False
                return out
            if kwargs.pop("pin_memory", None) and not SimDeviceMode.busy:
                SimDeviceMode.busy = True
                try:
                    return _to_pinned(func(*args, **kwargs))
                finally:
                    SimDeviceMode.busy = False
            # torch's own operations are stream-ordered with the library's work on a GPU; here they run eagerly on the host, so
            # everything queued on the simulated device has to have happened first
            name = getattr(func, "__name__", "")
            if name in _METADATA:
                return func(*args, **kwargs)
            if not SimDeviceMode.busy and any(t.numel() and sim().cpusim_is_device(t.data_ptr())
                                              for t in _flat(list(args) + list(kwargs.values()))):
                sim().cpusim_check()
            out = func(*args, **kwargs)
            if SimDeviceMode.busy or not isinstance(out, torch.Tensor) or out.numel() == 0:
                return out
            if name in ("cpu", "numpy", "item", "tolist"):
                return out
            if sim().cpusim_is_device(out.data_ptr()):
                return out
            if any(t.numel() and sim().cpusim_is_device(t.data_ptr()) for t in _flat(list(args) + list((kwargs or {}).values()))):
                SimDeviceMode.busy = True
                try:
                    return _to_sim(out)
                finally:
                    SimDeviceMode.busy = False
            return out

    torch._cpusim_mode = SimDeviceMode()
    torch._cpusim_mode.__enter__()

    real_init = dist.init_process_group

    def init_pg(backend=None, *a, **k):
        k.pop("device_id", None)
        return real_init("gloo", *a, **k)

    dist.init_process_group = init_pg
    torch._cpusim_installed = True
