"""TEST INFRASTRUCTURE: runs bench.py's single-GPU flow end to end with torch's CUDA surface faked on the
CPU and the product library replaced by the kernel-emulator build (tests/cudaemu).  It exists to catch
Python-level mistakes (names, argument lists, control flow, JSON shape) in bench.py without a GPU; the
numbers it prints are meaningless.  usage: bench_on_emulator.py [bench.py args]"""
import contextlib
import ctypes as C
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
EMU = os.path.join(HERE, "cudaemu", "_build")

_real_device = torch.device


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 1.0


def fake_device(*a, **k):
    return _real_device("cpu")


torch.device = fake_device
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.empty_cache = lambda: None
torch.cuda.Stream = FakeStream
torch.cuda.ExternalStream = FakeStream
torch.cuda.Event = FakeEvent
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.Tensor.pin_memory = lambda self: self

import torch.distributed as dist  # noqa: E402

_real_init = dist.init_process_group


def fake_init(backend=None, **kw):
    kw.pop("device_id", None)
    return _real_init("gloo", **kw)            # CPU tensors: gloo stands in for NCCL in the harness traffic


dist.init_process_group = fake_init
os.environ.setdefault("LIS_B200_TRANSPORT", "host")   # the library's own halo exchange: staged through host memory

import lis_b200  # noqa: E402
import lis_b200.capi as capi  # noqa: E402

_orig_cdll = C.CDLL


def load_library():
    return _orig_cdll(os.path.join(EMU, "liblis_emu.so"))


def load_kernels():
    capi.load_library = load_library
    return capi.load_kernels.__wrapped__() if hasattr(capi.load_kernels, "__wrapped__") else _load_kernels_orig()


_load_kernels_orig = capi.load_kernels
capi.load_library = load_library
lis_b200.load_library = load_library
lis_b200.load_kernels = lambda: _load_kernels_orig()
lis_b200.load_shim = lambda: lis_b200.Shim(os.path.join(EMU, "liblis_emu_shim.so"))

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    bench.main()
