"""TEST INFRASTRUCTURE ONLY: runs bench.py's own ("ours") arm on a machine without a GPU, against the mock library, to check the
bench's CONTROL FLOW and the JSON line it prints (keys, types, bookkeeping such as h2d/d2h byte counts, warm-up / registration
handling) on every CPU run.  The numbers it prints are meaningless (host memcpy, bytewise CRC) and are never recorded anywhere.
torch is told that "cuda" tensors are CPU tensors and that events are wall clocks; the product library is replaced by the mock."""
import os
import runpy
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from curvine_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.environ["CV_TEST_MOCK_CUDA_LIB"]

import torch  # noqa: E402


def _cpu_dev(kw):
    d = kw.get("device")
    if d is not None and str(d).startswith("cuda"):
        kw["device"] = "cpu"
    return kw


for _name in ("empty", "zeros", "full", "arange", "tensor"):
    _orig = getattr(torch, _name)
    setattr(torch, _name, (lambda f: lambda *a, **kw: f(*a, **_cpu_dev(kw)))(_orig))
torch.Tensor.cuda = lambda self, *a, **k: self


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.Event = _Event
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.current_device = lambda: 0
torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)

import torch.distributed as _dist  # noqa: E402

_init_pg = _dist.init_process_group
_dist.init_process_group = lambda backend=None, **kw: _init_pg("gloo", **{k: v for k, v in kw.items() if k != "device_id"})  # gloo stands in for NCCL

sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
