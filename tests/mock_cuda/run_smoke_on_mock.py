"""TEST INFRASTRUCTURE ONLY: __graft_entry__.smoke()'s control flow (files tier short-circuit + framed, arena tier) on a machine without a
GPU, against the mock library -- same scheme as run_bench_on_mock.py.  Says nothing about the kernels."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from curvine_b200 import _lib
_lib.LIB_PATH = os.environ["CV_TEST_MOCK_CUDA_LIB"]
import torch
def _cpu_dev(kw):
    d = kw.get("device")
    if d is not None and str(d).startswith("cuda"): kw["device"] = "cpu"
    return kw
for _name in ("empty", "zeros", "full", "arange", "tensor"):
    _orig = getattr(torch, _name)
    setattr(torch, _name, (lambda f: lambda *a, **kw: f(*a, **_cpu_dev(kw)))(_orig))
torch.cuda.is_available = lambda: True
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)
import __graft_entry__ as G
G.smoke()
