"""The GPU ingest pipeline's HOST logic on a machine without a GPU.

tests/test_gpu_reader.py is the parity suite of the reader through the C ABI (`-m gpu`, run on a B200 against the real
library).  Here the same tests run in a subprocess against tests/mock_cuda's library: the product's C++ host side
(csrc/host/*: job planning, pinned ring, copy groups, registrar, fetch threads, verify batching, result harvesting, the
worker, the writer, the HBM tier) compiled against a host-memory stand-in for the CUDA runtime, with plain-loop CPU stand-ins
for the cvk_* launchers instead of csrc/kernels.cu.  That checks the pipeline's bookkeeping (what lands where, which CRC is
compared with which manifest entry, slot reuse, cache revalidation, error paths) on every CPU run; it says nothing about the
kernels (tests/test_simt_emu.py runs the same suites against the kernel SOURCE on a SIMT shim; the B200 run checks the compiled code).  The mock is test infrastructure: nothing under curvine_b200/ can load it."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_reader_suite_against_the_mock_runtime():
    sys.path.insert(0, os.path.join(ROOT, "tests", "mock_cuda"))
    try:
        import build as mock_build
    finally:
        sys.path.pop(0)
    lib = mock_build.build()
    env = dict(os.environ, CV_TEST_MOCK_CUDA_LIB=lib)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_reader.py"),
                        os.path.join(ROOT, "tests", "test_zz_gpu_reader_faults.py"), os.path.join(ROOT, "tests", "test_arena_gpu.py"), os.path.join(ROOT, "tests", "test_gds_gpu.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-n", "6"], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 60, tail


def test_device_reader_releases_every_device_and_pinned_allocation_and_registration():
    """tests/mock_cuda/leak_check.py: mixed workload over four pipeline configurations incl. failed and abandoned reads and a device
    write; after every cv_fs_close the mock runtime must hold no device allocation, no pinned allocation and no registered range."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "mock_cuda"))
    try:
        import build as mock_build
    finally:
        sys.path.pop(0)
    env = dict(os.environ, CV_TEST_MOCK_CUDA_LIB=mock_build.build())
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock_cuda", "leak_check.py")], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "leak check ok" in r.stdout, r.stdout[-3000:]
