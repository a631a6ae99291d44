"""cv_read_device on a stream of the caller's: the read is ordered ON that stream (round 1's review found `run_jobs` ordering only its
end).  Work the caller enqueued BEFORE the call (here: fills of the destination, behind something slow) must not land on top of the
file's bytes, and work enqueued AFTER it (a copy out of the destination) must see the whole file -- without any host-side
synchronisation in between.  Written after round 2's last GPU run; sorts late on purpose.  On the host-side stand-ins (tests/mock_cuda with
MOCK_CUDA_ASYNC=1, tests/simt_emu) the caller's stream is a stand-in stream with its own thread and random pauses."""
import ctypes
import os
import shutil
import tempfile

import numpy as np
import pytest

from curvine_b200 import _lib, fs as F
from oracle import clib, synth

pytestmark = pytest.mark.gpu
MOCK = bool(os.environ.get("CV_TEST_MOCK_CUDA_LIB"))


class CallerStream:
    """a non-default stream + the three operations the test enqueues on it"""

    def __init__(self, torch):
        self.torch = torch
        if MOCK:
            self.L = _lib.lib()
            self.h = ctypes.c_void_p()
            assert self.L.cudaStreamCreateWithFlags(ctypes.byref(self.h), 1) == 0
            self.handle = self.h.value
        else:
            self.s = torch.cuda.Stream()
            self.handle = self.s.cuda_stream

    def fill(self, t, value):
        if MOCK:
            assert self.L.cudaMemsetAsync(ctypes.c_void_p(t.data_ptr()), value, ctypes.c_size_t(t.numel()), self.h) == 0
        else:
            with self.torch.cuda.stream(self.s):
                t.fill_(value)

    def copy(self, dst, src):
        if MOCK:
            assert self.L.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(src.numel()), 3, self.h) == 0
        else:
            with self.torch.cuda.stream(self.s):
                dst.copy_(src, non_blocking=True)

    def synchronize(self):
        if MOCK:
            assert self.L.cudaStreamSynchronize(self.h) == 0
        else:
            self.s.synchronize()

    def close(self):
        if MOCK:
            self.L.cudaStreamDestroy(self.h)


@pytest.mark.parametrize("sc,arena", [(True, False), (True, True), (False, False)])
def test_read_device_is_ordered_on_the_callers_stream(cuda, sc, arena):
    import torch
    n, bs, ino = (24 << 20) - 999, 1 << 20, 8301
    d = tempfile.mkdtemp(prefix="cvso", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    cs = None
    try:
        extra = 'mem_arena = true\narena_segment = "16MB"\n' if arena else ""
        with F.MiniWorker([("[MEM:64MB]" if arena else "[MEM]") + d + "/m"], extra_worker=extra) as w:
            man = w.create_file("/so", ino, n, bs, threads=4)
            want = synth.file_bytes(ino, n, bs)
            conf = F.client_conf(short_circuit=sc, b200='fetch_threads = 4\nverify_batch = 4\npinned_slots = 12\ncopy_group = 2\nzero_copy = %s\ngpu_chunk_size = "1MB"\n'
                                                      'register_threads = 2\narena_register_slice = "4MB"\narena_preregister = ["%s/m"]\n' % ("true" if arena else "false", d))
            with F.CurvineFileSystem(conf) as fs:
                fs.load_namespace(man)
                if arena:
                    fs.preregister()
                    fs.wait_registered()
                cs = CallerStream(torch)
                dst = torch.zeros(n, dtype=torch.uint8, device=cuda)
                out = torch.zeros(n, dtype=torch.uint8, device=cuda)
                slow = torch.zeros(64 << 20, dtype=torch.uint8, device=cuda)
                torch.cuda.synchronize()  # the allocations' own fills ran on the default stream, which the caller's stream does not wait for
                for rnd in range(3):
                    for _ in range(6):
                        cs.fill(slow, rnd)          # keeps the stream busy: the fills below are still pending when read_device is called
                    for v in (0xE0, 0xE1, 0xE2 + rnd):
                        cs.fill(dst, v)             # pending writes to the destination, enqueued BEFORE the read
                    r = fs.open("/so")
                    assert r.read_device(dst.data_ptr(), n, cs.handle) == n
                    cs.copy(out, dst)               # enqueued AFTER the read, same stream, no host synchronisation in between
                    cs.synchronize()
                    assert out.cpu().numpy().tobytes() == want, "round %d: the read is not ordered on the caller's stream" % rnd
                    s, bad, ver = r.verify()
                    r.complete()
                    assert bad == 0 and s == int(clib.crc_blocks(1, np.frombuffer(want, dtype=np.uint8), bs).astype(np.uint64).sum())
                    out.zero_()
    finally:
        if cs is not None:
            cs.close()
        shutil.rmtree(d, ignore_errors=True)


def test_device_reader_outlives_its_filesystem_handle(cuda):
    """cv_fs_close drops the handle's reference only (include/curvine_b200.h; the reference's FsReader holds an Arc<FsContext>): a reader
    keeps the context -- pinned ring, streams, connection pool -- alive until it is closed itself."""
    import torch
    n, bs, ino = (7 << 20) + 123, 1 << 20, 8302
    d = tempfile.mkdtemp(prefix="cvlt", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        with F.MiniWorker(["[MEM]" + d + "/m"]) as w:
            man = w.create_file("/lt", ino, n, bs, threads=4)
            want = synth.file_bytes(ino, n, bs)
            for sc in (True, False):
                fs = F.CurvineFileSystem(F.client_conf(short_circuit=sc, b200='fetch_threads = 2\nverify_batch = 2\npinned_slots = 8\ncopy_group = 1\ngpu_chunk_size = "1MB"\n'))
                fs.load_namespace(man)
                r = fs.open("/lt")
                dst = torch.zeros(n, dtype=torch.uint8, device=cuda)
                st = torch.cuda.current_stream().cuda_stream
                assert r.read_device(dst.data_ptr(), 2 * bs, st) == 2 * bs
                fs.close()                       # the handle goes first
                assert r.read_device(dst.data_ptr() + 2 * bs, n, st) == n - 2 * bs
                s, bad, ver = r.verify()
                torch.cuda.synchronize()
                assert bad == 0 and dst.cpu().numpy().tobytes() == want
                r.complete()                     # the last holder: the pipeline is torn down here
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_curvinefs_client_read_tensor(cuda):
    """curvine_b200/curvinefs.py (the reference's Python SDK names on the new ABI): ranges of a file as uint8 CUDA tensors."""
    import torch
    from curvine_b200 import curvinefs
    dev = "cpu" if MOCK else None  # host-side stand-ins: "device memory" is host memory, so the binding is asked for a CPU tensor
    n, bs, ino = (6 << 20) + 77, 1 << 20, 8303
    d = tempfile.mkdtemp(prefix="cvpy", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        with F.MiniWorker(["[MEM]" + d + "/m"]) as w:
            man = w.create_file("/py/a", ino, n, bs, threads=4)
            want = synth.file_bytes(ino, n, bs)
            open(d + "/ns", "w").write(man)
            open(d + "/conf.toml", "w").write('namespace_manifest = "%s/ns"\n' % d + F.client_conf())
            c = curvinefs.CurvineClient(d + "/conf.toml")
            t = c.read_range_tensor("/py/a", device=dev)
            assert (MOCK or t.is_cuda) and t.dtype == torch.uint8 and t.cpu().numpy().tobytes() == want
            t = c.read_range_tensor("/py/a", bs + 5, 2 * bs, device=dev)
            assert t.cpu().numpy().tobytes() == want[bs + 5:3 * bs + 5]
            r = c.open("/py/a")
            assert r.read(0, 100) == want[:100]
            assert r.read_tensor(1000, device=dev).cpu().numpy().tobytes() == want[100:1100]
            assert torch.from_dlpack(r.read_tensor(device=dev)).cpu().numpy().tobytes() == want[1100:]
            r.close()
            c.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
