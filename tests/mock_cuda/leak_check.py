"""TEST INFRASTRUCTURE ONLY: resource accounting of the device reader on the mock runtime.  After a mixed workload (zero-copy with
background registration, pinned-ring reads, framed reads, small-file batches, a device write, a failed read) every filesystem handle
is closed and the mock is asked what is still alive: device allocations, pinned allocations, registered host ranges."""
import ctypes
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from curvine_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.environ["CV_TEST_MOCK_CUDA_LIB"]
from curvine_b200 import fs as F  # noqa: E402
from oracle import synth  # noqa: E402


def counters():
    a = (ctypes.c_uint64 * 6)()
    _lib.lib().mock_cuda_counters(a)
    return dict(zip(["memcpy_calls", "memcpy_bytes", "registered_ranges", "register_calls", "device_allocs", "pinned_allocs"], a))


def main():
    base = counters()
    assert base["device_allocs"] == 0 and base["pinned_allocs"] == 0 and base["registered_ranges"] == 0, base
    d = tempfile.mkdtemp(prefix="cvleak", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        with F.MiniWorker(["[MEM]" + d]) as w:
            n, bs = (20 << 20) + 123, 1 << 20
            man = w.create_file("/a", 8001, n, bs) + "".join(w.create_file("/s%d" % i, 8100 + i, 300000, 1 << 20) for i in range(40))
            want = synth.file_bytes(8001, n, bs)
            for conf_kw in (dict(sc=True, b200='zero_copy = true\nregister_threads = 2\ncopy_group = 2\nfetch_threads = 4\nverify_batch = 4\nregister_cache = "64MB"\n'),
                            dict(sc=True, b200='zero_copy = true\nregister_threads = 0\ncopy_group = 4\nfetch_threads = 4\nverify_batch = 4\nregister_cache = "8MB"\n'),
                            dict(sc=True, b200='zero_copy = false\nfetch_threads = 4\nverify_batch = 4\n'),
                            dict(sc=False, b200='gpu_chunk_size = "256KB"\nfetch_threads = 4\nverify_batch = 4\n')):
                fs = F.CurvineFileSystem(F.client_conf(short_circuit=conf_kw["sc"], b200=conf_kw["b200"]))
                fs.load_namespace(man)
                dst = np.zeros(n, dtype=np.uint8)
                for rep in range(3):
                    r = fs.open("/a")
                    assert r.read_device(dst.ctypes.data, n, 0) == n
                    assert r.verify()[1] == 0 and dst.tobytes() == want
                    r.complete()
                    fs.wait_registered()
                big = np.zeros(40 * 300000, dtype=np.uint8)
                tot, s, bad, ver = fs.read_many_device(["/s%d" % i for i in range(40)], big.ctypes.data, [i * 300000 for i in range(40)], big.size, 0)
                assert tot == 40 * 300000 and bad == 0 and ver == 40
                r = fs.open("/a")  # a reader abandoned with results pending, and one failed call
                r.read_device(dst.ctypes.data, 5 << 20, 0)
                r.complete()
                os.rename(os.path.join(d, "curvine"), os.path.join(d, "hidden"))
                r = fs.open("/s0")
                try:
                    r.read_device(big.ctypes.data, 300000, 0)
                    r.verify()
                except F.FsError:
                    pass
                try:
                    r.complete()
                except F.FsError:
                    pass
                os.rename(os.path.join(d, "hidden"), os.path.join(d, "curvine"))
                mid = counters()
                fs.close()
                after = counters()
                assert after["device_allocs"] == 0 and after["pinned_allocs"] == 0 and after["registered_ranges"] == 0, (conf_kw, mid, after)
            # HBM tier: two resident blocks (one re-loaded, one evicted by capacity would need a conf; here: manual loads), framed read, then the
            # worker goes away with them
            from oracle import layout
            for i in (0, 1, 1):
                w.hbm_load(layout.create_block_id(8001, i), 0)
            assert counters()["device_allocs"] == 2
            fs = F.CurvineFileSystem(F.client_conf(short_circuit=False))
            fs.load_namespace(man)
            r = fs.open("/a")
            assert r.read_full(3 << 20) == want[:3 << 20]
            r.complete()
            fs.close()
            assert w.hbm_stats()["reads_from_hbm"] >= 2
            # write path: device bytes -> worker -> read back
            fs = F.CurvineFileSystem(F.client_conf(short_circuit=False))
            src = np.frombuffer(synth.file_bytes(8200, 3 << 20, 1 << 20), dtype=np.uint8).copy()
            wr = fs.create("/w", 8200, 1 << 20, w.port)
            wr.write_device(src.ctypes.data, src.size, 0)
            wr.complete()
            r = fs.open("/w")
            out = np.zeros(src.size, dtype=np.uint8)
            assert r.read_device(out.ctypes.data, out.size, 0) == out.size and r.verify()[1] == 0 and (out == src).all()
            r.complete()
            fs.close()
        final = counters()
        assert final["device_allocs"] == 0 and final["pinned_allocs"] == 0 and final["registered_ranges"] == 0, final
        print("leak check ok:", final)
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
