"""Host-side reader surface (cv_read / cv_read_full: the reference's Reader trait, A1-A3) against the oracle's reference-shaped CPU
reader, on CPU only.  The reference's curvine-bench loop without its checksum: read_full(128 KiB buffer) until EOF
(curvine_bench.rs:212-236; the crc32 is left out on both sides because the Python harness has no PCLMUL crc32).  Shows what a Rust UnifiedReader::Cuda::read_chunk0 bound to cv_read costs for HOST destinations, with the
prefetch threads of fs_reader_buffer.rs restated (read_chunk_num = 8) and without them (read_chunk_num = 1).
    python tools/host_read_bench.py [--gib 1]"""
import argparse
import ctypes
import json
import os
import shutil
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--block", type=int, default=4 << 20)
    a = ap.parse_args()
    from curvine_b200 import _lib, fs as F
    from oracle import clib, layout
    n = int(a.gib * (1 << 30)) // a.block * a.block
    d = tempfile.mkdtemp(prefix="cvhost", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    res = {"file_bytes": n, "block_bytes": a.block, "host_cpus": os.cpu_count()}
    try:
        with F.MiniWorker(["[MEM]" + d], hostname="localhost") as w:
            man = w.create_file("/h", 7100, n, a.block, threads=min(8, os.cpu_count() or 4))
            ids = [layout.create_block_id(7100, i) for i in range(n // a.block)]
            L = _lib.lib()
            buf = ctypes.create_string_buffer(131072)
            for sc in (True, False):
                for chunk_num, parallel in ((8, 1), (1, 1), (8, 4)):
                    with F.CurvineFileSystem(F.client_conf(hostname="localhost", short_circuit=sc, read_chunk_num=chunk_num, read_parallel=parallel)) as fs:
                        fs.load_namespace(man)
                        best = 0.0
                        for rep in range(3):
                            r = fs.open("/h")
                            t0 = time.perf_counter()
                            total, s, got = 0, 0, ctypes.c_int64()
                            while True:
                                rc = L.cv_read_full(r._h, buf, 131072, ctypes.byref(got))
                                assert rc == 0
                                if got.value == 0:
                                    break
                                total += got.value  # no checksum on either side: the Python harness has no PCLMUL crc32 to match the oracle's
                            dt = time.perf_counter() - t0
                            r.complete()
                            assert total == n
                            best = max(best, n / dt / 1e9)
                        res["product_%s_chunknum%d_parallel%d_GBps" % ("short_circuit" if sc else "framed", chunk_num, parallel)] = round(best, 2)
                for parallel in (1, 4):
                    best = 0.0
                    for rep in range(3):
                        t0 = time.perf_counter()
                        got, cks, threads = clib.cpu_read_file(w.port, sc, n, a.block, ids, 131072, 8, parallel, 131072, n, -1)
                        best = max(best, got / (time.perf_counter() - t0) / 1e9)
                    res["oracle_cpu_reader_%s_parallel%d_GBps" % ("short_circuit" if sc else "framed", parallel)] = round(best, 2)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
