"""Config C5: 4096 x 256 KiB single-block files, mem tier, random order, 1 GPU.  FUSE-shaped access per file:
open -> fuse_read(0, 256 KiB) -> close.  Reports files/s, GB/s and p50/p99 per-file latency for
  gpu   cv_open + cv_fuse_read_device (bytes into HBM scratch, scattered into 64 x 4 KiB page buffers by K3, CRC-verified) + close
  cpu   the oracle's reference-shaped CPU reader (oracle/cpu_reader.c) reading each file into host memory with crc32."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main_bench(args, emit):
    """bench.py --config c5: ONE line in the bench contract.  value = files/s of the per-file path (one FUSE-shaped call per file)."""
    a = argparse.Namespace(files=int(os.environ.get("CV_C5_FILES", "4096")), size=256 * 1024, zero_copy=1, arena=1 if args.tier == "arena" else 0, emit=emit)
    return run(a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4096)
    ap.add_argument("--size", type=int, default=256 * 1024)
    ap.add_argument("--zero-copy", type=int, default=1)
    ap.add_argument("--arena", type=int, default=1)
    a = ap.parse_args()
    a.emit = None
    return run(a)


def run(a):
    import ctypes
    import torch
    from curvine_b200 import fs as F
    from oracle import clib, layout
    torch.cuda.set_device(0)
    d = tempfile.mkdtemp(prefix="cvc5_", dir="/dev/shm")
    res = {}
    try:
        from curvine_b200 import _lib
        cap = a.files * a.size + (64 << 20)
        extra = 'mem_arena = true\narena_segment = "256MB"\narena_numa = [%d]\n' % int(_lib.lib().cv_gpu_numa_node(0)) if a.arena else ""
        with F.MiniWorker([("[MEM:%d]" % cap if a.arena else "[MEM]") + d], hostname="localhost", extra_worker=extra) as w:
            mans = [w.create_file("/small/f%d" % i, 100000 + i, a.size, a.size, threads=1) for i in range(a.files)]
            order = np.random.default_rng(7).permutation(a.files)  # Fisher-Yates, seed 7
            conf = F.client_conf(hostname="localhost", short_circuit=True,
                                 b200='zero_copy = %s\nregister_cache = "4GB"\nfetch_threads = 4\nverify_batch = 4\ncopy_group = 1\narena_preregister = ["%s"]\n'
                                      % ("true" if a.zero_copy else "false", d))
            with F.CurvineFileSystem(conf) as fs:
                fs.load_namespace("\n".join(mans))
                fs.preregister()
                fs.wait_registered()
                scratch = torch.empty(a.size, dtype=torch.uint8, device="cuda")
                npages = a.size // 4096
                pages = torch.empty(npages * 4096, dtype=torch.uint8, device="cuda")
                offs = [i * 4096 for i in range(npages)]
                c_offs = (ctypes.c_uint64 * npages)(*offs)
                stream = torch.cuda.current_stream().cuda_stream
                # ---- one C-ABI call per file: open -> fuse read into page buffers -> verify -> close
                from curvine_b200 import kernels as K
                for rep in range(2):
                    lat = []
                    launches0 = K.launch_count()
                    t0 = time.perf_counter()
                    for i in order:
                        t1 = time.perf_counter()
                        got, bad = fs.fuse_read_file_device("/small/f%d" % i, a.size, scratch.data_ptr(), pages.data_ptr(), c_offs, 4096, stream)
                        assert got == a.size and bad == 0
                        lat.append(time.perf_counter() - t1)
                    dt = time.perf_counter() - t0
                    lat = np.array(lat) * 1e6
                    res["gpu_one_call_rep%d" % rep] = {"files_per_s": a.files / dt, "GBps": a.files * a.size / dt / 1e9, "p50_us": float(np.percentile(lat, 50)),
                                                       "p90_us": float(np.percentile(lat, 90)), "p99_us": float(np.percentile(lat, 99)),
                                                       "kernel_launches": int(K.launch_count() - launches0)}
                for rep in range(2):  # rep 0 warms (registers mappings); rep 1 is reported
                    lat = []
                    t0 = time.perf_counter()
                    for i in order:
                        t1 = time.perf_counter()
                        r = fs.open("/small/f%d" % i)
                        got = r.fuse_read_device(0, a.size, scratch.data_ptr(), pages.data_ptr(), offs, 4096, stream)
                        s, bad, ver = r.verify()
                        r.complete()
                        torch.cuda.current_stream().synchronize()
                        assert got == a.size and bad == 0 and ver == 1
                        lat.append(time.perf_counter() - t1)
                    dt = time.perf_counter() - t0
                    lat = np.array(lat) * 1e6
                    res["gpu_rep%d" % rep] = {"files_per_s": a.files / dt, "GBps": a.files * a.size / dt / 1e9, "p50_us": float(np.percentile(lat, 50)),
                                              "p99_us": float(np.percentile(lat, 99))}
                # batched: all files in ONE pipelined call (cv_read_many_device), files land back to back in HBM
                big = torch.empty(a.files * a.size, dtype=torch.uint8, device="cuda")
                paths = ["/small/f%d" % i for i in order]
                doffs = [k * a.size for k in range(a.files)]
                for rep in range(3):
                    t0 = time.perf_counter()
                    total, s, bad, ver = fs.read_many_device(paths, big.data_ptr(), doffs, a.files * a.size, stream)
                    torch.cuda.current_stream().synchronize()
                    dt = time.perf_counter() - t0
                    assert bad == 0 and ver == a.files and total == a.files * a.size
                    res["gpu_batched_rep%d" % rep] = {"files_per_s": a.files / dt, "GBps": total / dt / 1e9, "ms_total": dt * 1e3}
            # the CPU reference port reads the same files from a reference-layout store (one tmpfs file per block) served by the
            # oracle's worker emulator: nothing of the product is on the baseline's path
            from oracle import refworker
            with refworker.RefWorker(d + "_ref") as rw:
                for i in range(a.files):
                    rw.create_file(100000 + i, a.size, a.size, threads=1)
                for i in order[:256]:  # warm-up
                    clib.cpu_read_file(rw.port, True, a.size, a.size, [layout.create_block_id(100000 + int(i), 0)], 131072, 8, 1, 131072, 0, 1)
                lat = []
                t0 = time.perf_counter()
                for i in order:
                    t1 = time.perf_counter()
                    clib.cpu_read_file(rw.port, True, a.size, a.size, [layout.create_block_id(100000 + int(i), 0)], 131072, 8, 1, 131072, 0, 1)
                    lat.append(time.perf_counter() - t1)
                dt = time.perf_counter() - t0
            shutil.rmtree(d + "_ref", ignore_errors=True)
            lat = np.array(lat) * 1e6
            res["cpu_reference_port"] = {"files_per_s": a.files / dt, "GBps": a.files * a.size / dt / 1e9, "p50_us": float(np.percentile(lat, 50)),
                                         "p99_us": float(np.percentile(lat, 99)), "note": "open+read+crc32 into host memory, 2 threads (1 sub-reader + caller)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    if a.emit is None:
        print(json.dumps(res, indent=1))
        return
    one, cpu, bat = res["gpu_one_call_rep1"], res["cpu_reference_port"], res["gpu_batched_rep2"]
    a.emit({"metric": "small-file random read into HBM pages (CRC-verified): files/s, one FUSE-shaped call per file", "value": one["files_per_s"], "unit": "files/s",
            "n_gpus": 1, "steps": a.files, "warmup": a.files, "ms_per_step": 1e3 / one["files_per_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C5: %d x %d KiB single-block files, mem tier (%s), Fisher-Yates order (seed 7), open -> fuse_read(0, size) into 4 KiB page buffers -> verify -> close"
                                   % (a.files, a.size >> 10, "arena" if a.arena else "files"), "files": a.files, "file_bytes": a.size},
            "per_file": one, "per_file_four_calls": res["gpu_rep1"], "batched_read_many": bat,
            "cpu_baseline": {"value": cpu["files_per_s"], "unit": "files/s", "cores": 2, "kind": "port", "p50_us": cpu["p50_us"], "p99_us": cpu["p99_us"],
                             "sample": "all %d files once: open + read + crc32 into host memory (oracle/cpu_reader.c)" % a.files},
            # every per-file call IS end to end: bytes start in the worker's host memory, the timed call includes the H2D of the file, the
            # page scatter, the CRC kernel and the D2H of the CRC and the mismatch count
            "e2e": {"value": one["files_per_s"], "unit": "files/s", "h2d_bytes_per_step": a.size, "d2h_bytes_per_step": 8,
                    "what": "cv_fuse_read_file_device per file: open -> device read into 4 KiB pages -> verify -> close"},
            "gpu_launches": one["kernel_launches"], "raw": res})


if __name__ == "__main__":
    main()
