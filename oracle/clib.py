"""ctypes handle on oracle/liboracle.so (C restatement).  Test infrastructure only."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".c")]
    if force or not os.path.exists(PATH) or any(os.path.getmtime(s) > os.path.getmtime(PATH) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(PATH)
        vp, sz, u32, u64, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
        L.cvo_crc.argtypes, L.cvo_crc.restype = [i, u32, vp, sz], u32
        L.cvo_crc_bitwise.argtypes, L.cvo_crc_bitwise.restype = [i, u32, vp, sz], u32
        L.cvo_crc_blocks.argtypes, L.cvo_crc_blocks.restype = [i, vp, sz, sz, vp], None
        L.cvo_bench_checksum.argtypes, L.cvo_bench_checksum.restype = [vp, sz, sz, i], u64
        L.cvo_synth_block.argtypes, L.cvo_synth_block.restype = [u64, u64, vp, sz], None
        L.cvo_crc32_pclmul.argtypes, L.cvo_crc32_pclmul.restype = [u32, vp, sz], u32
        i64 = ctypes.c_int64
        L.cvo_cpu_read_file.argtypes = [ctypes.c_char_p, i, i, i64, i64, vp, i64, i, i, i64, i64, i, vp]
        L.cvo_cpu_read_file.restype = i
        _lib = L
    return _lib


def _buf(b):
    import numpy as np
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return a, a.ctypes.data


def crc(poly_id: int, data, init: int = 0) -> int:
    a, p = _buf(data)
    return lib().cvo_crc(poly_id, init, p, a.size)


def crc_blocks(poly_id: int, data, block_size: int):
    import numpy as np
    a, p = _buf(data)
    n = (a.size + block_size - 1) // block_size
    out = np.zeros(n, dtype=np.uint32)
    lib().cvo_crc_blocks(poly_id, p, a.size, block_size, out.ctypes.data)
    return out


def bench_checksum(data, buf_size: int, stale_tail: bool = True) -> int:
    a, p = _buf(data)
    return lib().cvo_bench_checksum(p, a.size, buf_size, 1 if stale_tail else 0)


def synth_block(file_id: int, block_index: int, length: int):
    import numpy as np
    out = np.empty(length, dtype=np.uint8)
    lib().cvo_synth_block(file_id, block_index, out.ctypes.data, length)
    return out


def crc32_pclmul(data, init: int = 0) -> int:
    a, p = _buf(data)
    return lib().cvo_crc32_pclmul(init, p, a.size)


def cpu_read_file(port: int, short_circuit: bool, file_len: int, block_size: int, block_ids, chunk_size: int = 131072,
                  chunk_num: int = 8, read_parallel: int = 1, buf_size: int = 131072, limit: int = 0, checksum: int = 1,
                  ip: str = "127.0.0.1"):
    """The reference-shaped CPU reader (oracle/cpu_reader.c).  -> (bytes, sum_crc32, threads)."""
    import numpy as np
    ids = np.asarray(block_ids, dtype=np.int64)
    out = np.zeros(3, dtype=np.uint64)
    rc = lib().cvo_cpu_read_file(ip.encode(), port, 1 if short_circuit else 0, file_len, block_size, ids.ctypes.data, chunk_size,
                                 chunk_num, read_parallel, buf_size, limit, checksum, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("cvo_cpu_read_file failed")
    return int(out[0]), int(out[1]), int(out[2])


def reference_read_parallel(file_len: int, read_parallel: int = 1, large_file_size: int = 10 << 30, max_read_parallel: int = 8) -> int:
    """ReadDetector::with_conf (read_detector.rs:130-135) with the reference defaults."""
    if file_len >= large_file_size:
        return min(max_read_parallel, max(1, (file_len + large_file_size - 1) // large_file_size))
    return read_parallel
