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
