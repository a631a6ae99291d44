"""ctypes loader for libcurvine_b200.so (the C-ABI drop-in boundary).

Fails loudly: there is no CPU or PyTorch fallback for any entry point.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcurvine_b200.so")

_lib = None


class CvFrameDesc(ctypes.Structure):
    _fields_ = [("wire_off", ctypes.c_uint64), ("dst_off", ctypes.c_uint64), ("data_len", ctypes.c_uint32),
                ("header_len", ctypes.c_uint32), ("req_id", ctypes.c_int64), ("seq_id", ctypes.c_int32),
                ("block", ctypes.c_uint32), ("code", ctypes.c_uint8), ("status", ctypes.c_uint8),
                ("pad_", ctypes.c_uint8 * 6)]


class CvStreamDesc(ctypes.Structure):
    _fields_ = [("wire_off", ctypes.c_uint64), ("dst_off", ctypes.c_uint64), ("block_len", ctypes.c_uint64),
                ("req_id", ctypes.c_int64), ("chunk_size", ctypes.c_uint32), ("first_seq_id", ctypes.c_int32),
                ("block", ctypes.c_uint32), ("first_frame", ctypes.c_uint32), ("code", ctypes.c_uint8),
                ("status", ctypes.c_uint8), ("pad_", ctypes.c_uint8 * 6)]


class CvSeg(ctypes.Structure):
    _fields_ = [("src_off", ctypes.c_uint64), ("dst_off", ctypes.c_uint64), ("len", ctypes.c_uint64)]


assert ctypes.sizeof(CvFrameDesc) == 48 and ctypes.sizeof(CvStreamDesc) == 56 and ctypes.sizeof(CvSeg) == 24


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -m curvine_b200.build` (needs nvcc); "
                              "curvine_b200 has no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, u8p, u32, u64, i = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    L.cvk_init.argtypes, L.cvk_init.restype = [i], i
    L.cvk_launch_count.argtypes, L.cvk_launch_count.restype = [], u64
    L.cvk_crc_blocks.argtypes, L.cvk_crc_blocks.restype = [u8p, vp, vp, u32, i, u64, vp, vp], i
    L.cvk_verify_crcs.argtypes, L.cvk_verify_crcs.restype = [vp, vp, u32, vp, vp, vp], i
    L.cvk_unpack_frames.argtypes, L.cvk_unpack_frames.restype = [u8p, vp, u32, u32, u8p, i, u64, vp, vp, vp], i
    L.cvk_expand_streams.argtypes, L.cvk_expand_streams.restype = [vp, u32, vp, u32, vp], i
    L.cvk_gather_pages.argtypes, L.cvk_gather_pages.restype = [u8p, vp, u32, u64, u8p, vp], i
    L.cvk_pack_frames.argtypes, L.cvk_pack_frames.restype = [u8p, vp, u32, u32, u8p, i, u64, vp, vp], i
    L.cvk_deinterleave_blocks.argtypes = [u8p, u64, u32, u64, u64, u64, u8p, vp]
    L.cvk_deinterleave_blocks.restype = i
    for name, fn in list(_LATE.items()):
        fn(L)


_LATE = {}


class CudaError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        raise CudaError("%s failed: cudaError %d" % (what or "cvk call", rc))
