"""ctypes wrapper of oracle/ref_worker.c: a worker emulator over a reference-layout BlockStore (one file per block).

Test infrastructure only (see oracle/__init__.py): the reference arm of bench.py times oracle/cpu_reader.c against
this, so that nothing of the product library is on the baseline's path."""
import ctypes
import os

from . import clib


class RefWorker:
    def __init__(self, data_dir: str, cluster_id: str = "curvine"):
        L = clib.lib()
        L.cvo_ref_worker_start.argtypes, L.cvo_ref_worker_start.restype = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)], ctypes.c_void_p
        L.cvo_ref_worker_stop.argtypes, L.cvo_ref_worker_stop.restype = [ctypes.c_void_p], None
        L.cvo_ref_worker_create_file.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.cvo_ref_worker_create_file.restype = ctypes.c_int
        self._L = L
        self.base = os.path.join(data_dir, cluster_id)
        os.makedirs(self.base, exist_ok=True)
        p = ctypes.c_int()
        self._h = L.cvo_ref_worker_start(self.base.encode(), ctypes.byref(p))
        if not self._h:
            raise RuntimeError("ref worker: could not listen")
        self.port = p.value

    def create_file(self, inode_id: int, length: int, block_size: int, threads: int = 8):
        """Blocks of the oracle generator (synth.py / cvo_synth_block) as raw files blk_<inode<<24|b>."""
        if self._L.cvo_ref_worker_create_file(self._h, inode_id, length, block_size, threads) != 0:
            raise RuntimeError("ref worker: writing the block files failed")

    def stop(self):
        if self._h:
            self._L.cvo_ref_worker_stop(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
