"""Parity at a size closer to BASELINE.json's C2 than the other GPU tests reach (2 GiB of 4 MiB blocks instead of <= 256 MiB), through
size-independent properties: the u64 sum of the per-block CRCs the GPU computed equals the oracle's over the
generator's bytes, the two round-robin shards' sums add up to the file's, and sampled blocks are byte-identical.  Written after round 2's
last GPU run; sorts late on purpose."""
import os
import random
import shutil
import tempfile

import numpy as np
import pytest

from curvine_b200 import fs as F
from oracle import clib

pytestmark = pytest.mark.gpu
BLOCK = 4 << 20


def _blk(ino, b):
    """generator block through the C oracle (oracle/oracle.c; equal to oracle/synth.py by tests/test_oracle.py, 150x faster)"""
    return clib.synth_block(ino, b, BLOCK).tobytes()


@pytest.mark.parametrize("tier", ["arena", "files"])
def test_two_gib_file_crc_sums_and_sampled_blocks(cuda, tier):
    import torch
    # on the host-side stand-ins (tests/mock_cuda, tests/simt_emu: the kernels run on a fiber-per-thread shim) the same properties at 256 MiB
    n, ino = (256 << 20) if os.environ.get("CV_TEST_MOCK_CUDA_LIB") else (2 << 30), 8901
    nb = n // BLOCK
    d = tempfile.mkdtemp(prefix="cvscale", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        extra = 'mem_arena = true\narena_segment = "256MB"\n' if tier == "arena" else ""
        with F.MiniWorker([("[MEM:%d]" % (n + (256 << 20)) if tier == "arena" else "[MEM]") + d + "/m"], extra_worker=extra) as w:
            man = w.create_file("/big", ino, n, BLOCK, threads=16)
            # oracle side: per-block CRC-32C of the generator's bytes, block by block (no 2 GiB buffer on the host)
            want = np.array([clib.crc(1, clib.synth_block(ino, b, BLOCK)) for b in range(nb)], dtype=np.uint64)
            man_crc = [int(l.split()[5], 16) for l in man.splitlines() if l.startswith("block ")]
            assert man_crc == [int(x) for x in want]  # the manifest (written by the product's generator) agrees with the oracle
            conf = F.client_conf(b200='fetch_threads = 8\nverify_batch = 16\ncopy_group = 8\nzero_copy = true\narena_preregister = ["%s/m"]\n' % d)
            with F.CurvineFileSystem(conf) as fs:
                fs.load_namespace(man)
                dst = torch.empty(n, dtype=torch.uint8, device=cuda)
                r = fs.open("/big")
                assert r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == n
                s, bad, ver = r.verify()
                torch.cuda.synchronize()
                r.complete()
                assert bad == 0 and ver == nb and s == int(want.sum())
                for b in random.Random(3).sample(range(nb), 6) + [0, nb - 1]:
                    assert dst[b * BLOCK:(b + 1) * BLOCK].cpu().numpy().tobytes() == _blk(ino, b), b
                # round-robin shards: every block lands in its slot, the shard sums add up to the file's
                total = 0
                for rank in range(2):
                    r = fs.open("/big")
                    got = r.read_device_sharded(rank, 2, dst.data_ptr(), n // 2, torch.cuda.current_stream().cuda_stream)
                    s, bad, ver = r.verify()
                    torch.cuda.synchronize()
                    r.complete()
                    assert got == n // 2 and bad == 0 and ver == nb // 2 and s == int(want[rank::2].sum())
                    j = 5
                    assert dst[j * BLOCK:(j + 1) * BLOCK].cpu().numpy().tobytes() == _blk(ino, j * 2 + rank)
                    total += s
                assert total == int(want.sum())
    finally:
        shutil.rmtree(d, ignore_errors=True)
