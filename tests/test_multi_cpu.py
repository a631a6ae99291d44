"""The N>1 path on CPU: world_size-2 gloo.  Each rank reads its round-robin shard (block b -> rank b % world, the plan
cv_read_device_sharded executes) through the C ABI's host reader from one shared worker, the shards are all-gathered
(gloo stands in for NCCL) and de-interleaved with the same index map as cvk_deinterleave_blocks; the result must be
the file, and the per-rank sums of per-block CRCs must add up to the oracle's file-level figure."""
import os
import socket
import tempfile
import zlib

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, worker_port, manifest, n, bs, ino, out_dir):
    import torch
    import torch.distributed as dist
    from curvine_b200 import fs as F
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        nb = (n + bs - 1) // bs
        per = (nb + world - 1) // world
        shard = np.zeros(per * bs, dtype=np.uint8)
        crc_sum = 0
        with F.CurvineFileSystem(F.client_conf(short_circuit=(rank == 0))) as fs:  # one rank short-circuit, one framed
            fs.load_namespace(manifest)
            with fs.open("/mc") as r:
                plan = r.shard_plan(rank, world)
                assert [p[0] for p in plan] == list(range(rank, nb, world))
                for blk, file_off, ln, dst_off in plan:
                    assert file_off == blk * bs and dst_off == (blk // world) * bs
                    r.seek(file_off)
                    data = r.read_full(ln)
                    assert len(data) == ln
                    shard[dst_off:dst_off + ln] = np.frombuffer(data, dtype=np.uint8)
                    crc_sum += zlib.crc32(data)
        gathered = torch.zeros(world * per * bs, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, torch.from_numpy(shard))
        t = torch.tensor([crc_sum], dtype=torch.int64)
        dist.all_reduce(t)
        g = gathered.numpy()
        out = np.zeros(n, dtype=np.uint8)
        for b in range(nb):  # cvk_deinterleave_blocks' index map
            ln = min(bs, n - b * bs)
            src = (b % world) * per * bs + (b // world) * bs
            out[b * bs:b * bs + ln] = g[src:src + ln]
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), out)
        with open(os.path.join(out_dir, "sum%d.txt" % rank), "w") as f:
            f.write(str(int(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_read_allgather_deinterleave(tmp_path):
    import torch.multiprocessing as mp
    from curvine_b200 import fs as F
    from oracle import clib, synth
    world, bs, ino = 2, 1 << 20, 8100
    n = 7 * bs + 12345
    d = tempfile.mkdtemp(prefix="cvmc", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    with F.MiniWorker(["[MEM]" + d]) as w:
        man = w.create_file("/mc", ino, n, bs)
        mp.spawn(_rank_main, args=(world, _free_port(), w.port, man, n, bs, ino, str(tmp_path)), nprocs=world, join=True)
    want = np.frombuffer(synth.file_bytes(ino, n, bs), dtype=np.uint8)
    for rank in range(world):
        assert (np.load(tmp_path / ("rank%d.npy" % rank)) == want).all()
        assert int(open(tmp_path / ("sum%d.txt" % rank)).read()) == int(clib.crc_blocks(0, want, bs).astype(np.uint64).sum())
    __import__("shutil").rmtree(d, ignore_errors=True)
