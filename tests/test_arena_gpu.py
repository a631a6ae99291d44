"""Mem-arena tier through the GPU reader (C ABI): blocks are DMA'd straight out of arena segments pinned once per
context; the first read of a file -- even one written after the segments were pinned -- needs no per-file state.
Bytes in HBM must equal the oracle generator's, CRCs the oracle's (same bar as tests/test_gpu_reader.py)."""
import os
import shutil
import tempfile

import numpy as np
import pytest

from curvine_b200 import _lib, fs as F
from oracle import clib, synth

pytestmark = pytest.mark.gpu

ARENA = 'mem_arena = true\narena_segment = "8MB"\narena_reuse_delay = "0ms"\n'


@pytest.fixture()
def aw():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="cvagpu", dir=base)
    w = F.MiniWorker(["[MEM:32MB]" + d + "/m0", "[MEM:32MB]" + d + "/m1"], extra_worker=ARENA)
    yield w, d
    w.stop()
    shutil.rmtree(d, ignore_errors=True)


def _conf(d, sc=True, prereg=True, extra="", threads=4, copy_group=4):
    pre = 'arena_preregister = ["%s/m0", "[MEM]%s/m1"]\n' % (d, d) if prereg else ""
    return F.client_conf(short_circuit=sc, b200='fetch_threads = %d\nverify_batch = 4\npinned_slots = 12\nzero_copy = true\ncopy_group = %d\n'
                         'register_threads = 2\narena_register_slice = "2MB"\ngpu_chunk_size = "256KB"\n%s%s' % (threads, copy_group, pre, extra))


def _buf(n, cuda):
    import torch
    return torch.full((n,), 0x5A, dtype=torch.uint8, device=cuda)


def _sum(poly, data, bs):
    return int(clib.crc_blocks(poly, np.frombuffer(data, dtype=np.uint8), bs).astype(np.uint64).sum())


def test_first_read_of_a_file_is_dma_from_the_pinned_arena(cuda, aw):
    import torch
    w, d = aw
    n, bs, ino = (13 << 20) + 4097, 1 << 20, 6101
    man = w.create_file("/a", ino, n, bs)
    want = synth.file_bytes(ino, n, bs)
    with F.CurvineFileSystem(_conf(d)) as fs:
        fs.load_namespace(man)
        fs.preregister()
        fs.wait_registered()
        st = fs.arena_stats()
        assert st["segments"] == 8 and st["pinned_bytes"] == 64 << 20 and st["dma_jobs"] == 0
        r = fs.open("/a")
        dst = _buf(n + 32, cuda)
        assert r.read_device(dst.data_ptr(), n + 32, torch.cuda.current_stream().cuda_stream) == n
        s, bad, ver = r.verify()
        torch.cuda.synchronize()
        assert bad == 0 and ver == 14 and s == _sum(1, want, bs)
        assert dst[:n].cpu().numpy().tobytes() == want and (dst[n:] == 0x5A).all()
        ds = r.device_stats()
        assert ds["ring_alloc_sec"] == 0 and ds["reg_misses"] == 0  # no pinned ring, no per-file mapping cache involved
        r.complete()
        assert fs.arena_stats()["dma_jobs"] == 14 and fs.arena_stats()["dma_bytes"] == n


def test_a_file_written_after_the_arena_was_pinned_streams_without_any_new_registration(cuda, aw):
    import torch
    w, d = aw
    with F.CurvineFileSystem(_conf(d)) as fs:
        fs.preregister()
        fs.wait_registered()
        pinned = fs.arena_stats()
        for k in range(3):  # files that did not exist when the segments were pinned; the third reuses freed extents
            ino, n, bs = 6200 + k, (6 << 20) + 100 * k, 2 << 20
            man = w.create_file("/late%d" % k, ino, n, bs, threads=1)
            fs.load_namespace(man)
            r = fs.open("/late%d" % k)
            dst = _buf(n, cuda)
            assert r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == n
            s, bad, ver = r.verify()
            torch.cuda.synchronize()
            want = synth.file_bytes(ino, n, bs)
            assert bad == 0 and s == _sum(1, want, bs) and dst.cpu().numpy().tobytes() == want
            r.complete()
            if k == 1:
                w.delete_file(6200, 3)
        now = fs.arena_stats()
        assert now["segments"] == pinned["segments"] and now["pinned_bytes"] == pinned["pinned_bytes"]
        assert now["dma_jobs"] == 11  # 3 + 4 + 4 blocks


def test_segments_are_pinned_on_demand_without_preregistration(cuda, aw):
    import torch
    w, d = aw
    n, bs, ino = 5 << 20, 1 << 20, 6301
    man = w.create_file("/od", ino, n, bs)
    with F.CurvineFileSystem(_conf(d, prereg=False)) as fs:
        fs.load_namespace(man)
        r = fs.open("/od")
        dst = _buf(n, cuda)
        assert r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == n
        s, bad, ver = r.verify()
        torch.cuda.synchronize()
        assert bad == 0 and dst.cpu().numpy().tobytes() == synth.file_bytes(ino, n, bs)
        r.complete()
        st = fs.arena_stats()
        assert 1 <= st["segments"] <= 2 and st["dma_jobs"] == 5


def test_sharded_read_with_one_arena_per_rank(cuda, aw):
    import torch
    w, d = aw
    n, bs, ino = (9 << 20) + 17, 1 << 20, 6401
    _lib.lib().cv_synth_set_shard_world(2)
    try:
        man = w.create_file("/sh", ino, n, bs, threads=3)
    finally:
        _lib.lib().cv_synth_set_shard_world(0)
    want = synth.file_bytes(ino, n, bs)
    total = 0
    with F.CurvineFileSystem(_conf(d)) as fs:
        fs.load_namespace(man)
        for rank in range(2):
            r = fs.open("/sh")
            plan = r.shard_plan(rank, 2)
            cap = len(plan) * bs
            dst = _buf(cap, cuda)
            got = r.read_device_sharded(rank, 2, dst.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
            s, bad, ver = r.verify()
            torch.cuda.synchronize()
            assert bad == 0 and ver == len(plan)
            host = dst.cpu().numpy().tobytes()
            for (b, foff, ln, doff) in plan:
                assert host[doff:doff + ln] == want[foff:foff + ln]
            total += s
            r.complete()
    assert total == _sum(1, want, bs)


@pytest.mark.parametrize("chunk", ["64KB", "1MB"])
def test_framed_read_is_served_out_of_the_arena_mapping(cuda, aw, chunk):
    import torch
    w, d = aw
    n, bs, ino = (7 << 20) + 333, 1 << 20, 6501
    man = w.create_file("/fr", ino, n, bs)
    want = synth.file_bytes(ino, n, bs)
    conf = _conf(d, sc=False, prereg=False, extra='')
    conf = conf.replace('gpu_chunk_size = "256KB"', 'gpu_chunk_size = "%s"' % chunk)
    with F.CurvineFileSystem(conf) as fs:
        fs.load_namespace(man)
        r = fs.open("/fr")
        r.seek(12345)
        dst = _buf(n, cuda)
        got = r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        s, bad, ver = r.verify()
        torch.cuda.synchronize()
        assert got == n - 12345 and bad == 0 and ver == 7  # block 0 is partial: not comparable
        assert dst[:got].cpu().numpy().tobytes() == want[12345:]
        r.complete()
    assert w.metrics()["read_blocks_remote"] >= 8


def test_small_files_batch_and_fuse_scatter_over_the_arena(cuda, aw):
    import torch
    w, d = aw
    paths, want, offs, off = [], [], [], 0
    with F.CurvineFileSystem(_conf(d)) as fs:
        for i in range(40):
            n = 200000 + 4099 * i
            fs.load_namespace(w.create_file("/sf%d" % i, 6600 + i, n, 256 << 10, threads=1))
            paths.append("/sf%d" % i)
            want.append(synth.file_bytes(6600 + i, n, 256 << 10))
            offs.append(off)
            off += (n + 255) // 256 * 256
        dst = _buf(off, cuda)
        tot, s, bad, ver = fs.read_many_device(paths, dst.data_ptr(), offs, off, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert tot == sum(len(x) for x in want) and bad == 0
        host = dst.cpu().numpy().tobytes()
        for o, x in zip(offs, want):
            assert host[o:o + len(x)] == x
        # FUSE-shaped: 200,000 bytes of /sf0 scattered into 4 KiB pages in reverse page order
        r = fs.open("/sf0")
        npages = (200000 + 4095) // 4096
        pages = _buf(npages * 4096, cuda)
        scratch = _buf(200000, cuda)
        page_offs = [(npages - 1 - i) * 4096 for i in range(npages)]
        got = r.fuse_read_device(0, 200000, scratch.data_ptr(), pages.data_ptr(), page_offs, 4096, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert got == 200000
        ph = pages.cpu().numpy().tobytes()
        for i in range(npages):
            ln = min(4096, 200000 - i * 4096)
            assert ph[page_offs[i]:page_offs[i] + ln] == want[0][i * 4096:i * 4096 + ln]
        r.complete()
        # the one-call variant (open -> fuse read -> verify -> close) a FUSE daemon would use per small file
        import ctypes
        arr = (ctypes.c_uint64 * npages)(*page_offs)
        for i in (3, 17, 39):
            pages.fill_(0)
            ln = len(want[i])
            np_i = (ln + 4095) // 4096
            big_pages = _buf(np_i * 4096, cuda)
            offs_i = (ctypes.c_uint64 * np_i)(*[k * 4096 for k in range(np_i)])
            sc = _buf(ln, cuda)
            got, bad = fs.fuse_read_file_device("/sf%d" % i, ln, sc.data_ptr(), big_pages.data_ptr(), offs_i, 4096, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert got == ln and bad == 0
            assert big_pages.cpu().numpy().tobytes()[:ln] == want[i]
        del arr


def test_corruption_inside_the_arena_is_reported_by_the_gpu_verify(cuda, aw):
    import torch
    w, d = aw
    n, bs, ino = 4 << 20, 1 << 20, 6701
    man = w.create_file("/bad", ino, n, bs, threads=1)
    from oracle import layout
    stub = layout.block_path(d + "/m0/curvine", layout.create_block_id(ino, 2))  # two MEM dirs, round robin: block 2 is in m0
    magic, seg, off, ln = open(stub).read().split()
    assert magic == "CVARENA1" and int(ln) == bs
    with open(d + "/m0/curvine/arena/seg_%04d" % int(seg), "r+b") as f:  # flip one byte of block 2 inside its extent
        f.seek(int(off) + 4242)
        b = f.read(1)
        f.seek(int(off) + 4242)
        f.write(bytes([b[0] ^ 0x40]))
    with F.CurvineFileSystem(_conf(d)) as fs:
        fs.load_namespace(man)
        r = fs.open("/bad")
        dst = _buf(n, cuda)
        r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        s, bad, ver = r.verify()
        assert bad == 1 and ver == 4
        r.complete()


def test_blocks_without_a_manifest_crc_do_not_switch_the_comparison_off(cuda, aw):
    """one block without a CRC (or a hole) inside the range used to disable the whole comparison."""
    import torch
    w, d = aw
    n, bs, ino = 6 << 20, 1 << 20, 6801
    man = w.create_file("/nocrc", ino, n, bs, threads=1)
    lines = man.splitlines()
    blocks = [i for i, l in enumerate(lines) if l.startswith("block ")]
    f = lines[blocks[2]].split()
    f[4], f[5] = "-", "-"  # block 2 loses its manifest CRCs
    lines[blocks[2]] = " ".join(f)
    f = lines[blocks[4]].split()
    f[5] = "%08x" % (int(f[5], 16) ^ 1)  # block 4's expected CRC-32C is wrong
    lines[blocks[4]] = " ".join(f)
    with F.CurvineFileSystem(_conf(d)) as fs:
        fs.load_namespace("\n".join(lines) + "\n")
        r = fs.open("/nocrc")
        dst = _buf(n, cuda)
        r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        s, bad, ver = r.verify()
        assert ver == 5 and bad == 1
        r.complete()
