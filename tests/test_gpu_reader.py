"""End-to-end GPU reader through the C ABI vs the oracle: bytes in HBM bit-exact, on-GPU CRC verify, sharding."""
import os

import numpy as np
import pytest

from curvine_b200 import fs as F
from oracle import clib, layout, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cluster(tmp_path_factory):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tmp_path_factory.mktemp("gw") if base is None else __import__("pathlib").Path(__import__("tempfile").mkdtemp(prefix="cvgpu", dir=base))
    w = F.MiniWorker(["[MEM]" + str(d / "mem")])
    yield w, d
    w.stop()
    __import__("shutil").rmtree(str(d), ignore_errors=True)


def _dev_buf(n, cuda):
    import torch
    return torch.full((n,), 0xA5, dtype=torch.uint8, device=cuda)


def _conf(sc, poly=1, chunk="4MB", threads=4, batch=4, zero_copy=False, copy_group=2, register_threads=0, register_when_idle=True,
          register_cache="48MB"):
    return F.client_conf(short_circuit=sc, b200='verify_poly = %d\ngpu_chunk_size = "%s"\nfetch_threads = %d\nverify_batch = %d\npinned_slots = 12\n'
                         'zero_copy = %s\ncopy_group = %d\nregister_cache = "%s"\nregister_threads = %d\nregister_when_idle = %s\n'
                         % (poly, chunk, threads, batch, "true" if zero_copy else "false", copy_group, register_cache, register_threads,
                            "true" if register_when_idle else "false"))


@pytest.mark.parametrize("sc,chunk", [(True, "4MB"), (False, "128KB"), (False, "1MB"), (False, "4MB")])
@pytest.mark.parametrize("poly", [0, 1])
def test_c1_file_lands_bit_exact_and_verifies(cuda, cluster, sc, chunk, poly):
    """C1 shape: 64 MiB, 1 MiB blocks.  Bytes == oracle generator; per-block CRC == manifest == oracle."""
    import torch
    w, _ = cluster
    n, bs, ino = (64 << 20) - 4321, 1 << 20, 7001
    man = w.create_file("/c1", ino, n, bs)
    want = synth.file_bytes(ino, n, bs)
    with F.CurvineFileSystem(_conf(sc, poly, chunk)) as fs:
        fs.load_namespace(man)
        r = fs.open("/c1")
        dst = _dev_buf(n + 64, cuda)
        got = r.read_device(dst.data_ptr(), n + 64, torch.cuda.current_stream().cuda_stream)
        assert got == n and r.pos() == n
        torch.cuda.synchronize()
        assert dst[:n].cpu().numpy().tobytes() == want
        assert (dst[n:] == 0xA5).all()
        s, bad, ver = r.verify()
        nb = (n + bs - 1) // bs
        assert bad == 0 and ver == nb
        assert s == int(clib.crc_blocks(poly, np.frombuffer(want, dtype=np.uint8), bs).astype(np.uint64).sum())
        st = r.device_stats()
        assert st["bytes"] == n and st["blocks"] == nb and st["kernel_launches"] > 0
        assert st["h2d_bytes"] >= n
        assert r.read_device(dst.data_ptr(), 10) == 0  # EOF is not an error
        r.complete()
    m = w.metrics()
    assert (m["read_blocks_local"] if sc else m["read_blocks_remote"]) >= nb


@pytest.mark.parametrize("sc", [True, False])
def test_partial_ranges_and_seeks(cuda, cluster, sc):
    import torch
    w, _ = cluster
    n, bs, ino = (9 << 20) + 777, 1 << 20, 7002
    man = w.create_file("/p1", ino, n, bs)
    want = synth.file_bytes(ino, n, bs)
    with F.CurvineFileSystem(_conf(sc, 0, "128KB")) as fs:
        fs.load_namespace(man)
        r = fs.open("/p1")
        for pos, cap in [(0, 100), (bs - 1, 2), (bs + 12345, 3 * bs + 17), (n - 5, 100), (3, 5 * bs)]:
            r.seek(pos)
            dst = _dev_buf(cap + 8, cuda)
            got = r.read_device(dst.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            exp = want[pos:pos + cap]
            assert got == len(exp) and r.pos() == pos + got
            assert dst[:got].cpu().numpy().tobytes() == exp and (dst[got:] == 0xA5).all()
        # host and device reads share one position
        r.seek(2 * bs + 5)
        assert r.read(10) == want[2 * bs + 5:2 * bs + 15]
        dst = _dev_buf(64, cuda)
        assert r.read_device(dst.data_ptr(), 64, 0) == 64
        torch.cuda.synchronize()
        assert dst.cpu().numpy().tobytes() == want[2 * bs + 15:2 * bs + 79]
        assert r.read(5) == want[2 * bs + 79:2 * bs + 84]
        assert r.verify()[1] == 0
        r.complete()


@pytest.mark.parametrize("sc", [True, False])
def test_corruption_is_caught_on_the_gpu(cuda, cluster, sc):
    """A flipped byte in one block file after the manifest was written -> exactly one bad block."""
    import torch
    w, d = cluster
    n, bs, ino = 8 << 20, 1 << 20, 7003 + int(sc)
    man = w.create_file("/bad%d" % sc, ino, n, bs)
    path = layout.block_path(str(d / "mem" / "curvine"), layout.create_block_id(ino, 5))
    with open(path, "r+b") as f:
        f.seek(123456)
        b = f.read(1)
        f.seek(123456)
        f.write(bytes([b[0] ^ 0x10]))
    with F.CurvineFileSystem(_conf(sc)) as fs:
        fs.load_namespace(man)
        r = fs.open("/bad%d" % sc)
        dst = _dev_buf(n, cuda)
        assert r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == n
        s, bad, ver = r.verify()
        assert bad == 1 and ver == 8
        r.complete()


def test_holes_on_device(cuda, cluster):
    import torch
    w, _ = cluster
    n, bs, ino = (6 << 20) + 9, 1 << 20, 7010
    man = w.create_file("/dh", ino, n, bs, mode=2, hole_every=3)
    want = bytearray(synth.file_bytes(ino, n, bs))
    for i in (2, 5):
        want[i * bs:(i + 1) * bs] = bytes(min(bs, n - i * bs))
    with F.CurvineFileSystem(_conf(True)) as fs:
        fs.load_namespace(man)
        r = fs.open("/dh")
        dst = _dev_buf(n, cuda)
        assert r.read_device(dst.data_ptr(), n, 0) == n
        torch.cuda.synchronize()
        assert dst.cpu().numpy().tobytes() == bytes(want)
        assert r.verify()[1] == 0
        r.complete()


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("sc", [True, False])
def test_sharded_read_plus_deinterleave_restores_file_order(cuda, cluster, world, sc):
    """Block b -> rank b % world (fs_reader_parallel.rs:112-122 analogue); gather + K-deinterleave == file."""
    import torch
    from curvine_b200 import kernels as K
    w, _ = cluster
    bs, nb, ino = 1 << 20, 19, 7020
    n = bs * nb - 1000
    man = w.create_file("/sh", ino, n, bs)
    want = synth.file_bytes(ino, n, bs)
    per = (nb + world - 1) // world
    gathered = torch.zeros(world * per * bs, dtype=torch.uint8, device=cuda)
    total_sum, total_ver = 0, 0
    with F.CurvineFileSystem(_conf(sc, 1, "1MB")) as fs:
        fs.load_namespace(man)
        for rank in range(world):
            r = fs.open("/sh")
            shard = gathered[rank * per * bs:(rank + 1) * per * bs]
            got = r.read_device_sharded(rank, world, shard.data_ptr(), per * bs, 0)
            assert got == sum(min(bs, n - b * bs) for b in range(rank, nb, world))
            s, bad, ver = r.verify()
            assert bad == 0 and ver == len(range(rank, nb, world))
            total_sum += s
            total_ver += ver
            r.complete()
    assert total_ver == nb
    assert total_sum == int(clib.crc_blocks(1, np.frombuffer(want, dtype=np.uint8), bs).astype(np.uint64).sum())
    out = torch.zeros(n, dtype=torch.uint8, device=cuda)
    K.deinterleave_blocks(gathered, per * bs, world, bs, nb, n, out)
    torch.cuda.synchronize()
    assert out.cpu().numpy().tobytes() == want


def test_fuse_shaped_device_read_scatters_into_pages(cuda, cluster):
    """C5 shape: 256 KiB single-block files; reply scattered into 4 KiB page buffers (fuse_response.rs:49-60 analogue)."""
    import torch
    w, _ = cluster
    n, ino = 256 * 1024, 7030
    man = w.create_file("/small", ino, n, n)
    want = synth.file_bytes(ino, n, n)
    with F.CurvineFileSystem(_conf(True)) as fs:
        fs.load_namespace(man)
        r = fs.open("/small")
        scratch = _dev_buf(n, cuda)
        pages = torch.zeros(128 * 4096 + 64, dtype=torch.uint8, device=cuda)
        rng = np.random.default_rng(1)
        order = rng.permutation(128)[:60]  # 60 pages, scattered, not in order
        offs = [int(p) * 4096 + 7 for p in order]  # deliberately unaligned page buffers
        got = r.fuse_read_device(4096, 60 * 4096 - 100, scratch.data_ptr(), pages.data_ptr(), offs, 4096, 0)
        torch.cuda.synchronize()
        assert got == 60 * 4096 - 100 and r.pos() == 4096 + got
        host = pages.cpu().numpy()
        for i, o in enumerate(offs):
            seg = want[4096 + i * 4096:4096 + min((i + 1) * 4096, got)]
            assert host[o:o + len(seg)].tobytes() == seg
        r.complete()


@pytest.mark.parametrize("copy_group,register_threads,when_idle", [(1, 0, True), (4, 0, True), (4, 2, True), (2, 2, False)])
def test_zero_copy_registered_mappings(cuda, cluster, copy_group, register_threads, when_idle):
    """Mem-tier zero-copy: DMA straight from cudaHostRegister'ed mmaps of the block files.  Same bytes, same CRCs;
    in-place corruption and file replacement are both seen (mappings revalidated by inode/size/mtime); the LRU
    (48 MB here, file 64 MiB) evicts without breaking anything."""
    import torch
    w, d = cluster
    n, bs, ino = (64 << 20) - 4096 * 3 - 5, 1 << 20, 7040 + copy_group + 10 * register_threads
    idle_bg = register_threads > 0 and when_idle  # registrar yields to reads in flight: the first pass is ring-only
    man = w.create_file("/zc%d" % copy_group, ino, n, bs)
    want = bytearray(synth.file_bytes(ino, n, bs))
    nb = (n + bs - 1) // bs
    # eager background registration races the reader for the LRU: with a cache smaller than the file, a sequential re-read can
    # find every group evicted just before it gets there (0 hits, by timing), so that case gets a cache that holds the file;
    # the other three keep the 48 MB cache and cover eviction
    cache = "48MB" if (when_idle or register_threads == 0) else "128MB"
    with F.CurvineFileSystem(_conf(True, 1, zero_copy=True, copy_group=copy_group, register_threads=register_threads,
                                   register_when_idle=when_idle, register_cache=cache)) as fs:
        fs.load_namespace(man)
        for rep in range(3):
            fs.wait_registered()  # background mode: rep 0 goes through the ring while the registrar maps the files
            r = fs.open("/zc%d" % copy_group)
            dst = _dev_buf(n + 16, cuda)
            assert r.read_device(dst.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == n
            s, bad, ver = r.verify()
            torch.cuda.synchronize()
            assert bad == 0 and ver == nb
            assert dst[:n].cpu().numpy().tobytes() == bytes(want) and (dst[n:] == 0xA5).all()
            st = r.device_stats()
            assert st["reg_misses"] > 0, "the registered-mapping path did not run (silent fallback to the pinned ring)"
            if rep == 0 and idle_bg:
                assert st["reg_hits"] == 0, "idle-priority registrar ran during the cold pass"
            if rep == 2:
                assert st["reg_hits"] > 0, "no read was served from a registered mapping"
            r.complete()
        # partial range through the mapped path
        r = fs.open("/zc%d" % copy_group)
        r.seek(3 * bs + 17)
        dst = _dev_buf(2 * bs + 100, cuda)
        assert r.read_device(dst.data_ptr(), 2 * bs + 100, 0) == 2 * bs + 100
        torch.cuda.synchronize()
        assert dst.cpu().numpy().tobytes() == bytes(want[3 * bs + 17:5 * bs + 117])
        r.complete()
        # corrupt block 7 in place, replace block 9 by a new file with different content
        p7 = layout.block_path(str(d / "mem" / "curvine"), layout.create_block_id(ino, 7))
        with open(p7, "r+b") as f:
            f.seek(4096)
            f.write(b"\x00\x01\x02")
        p9 = layout.block_path(str(d / "mem" / "curvine"), layout.create_block_id(ino, 9))
        blk9 = bytes(bs)
        os.remove(p9)
        with open(p9, "wb") as f:
            f.write(blk9)
        want[7 * bs + 4096:7 * bs + 4099] = b"\x00\x01\x02"
        want[9 * bs:10 * bs] = blk9
        r = fs.open("/zc%d" % copy_group)
        dst = _dev_buf(n, cuda)
        assert r.read_device(dst.data_ptr(), n, 0) == n
        s, bad, ver = r.verify()
        torch.cuda.synchronize()
        assert bad == 2 and dst.cpu().numpy().tobytes() == bytes(want)
        r.complete()


@pytest.mark.parametrize("zero_copy", [False, True])
def test_read_many_small_files_in_one_pass(cuda, cluster, zero_copy):
    """C5 batching: many single-block files in one pipelined call; bytes, CRC sum and verify count match the oracle."""
    import torch
    w, _ = cluster
    size, nfiles = 256 * 1024, 48
    mans = [w.create_file("/many/f%d" % i, 7100 + i, size - (i % 3) * 4096, size, threads=1) for i in range(nfiles)]
    order = [int(x) for x in np.random.default_rng(7).permutation(nfiles)]
    with F.CurvineFileSystem(_conf(True, 1, zero_copy=zero_copy, copy_group=1)) as fs:
        fs.load_namespace("\n".join(mans))
        dst = _dev_buf(nfiles * size, cuda)
        paths = ["/many/f%d" % i for i in order]
        offs = [k * size for k in range(nfiles)]
        total, s, bad, ver = fs.read_many_device(paths, dst.data_ptr(), offs, nfiles * size, 0)
        torch.cuda.synchronize()
        host = dst.cpu().numpy()
        exp_sum = 0
        for k, i in enumerate(order):
            n = size - (i % 3) * 4096
            want = synth.block_bytes(7100 + i, 0, n)
            assert host[k * size:k * size + n].tobytes() == want
            assert (host[k * size + n:(k + 1) * size] == 0xA5).all()
            exp_sum += clib.crc(1, want)
        assert bad == 0 and ver == nfiles and s == exp_sum and total == sum(size - (i % 3) * 4096 for i in range(nfiles))
        with pytest.raises(F.FsError):
            fs.read_many_device(["/many/nope"], dst.data_ptr(), [0], size, 0)


@pytest.mark.parametrize("sc", [True, False])
def test_edge_shapes_on_device(cuda, cluster, sc):
    """Empty file, 5-byte file, ragged last block; framed mode with the 16 MiB maximum frame (rpc_message.rs:41)."""
    import torch
    w, _ = cluster
    bs = 20 << 20
    n = bs + (16 << 20) + 3
    man = w.create_file("/dedge/empty", 7200, 0, 1 << 20) + w.create_file("/dedge/tiny", 7201, 5, 1 << 20) + w.create_file("/dedge/big", 7202, n, bs)
    with F.CurvineFileSystem(_conf(sc, 1, "16MB", threads=2, batch=2)) as fs:
        fs.load_namespace(man)
        dst = _dev_buf(n + 8, cuda)
        with fs.open("/dedge/empty") as r:
            assert r.read_device(dst.data_ptr(), 100, 0) == 0 and r.verify() == (0, 0, 0)
        with fs.open("/dedge/tiny") as r:
            assert r.read_device(dst.data_ptr(), 100, 0) == 5
            s, bad, ver = r.verify()
            torch.cuda.synchronize()
            assert bad == 0 and ver == 1 and dst[:5].cpu().numpy().tobytes() == synth.block_bytes(7201, 0, 5) and s == clib.crc(1, synth.block_bytes(7201, 0, 5))
        with fs.open("/dedge/big") as r:  # 20 MiB block = a 16 MiB frame + a 4 MiB frame when framed
            assert r.read_device(dst.data_ptr(), n, 0) == n
            s, bad, ver = r.verify()
            torch.cuda.synchronize()
            assert bad == 0 and ver == 2 and dst[:n].cpu().numpy().tobytes() == synth.file_bytes(7202, n, bs)


def test_write_device_then_read_device_roundtrip(cuda, cluster):
    """Write-side mirror on the GPU: HBM bytes -> K4 (prefix write + copy + CRC at source) -> wire image -> worker;
    then read back into HBM through K2/K1.  Bytes identical, write-time CRCs in the manifest == read-time CRCs."""
    import torch
    w, d = cluster
    bs, n, ino = 1 << 20, (5 << 20) + 12345, 7300
    g = torch.Generator(device=cuda).manual_seed(11)
    src = torch.randint(0, 256, (n,), dtype=torch.uint8, device=cuda, generator=g)
    host = src.cpu().numpy()
    for sc in (True, False):
        with F.CurvineFileSystem(_conf(sc, 1, "128KB")) as fs:
            wr = fs.create("/wd/f%d" % sc, ino + int(sc), bs, w.port, chunk_size=131072)
            wr.write_device(src.data_ptr(), 3 * bs + 100, torch.cuda.current_stream().cuda_stream)  # ends mid-block
            wr.write_device(src.data_ptr() + 3 * bs + 100, n - 3 * bs - 100, torch.cuda.current_stream().cuda_stream)
            man = wr.complete()
            blocks = [l.split() for l in man.splitlines() if l.startswith("block ")]
            assert [int(b[4], 16) for b in blocks] == clib.crc_blocks(0, host, bs).tolist()
            assert [int(b[5], 16) for b in blocks] == clib.crc_blocks(1, host, bs).tolist()
            for i in range(len(blocks)):  # the worker wrote the reference layout
                p = layout.block_path(str(d / "mem" / "curvine"), layout.create_block_id(ino + int(sc), i))
                assert open(p, "rb").read() == host[i * bs:(i + 1) * bs].tobytes()
            r = fs.open("/wd/f%d" % sc)
            dst = _dev_buf(n, cuda)
            assert r.read_device(dst.data_ptr(), n, 0) == n
            s, bad, ver = r.verify()
            torch.cuda.synchronize()
            assert bad == 0 and ver == len(blocks) and torch.equal(dst, src)
            assert s == int(clib.crc_blocks(1, host, bs).astype(np.uint64).sum())
            r.complete()


def test_hbm_tier_serves_gpu_packed_frames(cuda, cluster):
    """HBM as a worker tier: blocks made resident in device memory are served to remote readers as frames packed on the
    GPU (K4).  The block files are DELETED first, so every byte must come from HBM; host reader (reference chunking,
    seeks) and GPU reader (pipelined requests, K2 unpack + CRC) both get the oracle's bytes."""
    import torch
    w, d = cluster
    bs, n, ino = 1 << 20, (4 << 20) + 4321, 7400
    man = w.create_file("/hbm", ino, n, bs)
    want = synth.file_bytes(ino, n, bs)
    for i in range(5):
        w.hbm_load(layout.create_block_id(ino, i), 0)
        os.remove(layout.block_path(str(d / "mem" / "curvine"), layout.create_block_id(ino, i)))
    assert w.hbm_stats()["resident_blocks"] >= 5
    with F.CurvineFileSystem(_conf(False, 1, "128KB")) as fs:  # remote (framed) reads
        fs.load_namespace(man)
        r = fs.open("/hbm")
        assert r.read_full(n) == want  # host path: 128 KiB ping-pong chunks out of the packed stream
        r.seek(bs + 777)  # mid-block seek -> DataHeaderProto.offset -> re-pack from the new offset
        assert r.read_full(300000) == want[bs + 777:bs + 777 + 300000]
        dst = _dev_buf(n, cuda)
        r.seek(0)
        assert r.read_device(dst.data_ptr(), n, 0) == n
        s, bad, ver = r.verify()
        torch.cuda.synchronize()
        assert bad == 0 and ver == 5 and dst.cpu().numpy().tobytes() == want
        r.complete()
    st = w.hbm_stats()
    assert st["reads_from_hbm"] >= 10 and st["packed_bytes"] >= 2 * n


def test_read_to_tensor_binding(cuda, cluster):
    import torch
    w, _ = cluster
    n, ino = (3 << 20) + 77, 7500
    man = w.create_file("/t2t", ino, n, 1 << 20)
    with F.CurvineFileSystem(_conf(True)) as fs:
        fs.load_namespace(man)
        mock = bool(os.environ.get("CV_TEST_MOCK_CUDA_LIB"))  # host-side stand-ins: "device memory" is host memory, the binding is asked for a CPU tensor
        t = fs.read_to_tensor("/t2t", device="cpu" if mock else None)
        assert (mock or t.is_cuda) and t.dtype == torch.uint8 and t.numel() == n
        assert t.cpu().numpy().tobytes() == synth.file_bytes(ino, n, 1 << 20)
        assert torch.utils.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t)).data_ptr() == t.data_ptr()
