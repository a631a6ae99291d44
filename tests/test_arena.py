"""Mem-arena tier (curvine_b200/csrc/host/arena.h) through the host reader and the worker protocol -- no GPU.

The arena replaces the reference's one-tmpfs-file-per-block mem tier (block_meta.rs:199-237) with extents inside a few
large segment files; everything a reader sees must stay what the reference delivers: the same bytes, chunk boundaries,
positions and error behaviour, short-circuit and framed (block_test.rs:33-103 scenarios)."""
import os
import shutil
import tempfile

import pytest

from curvine_b200 import fs as F
from oracle import synth


def _mk():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    return tempfile.mkdtemp(prefix="cvarena", dir=base)


ARENA = 'mem_arena = true\narena_segment = "%s"\narena_reuse_delay = "%s"\n'


@pytest.fixture()
def arena_worker():
    d = _mk()
    w = F.MiniWorker(["[MEM:24MB]" + d + "/m0"], extra_worker=ARENA % ("8MB", "0ms"))
    yield w, d
    w.stop()
    shutil.rmtree(d, ignore_errors=True)


def _read_all(fs, path, size=1 << 16):
    r = fs.open(path)
    out = bytearray()
    while True:
        b = r.read(size)
        if not b:
            break
        out += b
    r.complete()
    return bytes(out)


@pytest.mark.parametrize("sc", [True, False])
def test_arena_file_reads_bit_exact_short_circuit_and_framed(arena_worker, sc):
    w, d = arena_worker
    n, bs, ino = (5 << 20) + 12345, 1 << 20, 5101
    man = w.create_file("/a", ino, n, bs)
    st = w.arena_stats()
    assert st["arenas"] == 1 and st["segments"] == 3 and st["segment_bytes"] == 8 << 20
    assert st["used_bytes"] == 5 * (1 << 20) + 16384  # 4 KiB granules: the 12345-byte tail block holds 16 KiB
    # the reference-layout path holds an extent descriptor, not the bytes
    from oracle import layout
    stub = layout.block_path(d + "/m0/curvine", layout.create_block_id(ino, 0))
    text = open(stub).read()
    assert text.startswith("CVARENA1 0 0 1048576")
    with F.CurvineFileSystem(F.client_conf(short_circuit=sc)) as fs:
        fs.load_namespace(man)
        assert _read_all(fs, "/a") == synth.file_bytes(ino, n, bs)
        # seeks + chunk semantics as block_test.rs:209-302
        r = fs.open("/a")
        r.seek(bs + 7)
        assert r.read(10) == synth.file_bytes(ino, n, bs)[bs + 7:bs + 17]
        assert r.pos() == bs + 17
        r.seek(n - 3)
        assert r.read_full(100) == synth.file_bytes(ino, n, bs)[n - 3:]
        r.complete()
    m = w.metrics()
    assert (m["read_blocks_local"] if sc else m["read_blocks_remote"]) >= 6


def test_client_that_is_not_arena_aware_gets_unsupported_on_short_circuit_and_works_framed(arena_worker):
    w, _ = arena_worker
    n, bs, ino = 3 << 20, 1 << 20, 5102
    man = w.create_file("/b", ino, n, bs)
    with F.CurvineFileSystem(F.client_conf(short_circuit=True, b200="arena = false\n")) as fs:
        fs.load_namespace(man)
        r = fs.open("/b")
        with pytest.raises(F.FsError) as ei:
            r.read(10)
        assert ei.value.kind == 19 and "arena" in ei.value.msg  # Unsupported
    with F.CurvineFileSystem(F.client_conf(short_circuit=False, b200="arena = false\n")) as fs:
        fs.load_namespace(man)
        assert _read_all(fs, "/b") == synth.file_bytes(ino, n, bs)


def test_extents_never_straddle_a_segment_and_free_space_is_reused(arena_worker):
    w, _ = arena_worker
    bs = 3 << 20  # 2 blocks of 3 MiB per 8 MiB segment, 2 MiB tail left over per segment
    man = w.create_file("/c", 5103, 6 * bs, bs)
    assert w.arena_stats()["used_bytes"] == 6 * bs
    with F.CurvineFileSystem(F.client_conf()) as fs:
        fs.load_namespace(man)
        assert _read_all(fs, "/c") == synth.file_bytes(5103, 6 * bs, bs)
    # the three 2 MiB tails are free: three 2 MiB blocks fit, a fourth does not (capacity 24 MiB = 3 segments)
    man2 = w.create_file("/d", 5104, 3 * (2 << 20), 2 << 20)
    with pytest.raises(F.FsError) as ei:
        w.create_file("/e", 5105, 1 << 20, 1 << 20)
    assert ei.value.kind == 17 and "full" in ei.value.msg  # DiskOutOfSpace
    w.delete_file(5103, 6)
    assert w.arena_stats()["used_bytes"] == 3 * (2 << 20)
    man3 = w.create_file("/f", 5106, 4 * bs, bs)  # lands in the freed (coalesced) space
    with F.CurvineFileSystem(F.client_conf()) as fs:
        fs.load_namespace(man2)
        fs.load_namespace(man3)
        assert _read_all(fs, "/d") == synth.file_bytes(5104, 3 * (2 << 20), 2 << 20)
        assert _read_all(fs, "/f") == synth.file_bytes(5106, 4 * bs, bs)
    with pytest.raises(F.FsError) as ei:
        w.create_file("/g", 5107, 9 << 20, 9 << 20)  # a block larger than a segment
    assert "does not fit an arena segment" in ei.value.msg


def test_worker_restart_rebuilds_the_arena_from_the_descriptors():
    d = _mk()
    try:
        dirs = ["[MEM:16MB]" + d + "/m0"]
        w = F.MiniWorker(dirs, extra_worker=ARENA % ("8MB", "0ms"))
        man = w.create_file("/r1", 5201, (3 << 20) + 5, 1 << 20)
        man_b = w.create_file("/r2", 5202, 2 << 20, 1 << 20)
        w.delete_file(5201, 1)  # only block 0 of r1: a hole in the allocation map
        used = w.arena_stats()["used_bytes"]
        w.stop()
        w = F.MiniWorker(dirs, extra_worker=ARENA % ("8MB", "0ms"), port=0)
        assert w.metrics()["num_blocks"] == 5
        assert w.arena_stats()["used_bytes"] == used
        # the new worker listens on another port: the manifests' addresses are rewritten (_port)
        with F.CurvineFileSystem(F.client_conf()) as fs:
            fs.load_namespace(_port(man_b, w.port))
            assert _read_all(fs, "/r2") == synth.file_bytes(5202, 2 << 20, 1 << 20)
            fs.load_namespace(_port(man, w.port))
            r = fs.open("/r1")
            r.seek(1 << 20)
            assert r.read_full(3 << 20) == synth.file_bytes(5201, (3 << 20) + 5, 1 << 20)[1 << 20:]
            r.complete()
        # the freed 1 MiB is handed out again
        man_c = w.create_file("/r3", 5203, 1 << 20, 1 << 20)
        assert w.arena_stats()["used_bytes"] == used + (1 << 20)
        w.stop()
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _port(manifest, port):
    """block <id> <len> <type> <crc32> <crc32c> <h|-> <host:port:worker_id,...>: point every location at `port`."""
    out = []
    for line in manifest.splitlines():
        f = line.split()
        if f and f[0] == "block" and f[-1] != "-":
            f[-1] = ",".join(":".join([a.split(":")[0], str(port), a.split(":")[2]]) for a in f[-1].split(","))
        out.append(" ".join(f))
    return "\n".join(out) + "\n"


def test_write_path_lands_in_the_arena_and_reads_back(arena_worker):
    """WriteBlock Open -> Running x N -> Complete (write_handler.rs:90-300) into an arena dir; write-side CRCs == read-side."""
    w, _ = arena_worker
    data = synth.file_bytes(77, (2 << 20) + 999, 1 << 20)
    with F.CurvineFileSystem(F.client_conf(short_circuit=True)) as fs:
        wr = fs.create("/w1", 5301, 1 << 20, w.port, chunk_size=65536)
        for o in range(0, len(data), 300000):
            wr.write(data[o:o + 300000])
        man = wr.complete()
        assert w.arena_stats()["used_bytes"] == (2 << 20) + 4096
        assert _read_all(fs, "/w1") == data
        # a cancelled writer gives its extent back
        wr = fs.create("/w2", 5302, 1 << 20, w.port)
        wr.write(b"x" * 1000)
        wr.complete(cancel=True)
        assert w.arena_stats()["used_bytes"] == (2 << 20) + 4096
    with F.CurvineFileSystem(F.client_conf(short_circuit=False)) as fs:
        fs.load_namespace(man)
        assert _read_all(fs, "/w1") == data


def test_two_mem_dirs_place_blocks_round_robin_by_hint():
    d = _mk()
    try:
        from curvine_b200 import _lib
        w = F.MiniWorker(["[MEM:8MB]" + d + "/m0", "[MEM:8MB]" + d + "/m1"], extra_worker=ARENA % ("8MB", "0ms"))
        _lib.lib().cv_synth_set_shard_world(2)
        try:
            man = w.create_file("/s", 5401, 8 << 20, 1 << 20, threads=4)
        finally:
            _lib.lib().cv_synth_set_shard_world(0)
        from oracle import layout
        for b in range(8):
            stub = layout.block_path("%s/m%d/curvine" % (d, b % 2), layout.create_block_id(5401, b))
            assert os.path.exists(stub), (b, stub)
        with F.CurvineFileSystem(F.client_conf()) as fs:
            fs.load_namespace(man)
            assert _read_all(fs, "/s") == synth.file_bytes(5401, 8 << 20, 1 << 20)
        w.stop()
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("arena", [True, False])
def test_framed_read_over_the_same_host_unix_socket(arena):
    """[b200] local_unix_socket: block connections to a worker on this host go to its abstract unix socket (net.h), same frames and
    handlers as TCP; a client without the option, and a worker without the socket, keep using TCP."""
    d = _mk()
    try:
        w = F.MiniWorker(["[MEM:16MB]" + d + "/m0"], extra_worker=(ARENA % ("8MB", "0ms")) if arena else "")
        name = "@curvine-b200-worker-%d" % w.port
        assert any(name in line for line in open("/proc/net/unix")), "the worker listens on its abstract socket"
        n, bs, ino = (5 << 20) + 321, 1 << 20, 5501
        man = w.create_file("/u", ino, n, bs)
        want = synth.file_bytes(ino, n, bs)
        for unix in (True, False):
            with F.CurvineFileSystem(F.client_conf(short_circuit=False, b200="local_unix_socket = %s\n" % ("true" if unix else "false"))) as fs:
                fs.load_namespace(man)
                before = sum(1 for line in open("/proc/net/unix") if name in line)
                r = fs.open("/u")
                assert r.read(100) == want[:100]
                during = sum(1 for line in open("/proc/net/unix") if name in line)
                assert (during > before) == unix  # a connected server-side endpoint carries the listener's name
                r.seek(bs + 5)
                assert r.read_full(n) == want[bs + 5:]
                r.complete()
        w.stop()
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_random_create_delete_sequences_never_overlap_extents_and_account_every_byte():
    """Allocator model check through the worker: files of random block sizes come and go in a small arena (tiny segments, so that
    segment tails, first-fit reuse and coalescing all happen); after every step EVERY live file still reads back bit-exact (an
    overlapping extent would corrupt one of them) and used_bytes equals the sum of the 4 KiB-rounded block lengths."""
    import random
    rnd = random.Random(20260921)
    d = _mk()
    try:
        w = F.MiniWorker(["[MEM:12MB]" + d + "/m0"], extra_worker=ARENA % ("2MB", "0ms"))
        live = {}  # inode -> (n, bs, manifest)
        ino = 9000
        with F.CurvineFileSystem(F.client_conf()) as fs:
            for step in range(120):
                if live and (rnd.random() < 0.45 or sum(_rounded(n, bs) for n, bs, _ in live.values()) > (9 << 20)):
                    victim = rnd.choice(sorted(live))
                    n, bs, _ = live.pop(victim)
                    w.delete_file(victim, (n + bs - 1) // bs)
                else:
                    bs = rnd.choice([4096, 8192, 65536, 100 * 4096, 1 << 20, (2 << 20)])
                    n = rnd.randint(1, 5 * bs) if bs < (1 << 20) else rnd.randint(1, 2 * bs)
                    ino += 1
                    try:
                        man = w.create_file("/r%d" % ino, ino, n, bs, threads=rnd.choice([1, 3]))
                    except F.FsError as e:
                        assert e.kind == 17, e  # arena full (fragmentation included): a legal answer, nothing may be half-created
                        nb = (n + bs - 1) // bs
                        w.delete_file(ino, nb)  # blocks committed before the failing one
                        continue
                    live[ino] = (n, bs, man)
                    fs.load_namespace(man)
                assert w.arena_stats()["used_bytes"] == sum(_rounded(n, bs) for n, bs, _ in live.values()), step
                for i in rnd.sample(sorted(live), min(3, len(live))):
                    n, bs, _ = live[i]
                    assert _read_all(fs, "/r%d" % i, size=1 << 20) == synth.file_bytes(i, n, bs), (step, i)
            for i, (n, bs, _) in live.items():
                assert _read_all(fs, "/r%d" % i, size=1 << 20) == synth.file_bytes(i, n, bs), i
        w.stop()
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _rounded(n, bs):
    total, left = 0, n
    while left > 0:
        b = min(bs, left)
        total += (b + 4095) // 4096 * 4096
        left -= b
    return total
