"""curvine_b200/curvinefs.py: the read side of the reference's Python SDK surface (curvine-libsdk/python/curvinefs: CurvineClient,
CurvineReader) on the new C ABI.  Names, argument meaning and error behaviour of open / read / seek / close / read_range / head / tail /
get_file_status; bytes against the oracle generator."""
import os
import tempfile

import pytest

from curvine_b200 import curvinefs, fs as F
from oracle import synth


@pytest.fixture(scope="module")
def client():
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        with F.MiniWorker(["[MEM]" + d + "/m"]) as w:
            n, bs, ino = (5 << 20) + 4321, 1 << 20, 4501
            man = w.create_file("/data/a.bin", ino, n, bs)
            man += w.create_file("/data/empty", 4502, 0, bs)
            mpath = os.path.join(d, "namespace.manifest")
            open(mpath, "w").write(man)
            conf = os.path.join(d, "curvine.toml")
            open(conf, "w").write('namespace_manifest = "%s"\n' % mpath + F.client_conf())
            c = curvinefs.CurvineClient(conf, 8, 131072)
            yield c, synth.file_bytes(ino, n, bs)
            c.close()


def test_open_read_seek_close(client):
    c, want = client
    n = len(want)
    r = c.open("/data/a.bin")
    assert r.file_size == n
    assert r.read(0, 1000) == want[:1000]
    assert r.read(0, 70000) == want[1000:71000]            # sequential
    assert r.read(500, 10) == want[71500:71510]             # offset = bytes to skip from the current position
    r.seek((3 << 20) - 5)
    assert r.read(0, 11) == want[(3 << 20) - 5:(3 << 20) + 6]   # across a block boundary
    r.seek(n - 3)
    assert r.read(0, 100) == want[n - 3:] and r.read(0, 5) == b""   # end of file is not an error
    with pytest.raises(ValueError):
        r.seek(-1)
    with pytest.raises(ValueError):
        r.seek(n + 1)
    r.seek(0)
    with pytest.raises(IOError):
        r.read(-5, 1)                                        # "Position is negative"
    r.close()
    r.close()                                                # closing twice is not an error
    with pytest.raises(IOError):
        r.read(0, 1)


def test_read_range_head_tail_and_status(client):
    c, want = client
    n = len(want)
    st = c.get_file_status("/data/a.bin")
    assert st["len"] == n and st["name"] == "a.bin" and st["is_dir"] is False
    assert c.get_file_status("/data/nope") is None
    assert c.read_range("/data/a.bin", 0, 10) == want[:10]
    assert c.read_range("/data/a.bin", 123456, 2 << 20) == want[123456:123456 + (2 << 20)]
    assert c.read_range("/data/a.bin", -100, None) == want[-100:]      # negative offsets count from the end, None/-1 read to the end
    assert c.read_range("/data/a.bin", n - 7, -1) == want[n - 7:]
    assert c.read_range("/data/a.bin", 5, 0) == b""
    assert c.head("/data/a.bin", 4096) == want[:4096]
    assert c.tail("/data/a.bin", 4097) == want[-4097:]
    assert c.tail("/data/a.bin", n + 10) == want and c.tail("/data/empty", 10) == b""
    with pytest.raises(FileNotFoundError):
        c.read_range("/data/nope", 0, 1)
    with pytest.raises(ValueError):
        c.read_range("/data/a.bin", n, None)                             # "Offset exceeds file size"
    with pytest.raises(ValueError):
        c.read_range("/data/a.bin", 0, -2)
    with pytest.raises(ValueError):
        c.head("/data/a.bin", -1)
    with pytest.raises(IOError):
        c.open("/data/nope")


def test_control_plane_calls_are_unsupported(client):
    c, _ = client
    for call in (lambda: c.mkdir("/x", True), lambda: c.rm("/x"), lambda: c.ls("/"), lambda: c.rename("/a", "/b"), lambda: c.get_master_info()):
        with pytest.raises(F.FsError) as e:
            call()
        assert e.value.kind == 19  # ErrorKind::Unsupported (fs_error.rs:35-66)
