"""The oracle against the reference's own known answers and the committed golden vectors (CPU only)."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import clib, layout, synth
from oracle import crc as C
from oracle import reader_model as RM
from oracle import wire as W

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wire_vectors.json")))


def test_status_kat_orpc_common_test():
    """orpc/tests/common_test.rs:18-30 -- the only literal known answer on the path."""
    assert W.status_encode(W.REQ_RUNNING, W.RESP_ERROR) == 19 == GOLD["status_running_error"]
    assert W.status_decode(19) == (W.REQ_RUNNING, W.RESP_ERROR)
    for rq in (W.REQ_HEARTBEAT, W.REQ_RPC, W.REQ_OPEN, W.REQ_RUNNING, W.REQ_CANCEL, W.REQ_COMPLETE):
        for rs in (W.RESP_SUCCESS, W.RESP_ERROR, W.RESP_UNDEFINED):
            assert W.status_decode(W.status_encode(rq, rs)) == (rq, rs)
    assert W.status_encode(W.REQ_OPEN, W.RESP_UNDEFINED) & 0xFF == 0xF2
    assert W.status_encode(W.REQ_COMPLETE, W.RESP_SUCCESS) & 0xFF == 0x05


def test_golden_wire_vectors_appendix_a():
    blk = synth.block_bytes(1001, 0, 4 << 20)
    bid = layout.create_block_id(1001, 0)
    assert bid == GOLD["block_id"] == 16793993216
    reqs, resps = W.block_read_exchange(bid, blk, 131072, 0x0102030405060708)
    assert reqs[0].hex() == GOLD["open_request"] and len(reqs[0]) == 51
    assert resps[0].hex() == GOLD["open_response"]
    assert reqs[1].hex() == GOLD["running_request_1"] and len(reqs[1]) == 22
    assert resps[1][:22].hex() == GOLD["running_response_1_prefix"] == "00020012000000005103010203040506070800000001"
    assert reqs[-1].hex() == GOLD["complete_request"] and len(reqs[-1]) == 47
    assert len(reqs) == 34 and sum(map(len, resps)) == GOLD["response_stream_bytes"] == (4 << 20) + 34 * 22 + 13
    assert W.DataHeaderProto(65536).encode().hex() == "0880800410001800"
    assert layout.block_path("/data/curvine", bid) == GOLD["block_path"]


def test_golden_write_protocol_vectors():
    """WriteBlock (80) frames: BlockWriteRequest embeds ExtendedBlockProto (worker.proto:10-18, common.proto:98-104)."""
    bid = GOLD["block_id"]
    req = W.BlockWriteRequest(bid, 0, W.STORAGE_MEM, 1, 0, 4 << 20, False, "cv", 131072)
    enc = W.encode(W.request(80, W.REQ_OPEN, 0x0102030405060708, 0, req.encode()))
    assert enc.hex() == GOLD["write_open_request"]
    assert W.BlockWriteRequest.decode(req.encode()) == req
    assert enc[8] == 80 and enc[9] == 0xF2
    rsp = W.BlockWriteResponse(bid, None, 0, 4 << 20, W.STORAGE_MEM)
    assert W.BlockWriteResponse.decode(rsp.encode()) == rsp
    assert W.encode(W.success(W.request(80, W.REQ_OPEN, 0x0102030405060708, 0), rsp.encode())).hex() == GOLD["write_open_response"]
    # required string client_name is emitted even when empty (prost: required fields always encoded)
    assert W.BlockWriteRequest(1, 0, 0, 1, 0, 1, False, "", 1).encode().endswith(bytes([0x2a, 0x00, 0x30, 0x01]))


def test_frame_roundtrip_and_limits():
    rng = np.random.default_rng(0)
    for hl, dl in [(0, 0), (5, 0), (0, 1), (13, 131072), (29, 777)]:
        m = W.Message(81, W.REQ_RUNNING, W.RESP_SUCCESS, -123456789, 42, rng.bytes(hl), rng.bytes(dl))
        enc = W.encode(m)
        total = int.from_bytes(enc[:4], "big", signed=True)
        assert total == 18 + hl + dl and len(enc) == 4 + total  # rpc_message.rs:305
        (d,), used = W.decode_stream(enc)
        assert used == len(enc) and d == m
    # data_len < 0 and > 16 MiB rejected (rpc_message.rs:329-334)
    bad = bytearray(W.encode(W.Message(81, 3, 0, 1, 1)))
    bad[0:4] = (17).to_bytes(4, "big")
    with pytest.raises(W.WireError):
        W.decode_protocol(bytes(bad))
    bad[0:4] = (18 + 16 * 1024 * 1024 + 1).to_bytes(4, "big")
    with pytest.raises(W.WireError):
        W.decode_protocol(bytes(bad))
    bad[0:4] = (18 + 16 * 1024 * 1024).to_bytes(4, "big")
    W.decode_protocol(bytes(bad))
    # heartbeats are skipped by receivers (rpc_frame.rs:255-259)
    hb = W.encode(W.Message(0, W.REQ_HEARTBEAT, W.RESP_UNDEFINED, -1, -1))
    msgs, _ = W.decode_stream(hb + W.encode(W.Message(81, 3, 0, 5, 6, b"", b"xyz")))
    assert len(msgs) == 1 and msgs[0].data == b"xyz"


def test_error_body_and_proto_roundtrip():
    assert W.decode_error(W.encode_error(8, "File /x not exists")) == (8, "File /x not exists")
    assert bytes.fromhex(GOLD["error_response"])[9] == 19
    r = W.BlockReadRequest(-5, 1 << 40, 7, 65536, True, False, 123, 456)
    assert W.BlockReadRequest.decode(r.encode()) == r
    assert W.BlockReadRequest(id=9).encode().hex() == "0809100018002000280040014880808002" + "50808040"  # proto defaults kept
    for p in (W.BlockReadResponse(1, 2, "/p/blk_1", 0), W.BlockReadResponse(1, 2, None, 4)):
        assert W.BlockReadResponse.decode(p.encode()) == p
    with pytest.raises(W.WireError):
        W.BlockReadRequest.decode(b"\x08\x01")  # required fields missing


def test_crc_check_values_and_implementations_agree():
    assert C.crc32(C.CHECK_INPUT) == C.CHECK_IEEE == 0xCBF43926 == GOLD["crc_check"]["crc32"]
    assert C.crc32c(C.CHECK_INPUT) == C.CHECK_CASTAGNOLI == 0xE3069283 == GOLD["crc_check"]["crc32c"]
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 8, 9, 15, 16, 17, 4095, 4096, 65537):
        d = rng.bytes(n)
        for pid, poly in C.POLYS.items():
            vals = {C.crc_bitwise(d, poly) if n < 5000 else C.crc_table(d, poly), C.crc_table(d, poly), clib.crc(pid, d),
                    clib.lib().cvo_crc_bitwise(pid, 0, d, n)}
            assert len(vals) == 1
        assert clib.crc(0, d) == zlib.crc32(d)
    blk = synth.block_bytes(1001, 0, 4 << 20)
    assert clib.crc(0, blk) == GOLD["block_crc32"] and clib.crc(1, blk) == GOLD["block_crc32c"]
    assert blk[:32].hex() == GOLD["synth_block_1001_0_first32"]


def test_bench_checksum_semantics():
    """curvine_bench.rs:37-48,222-231: u64 sum of crc32 over whole read buffers (stale tail on a short last read)."""
    d = np.random.default_rng(2).bytes(300000)
    assert C.bench_checksum(d, 131072) == clib.bench_checksum(d, 131072)
    assert C.bench_checksum(d, 131072, stale_tail=False) == sum(zlib.crc32(d[i:i + 131072]) for i in range(0, len(d), 131072))
    assert C.bench_checksum(d[:262144], 131072) == C.bench_checksum(d[:262144], 131072, stale_tail=False)
    assert clib.bench_checksum(synth.block_bytes(1001, 0, 4 << 20), 131072) == GOLD["bench_sum_crc32_128k"]


def test_python_bench_checksum_chaining_equals_block_combine():
    """curvineBench.py:30-52 chains zlib.crc32 across buffers; the product never chains -- it CRCs pieces independently (per
    warp segment on the GPU, per block on the host) and combines them with x^(8*len) multipliers.  Same numbers: the reference's
    own Python tool is therefore a second in-tree anchor (next to crc32fast) for CRC-32 == zlib and for the combine algebra."""
    from curvine_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(11)
    bufs = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in (131072, 131072, 5, 0, 4097, 131072 - 1)]
    chained = 0
    for b in bufs:
        chained = zlib.crc32(b, chained)
    assert chained == zlib.crc32(b"".join(bufs))
    # combine of independent per-buffer CRCs (oracle restatement, and the product's host CRC feeding the same algebra)
    acc = 0
    for b in bufs:
        c_or = C.crc32(b)
        arr = np.frombuffer(b, dtype=np.uint8)
        c_host = L.cv_host_crc(0, arr.ctypes.data if len(b) else None, len(b))
        assert c_or == c_host == zlib.crc32(b)
        acc = C.crc_combine(acc, c_or, len(b), C.POLY_IEEE)
    assert acc == chained
    # two "threads": the tool's fold of per-thread values
    want = zlib.crc32(zlib.crc32(b"".join(bufs[3:])).to_bytes(4, "big"), zlib.crc32(zlib.crc32(b"".join(bufs[:3])).to_bytes(4, "big"), 0))
    assert C.python_bench_checksum([bufs[:3], bufs[3:]]) == want


def test_block_id_and_layout():
    """inode_id.rs:100-118 + block_meta.rs:199-237."""
    rng = np.random.default_rng(3)
    for _ in range(100):
        ino, seq = int(rng.integers(1, layout.ID_MASK)), int(rng.integers(0, layout.SEQ_MASK))
        bid = layout.create_block_id(ino, seq)
        assert layout.block_inode(bid) == ino and layout.block_seq(bid) == seq
    with pytest.raises(ValueError):
        layout.create_block_id(layout.ID_MASK + 1, 0)
    assert layout.block_dir("/b", (5 << 48) | (9 << 32) | 1) == "/b/active/b5/b9"
    assert layout.block_path("/b", 7, recovering=True) == "/b/staging/blk_7"


def test_split_slices_fs_reader_parallel():
    """fs_reader_parallel.rs:194-220."""
    flat = sorted(s for sub in RM.split(1000, 300, 3) for s in sub)
    cur = 0
    for a, b in flat:
        assert a == cur and b > a
        cur = b
    assert cur == 1000
    assert RM.split(1000, 300, 1) == [[(0, 1000)]]
    assert RM.split(0, 300, 3) == []
    assert RM.split(1000, 300, 3)[0] == [(0, 300), (900, 1000)]


def _det(threshold=10, enabled=True, size=1000, **kw):
    return RM.ReadDetector(RM.ClientConf(enable_smart_prefetch=enabled, sequential_read_threshold=threshold, **kw), size)


def test_read_detector_state_machine():
    """read_detector.rs:242-528, restated."""
    d = _det()
    assert d.enabled and d.is_sequential() and d.seq_count == 0 and d.read_parallel == 1
    d = _det(enabled=False)
    d.record_seek()
    assert d.is_sequential() and not d.record_read(0, 100)
    d = _det()
    d.record_seek()
    assert d.is_random() and d.seq_count == 0
    d.record_seek()
    assert d.is_random()
    d = _det(3)
    assert [d.record_read(i * 100, (i + 1) * 100) for i in range(4)] == [False] * 4 and d.seq_count == 4 and d.is_sequential()
    d = _det(3)
    d.record_seek()
    assert not d.record_read(0, 100) and not d.record_read(100, 200) and d.is_random()
    assert d.record_read(200, 300) and d.is_sequential() and d.seq_count == 3
    d = _det()
    d.record_read(0, 100), d.record_read(100, 200)
    d.record_read(500, 600)
    assert d.seq_count == 0 and d.is_sequential()
    d = _det(3)
    d.record_seek()
    d.record_read(0, 100), d.record_read(500, 600)
    assert d.seq_count == 0
    d.record_read(600, 700), d.record_read(700, 800), d.record_read(800, 900)
    assert d.seq_count == 3 and d.is_sequential()
    d = _det(5)
    d.record_seek()
    assert [d.record_read(i * 100, (i + 1) * 100) for i in range(5)] == [False] * 4 + [True]
    d = _det(large_file_size=1 << 30, max_read_parallel=8, size=100)
    assert d.read_parallel == 1
    assert 1 < _det(large_file_size=1 << 30, max_read_parallel=8, size=10 << 30).read_parallel <= 8
    d = _det()
    d.record_read(100, 100)
    assert d.seq_count == 1 and d.is_sequential()
    # reference defaults: 16 GiB -> 2, 128 GiB -> 8, 70 GiB -> 7
    assert [RM.ReadDetector(RM.ClientConf(), n << 30).read_parallel for n in (16, 128, 70)] == [2, 8, 7]


def test_reader_model_semantics():
    """block_test.rs-shaped walk over the model itself: chunks never span blocks, seek rules, holes."""
    bs, n = 1 << 20, (3 << 20) + 12345
    data = bytearray(synth.file_bytes(5, n, bs))
    blocks = [RM.BlockSpec(layout.create_block_id(5, i), min(bs, n - i * bs), hole=(i == 2)) for i in range(4)]
    data[2 * bs:3 * bs] = bytes(bs)
    data = bytes(data)
    r = RM.ReaderModel(RM.FileModel(blocks, data), RM.ClientConf(read_chunk_size=65536))
    first = r.read(1 << 30)
    assert len(first) == 65536  # read() hands out at most the current chunk
    assert first + r.read_full(n) == data and r.pos == n and r.read(10) == b""
    r.seek(bs - 1)
    assert r.read(1 << 30) == data[bs - 1:bs]  # last byte of a block: chunk ends at the block boundary
    r.seek(n)
    assert r.read(1) == b""
    # FsReaderBase rejects pos > len (fs_reader_base.rs:123-133) but FsReaderParallel::seek clamps to its last
    # slice end first (fs_reader_parallel.rs:175-181), so through FsReader a seek past EOF succeeds and reads return 0
    r.seek(n + 1)
    assert r.pos == n + 1 and r.read(1) == b""
    with pytest.raises(RM.FsError):
        RM.BaseModel(r.f, 65536).seek(n + 1)
    with pytest.raises(RM.FsError):
        r.seek(-1)
    r.seek(0)
    segs = r.fuse_read(bs + 65536 - 1024, 200000)
    assert b"".join(segs) == data[bs + 65536 - 1024:bs + 65536 - 1024 + 200000]
    # after a seek the worker streams chunk_size pieces from the seek offset (no re-alignment): read_handler.rs:143-158
    assert [len(x) for x in segs] == [65536, 65536, 65536, 200000 - 3 * 65536]


def test_c_generator_equals_python_generator():
    """oracle/oracle.c cvo_synth_block == oracle/synth.py block_bytes (xoshiro256** seeded by splitmix64) -- the GPU
    scale test uses the C one for speed."""
    for fid, b, n in [(1001, 0, 4096), (8901, 17, 100001), (5, 123456, 1), (2 ** 39, 2 ** 23, 65536 + 7)]:
        assert clib.synth_block(fid, b, n).tobytes() == synth.block_bytes(fid, b, n)
