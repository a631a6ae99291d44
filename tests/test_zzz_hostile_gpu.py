"""The GPU reader's framed path against a worker that lies (tests/test_hostile_peers.py has the host reader's side).  Frames are received
verbatim and validated ON THE GPU (K2: total_len / header_len / code / status / request-id and sequence-id echoes, rpc_message.rs:329-334,
raw_client.rs:100-116): a wrong echo, an error response or a frame of the wrong size must surface as an error from the read or its verify,
a truncated stream or a silent worker as an I/O error within the timeouts -- never as silently wrong bytes, a crash or a stall.
Written after round 2's last GPU run; sorts late on purpose."""
import random
import time

import pytest

from curvine_b200 import fs as F
from oracle import layout, synth
from test_hostile_peers import _LyingWorker

pytestmark = pytest.mark.gpu

MODES = ["honest", "wrong_req_id", "wrong_seq_id", "error_response", "truncated_payload", "huge_data_len", "negative_total", "random_bytes", "silence",
         "open_garbage_header", "open_len_lies", "longer_than_chunk", "negative_header_len", "open_error_body_garbage"]


@pytest.mark.parametrize("mode", MODES)
def test_device_reader_survives_a_lying_worker(cuda, mode):
    import torch
    ino, bs = 8401, 1 << 20
    n = bs
    data = synth.file_bytes(ino, n, bs)
    bid = layout.create_block_id(ino, 0)
    lw = _LyingWorker(data, mode, random.Random(5))
    man = "# m\nfile /lie %d %d %d 0\nblock %d %d 0 - - - localhost:%d:1\n" % (ino, n, bs, bid, n, lw.port)
    t0 = time.time()
    try:
        conf = F.client_conf(short_circuit=False, extra_client='conn_timeout_ms = 1000\ndata_timeout_ms = 1000\nrpc_timeout_ms = 1000\n',
                             b200='fetch_threads = 2\nverify_batch = 2\npinned_slots = 8\ncopy_group = 1\ngpu_chunk_size = "64KB"\n')
        with F.CurvineFileSystem(conf) as fs:
            fs.load_namespace(man)
            r = fs.open("/lie")
            dst = torch.zeros(n, dtype=torch.uint8, device=cuda)
            st = torch.cuda.current_stream().cuda_stream
            if mode == "honest":
                assert r.read_device(dst.data_ptr(), n, st) == n
                s, bad, ver = r.verify()
                torch.cuda.synchronize()
                assert bad == 0 and dst.cpu().numpy().tobytes() == data
            else:
                with pytest.raises(F.FsError) as e:
                    r.read_device(dst.data_ptr(), n, st)
                    r.verify()  # frame validation results come back with the verify
                assert e.value.kind > 0
            try:
                r.complete()
            except F.FsError:
                pass
    finally:
        lw.close()
    assert time.time() - t0 < 30


def test_connection_parked_with_an_answer_and_a_hang_up_behind_it_is_not_reused(cuda):
    """A worker that answers a block's Open / Running / Complete and then closes the connection (a restart between two reads): the client has
    parked that connection with the Complete's answer still unread, so the socket holds data AND the hang-up.  The next read must notice the
    hang-up and connect afresh -- not consume the answer, send its requests into a dead connection and fail."""
    import torch
    ino, bs = 8402, 1 << 20
    n = 3 * bs
    data = synth.file_bytes(ino, n, bs)
    # three blocks served by one lying worker that holds block 0's bytes for every id: use one block per file instead
    bid = layout.create_block_id(ino, 0)
    lw = _LyingWorker(data[:bs], "honest_then_hang_up", random.Random(1))
    man = "# m\nfile /hang %d %d %d 0\nblock %d %d 0 - - - localhost:%d:1\n" % (ino, bs, bs, bid, bs, lw.port)
    try:
        conf = F.client_conf(short_circuit=False, extra_client='conn_timeout_ms = 1000\ndata_timeout_ms = 1000\nrpc_timeout_ms = 1000\n',
                             b200='fetch_threads = 1\nverify_batch = 1\npinned_slots = 8\ncopy_group = 1\ngpu_chunk_size = "64KB"\n')
        with F.CurvineFileSystem(conf) as fs:
            fs.load_namespace(man)
            st = torch.cuda.current_stream().cuda_stream
            for attempt in range(4):
                r = fs.open("/hang")
                dst = torch.zeros(bs, dtype=torch.uint8, device=cuda)
                assert r.read_device(dst.data_ptr(), bs, st) == bs, attempt
                s, bad, ver = r.verify()
                torch.cuda.synchronize()
                assert bad == 0 and dst.cpu().numpy().tobytes() == data[:bs], attempt
                r.complete()
                time.sleep(0.2)  # the hang-up has certainly arrived at the parked connection
            assert fs.pool_stats()["opened"] >= 4  # one fresh connection per read: none of the dead ones was reused
    finally:
        lw.close()
