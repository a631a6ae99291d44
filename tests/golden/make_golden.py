"""Regenerates tests/golden/wire_vectors.json from the oracle restatement (oracle/wire.py).

The reference holds no golden frame dump: these are OUR restatement's vectors, derived from
rpc_message.rs:301-338 / worker.proto:38-60 / block_client.rs:222-300 and pinned by
  * the reference's only literal KAT, Status(Running, Error).encode() == 19 (orpc/tests/common_test.rs:18-30)
  * SURVEY.md Appendix A's hand-worked example (same bytes),
  * the total_len = 18 + header_len + data_len identity and encode->decode round trips (tests/test_oracle.py).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import crc as C  # noqa: E402
from oracle import layout, synth, wire as W  # noqa: E402


def main():
    block_id = layout.create_block_id(1001, 0)
    blk = synth.block_bytes(1001, 0, 4 * 1024 * 1024)
    reqs, resps = W.block_read_exchange(block_id, blk, 131072, 0x0102030405060708)
    seek_hdr = W.DataHeaderProto(65536, False, False).encode()
    seek_req = W.encode(W.request(W.RPC_CODE_READ_BLOCK, W.REQ_RUNNING, 0x0102030405060708, 2, seek_hdr))
    err = W.encode(W.error(W.request(W.RPC_CODE_READ_BLOCK, W.REQ_RUNNING, 7, 3), 10000, "block 5 not exits"))
    out = {
        "status_running_error": W.status_encode(W.REQ_RUNNING, W.RESP_ERROR),
        "status_bytes": {"open_req": W.status_encode(W.REQ_OPEN, W.RESP_UNDEFINED) & 0xFF,
                         "running_req": W.status_encode(W.REQ_RUNNING, W.RESP_UNDEFINED) & 0xFF,
                         "complete_req": W.status_encode(W.REQ_COMPLETE, W.RESP_UNDEFINED) & 0xFF,
                         "running_ok": W.status_encode(W.REQ_RUNNING, W.RESP_SUCCESS) & 0xFF,
                         "running_err": W.status_encode(W.REQ_RUNNING, W.RESP_ERROR) & 0xFF},
        "block_id": block_id,
        "block_path": layout.block_path("/data/curvine", block_id),
        "open_request": reqs[0].hex(),
        "open_response": resps[0].hex(),
        "running_request_1": reqs[1].hex(),
        "running_response_1_prefix": resps[1][:22].hex(),
        "running_response_1_len": len(resps[1]),
        "running_request_after_seek_65536": seek_req.hex(),
        "complete_request": reqs[-1].hex(),
        "complete_response": resps[-1].hex(),
        "error_response": err.hex(),
        "n_request_frames": len(reqs),
        "response_stream_bytes": sum(len(r) for r in resps),
        "block_crc32": C.crc32(blk),
        "block_crc32c": C.crc32c(blk),
        "bench_sum_crc32_128k": C.bench_checksum(blk, 131072),
        "synth_block_1001_0_first32": blk[:32].hex(),
        "write_open_request": W.encode(W.request(W.RPC_CODE_WRITE_BLOCK, W.REQ_OPEN, 0x0102030405060708, 0,
                                                 W.BlockWriteRequest(block_id, 0, W.STORAGE_MEM, 1, 0, 4 << 20, False, "cv", 131072).encode())).hex(),
        "write_open_response": W.encode(W.success(W.request(W.RPC_CODE_WRITE_BLOCK, W.REQ_OPEN, 0x0102030405060708, 0),
                                                  W.BlockWriteResponse(block_id, None, 0, 4 << 20, W.STORAGE_MEM).encode())).hex(),
        "write_complete_request": W.encode(W.request(W.RPC_CODE_WRITE_BLOCK, W.REQ_COMPLETE, 0x0102030405060708, 33,
                                                     W.BlockWriteRequest(block_id, 4 << 20, W.STORAGE_MEM, 1, 4 << 20, 4 << 20, False, "cv", 0).encode())).hex(),
        "crc_check": {"input": "123456789", "crc32": C.CHECK_IEEE, "crc32c": C.CHECK_CASTAGNOLI},
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wire_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
