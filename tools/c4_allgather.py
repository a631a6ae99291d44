"""Config C4 (model distribution): a checkpoint file, mem tier, every GPU ends up holding ALL bytes in file order.

  torchrun --nproc-per-node G tools/c4_allgather.py --gib 70

Each rank ingests its round-robin shard (cv_read_device_sharded, CRC-32C verified on the GPU), then the exchange runs two ways:
  A  NCCL all_gather_into_tensor (in place) + cvk_deinterleave_blocks            (collective, then a 2N HBM pass)
  B  cvk_gather_shards_p2p: ONE kernel pulls every block straight from its owner's HBM over NVLink (peer pointers from
     torch symmetric memory) into file order -- no gathered staging buffer, no second pass
and every GPU re-verifies the whole file with K1 against the manifest.  Times are CUDA events, max over ranks."""
import argparse
import ctypes
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BLOCK = 4 << 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=70.0)
    ap.add_argument("--skip-nccl", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from curvine_b200 import _lib, fs as F, kernels as K
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    os.dup2(2, 1) if rank != 0 else None
    n = int(a.gib * (1 << 30)) // BLOCK * BLOCK
    nb = n // BLOCK
    per = (nb + world - 1) // world
    payload = [None, None]
    d = None
    if rank == 0:
        d = tempfile.mkdtemp(prefix="cvc4_", dir="/dev/shm")
        w = F.MiniWorker(["[MEM]" + d], hostname="localhost")
        L.cv_synth_set_shard_world(world)
        man = w.create_file("/ckpt", 777, n, BLOCK, threads=64)
        payload = [man, w.port]
    dist.broadcast_object_list(payload, src=0)
    man = payload[0]
    exp = np.zeros(nb, dtype=np.uint32)
    for line in man.splitlines():
        if line.startswith("block "):
            f = line.split()
            exp[int(f[1]) & 0xFFFFFF] = int(f[5], 16)
    res = {"world": world, "file_bytes": n, "blocks": nb}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    try:
        conf = F.client_conf(hostname="localhost", short_circuit=True,
                             b200='device = %d\nfetch_threads = 8\nzero_copy = true\nregister_cache = "%dGB"\ncopy_group = 8\nverify_batch = 16\n' % (local, int(a.gib / world * 1.5) + 2))
        fs = F.CurvineFileSystem(conf)
        fs.load_namespace(man)
        final = torch.empty(n, dtype=torch.uint8, device="cuda")
        # the shard lives in symmetric memory: every rank gets a directly loadable pointer to every peer's shard
        import torch.distributed._symmetric_memory as symm_mem
        shard = symm_mem.empty(per * BLOCK, dtype=torch.uint8, device="cuda")
        hdl = symm_mem.rendezvous(shard, dist.group.WORLD.group_name)
        peers = [int(p) for p in hdl.buffer_ptrs]
        stream = torch.cuda.current_stream().cuda_stream
        # ---- ingest (twice: the second pass has the mappings registered)
        for rep in range(2):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            r = fs.open("/ckpt")
            got = r.read_device_sharded(rank, world, shard.data_ptr(), per * BLOCK, stream)
            s, bad, ver = r.verify()
            e1.record(); e1.synchronize()
            r.complete()
            assert bad == 0 and ver == len(range(rank, nb, world))
            res["ingest_ms_rep%d" % rep] = maxr(e0.elapsed_time(e1))
        res["ingest_GBps"] = n / res["ingest_ms_rep1"] / 1e6
        # ---- full-file verify helper (K1 over the final buffer on every GPU)
        d_off = torch.arange(nb, dtype=torch.int64, device="cuda") * BLOCK
        d_len = torch.full((nb,), BLOCK, dtype=torch.int64, device="cuda")
        d_exp = torch.from_numpy(exp.view(np.int32)).cuda()
        d_crc = torch.empty(nb, dtype=torch.int32, device="cuda")

        def verify_final(tag):
            d_bad = torch.zeros(1, dtype=torch.int32, device="cuda")
            e0, e1 = ev(), ev()
            e0.record()
            K.crc_blocks_raw(final.data_ptr(), d_off, d_len, nb, 1, n, d_crc)
            K.verify_crcs(d_crc, d_exp, d_bad)
            e1.record(); e1.synchronize()
            assert int(d_bad.item()) == 0, "%s: %d bad blocks" % (tag, int(d_bad.item()))
            res[tag + "_verify_ms"] = maxr(e0.elapsed_time(e1))

        # ---- B: fused peer gather over NVLink
        for rep in range(3):
            final.zero_()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            K.gather_shards_p2p(peers, BLOCK, nb, n, final)
            e1.record(); e1.synchronize()
            res["p2p_gather_ms_rep%d" % rep] = maxr(e0.elapsed_time(e1))
        verify_final("p2p")
        torch.cuda.synchronize(); dist.barrier()
        res["p2p_gather_GBps_into_each_gpu"] = n / res["p2p_gather_ms_rep2"] / 1e6
        res["p2p_nvlink_GBps_per_gpu"] = n * (world - 1) / world / res["p2p_gather_ms_rep2"] / 1e6
        # ---- A: NCCL all-gather + de-interleave
        if not a.skip_nccl:
            gathered = torch.empty(world * per * BLOCK, dtype=torch.uint8, device="cuda")
            mine = gathered[rank * per * BLOCK:(rank + 1) * per * BLOCK]
            mine.copy_(shard)
            for rep in range(2):
                final.zero_()
                torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                dist.all_gather_into_tensor(gathered, mine)
                e1.record()
                K.deinterleave_blocks(gathered, per * BLOCK, world, BLOCK, nb, n, final)
                e2.record(); e2.synchronize()
                res["nccl_allgather_ms_rep%d" % rep] = maxr(e0.elapsed_time(e1))
                res["deinterleave_ms_rep%d" % rep] = maxr(e1.elapsed_time(e2))
            verify_final("nccl")
            res["nccl_total_ms"] = res["nccl_allgather_ms_rep1"] + res["deinterleave_ms_rep1"]
        fs.close()
    finally:
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            w.stop()
            shutil.rmtree(d, ignore_errors=True)
    if rank == 0:
        sys.stdout.write(json.dumps(res) + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
