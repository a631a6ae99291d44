import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from curvine_b200 import _lib, kernels as K
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 8 << 20
def log(*a): print(rank, *a, file=sys.stderr, flush=True)
try:
    import torch.distributed._symmetric_memory as symm_mem
    shard = symm_mem.empty(n, dtype=torch.uint8, device="cuda")
    shard.fill_(rank + 1)
    hdl = symm_mem.rendezvous(shard, dist.group.WORLD.group_name)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    log("symm ptrs", [hex(p) for p in ptrs], "mine", hex(shard.data_ptr()))
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    dst = torch.zeros(n * world, dtype=torch.uint8, device="cuda")
    K.gather_shards_p2p(ptrs, 1 << 20, (n >> 20) * world, n * world, dst)
    torch.cuda.synchronize()
    log("symm gather ok", dst[::(1 << 20)].tolist()[:8])
except Exception as e:
    log("symm failed:", repr(e)[:400])
dist.barrier()
dist.destroy_process_group()
