"""torchrun probe: what the row-sharded step costs next to its parts (broadcast alone through torch's NCCL and through the
library's communicator, the single-GPU product alone, the row-sharded call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import laser_b200 as L
from laser_b200 import rowshard as RS
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local)); L.init()
comm = RS.comm_from_torch_distributed()
M = N = K = 8192
A = torch.rand(M, K, device="cuda") - 0.5; B = torch.rand(K, N, device="cuda") - 0.5; C = torch.empty(M, N, device="cuda")
def timed(fn, steps=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()
res = {}
res["torch_bcast_only"] = timed(lambda: dist.broadcast(B, src=0))
res["gemm_only"] = timed(lambda: L.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, C, N, 1))
res["rowsharded (C ABI)"] = timed(lambda: RS.gemm_rowsharded(M, N, K, 1.0, A, B, 0.0, C, src=0, comm=comm))
res["rowsharded M=0 (bcast only, C ABI)"] = timed(lambda: RS.gemm_rowsharded_dev(comm, 0, N, K, 1.0, None, K, 1, B, N, 1, 0, 0.0, None, N, 1))
if rank == 0:
    for k, v in res.items(): print("world=%d %-36s %.3f ms" % (world, k, v), flush=True)
dist.barrier(); dist.destroy_process_group()
