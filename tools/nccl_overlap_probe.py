"""torchrun probe (2+ GPUs): does an NCCL broadcast make progress next to the persistent tensor-core GEMM?
Times (CUDA events on each stream) a 134 MB broadcast on a side stream alone, next to the library's GEMM (either launch order),
and next to a cuBLAS bf16 GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import laser_b200 as L
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local)); L.init()
n = 8192
A = torch.rand(n, n, device="cuda") - 0.5; B = torch.rand(n, n, device="cuda") - 0.5; C = torch.empty(n, n, device="cuda")
Ab = A.bfloat16(); Bb = B.bfloat16()
X = torch.rand(n // 2, n, device="cuda")          # 134 MB payload
cs = torch.cuda.Stream(); s = torch.cuda.current_stream()
def gemm(): L.gemm_strided(n, n, n, 1.0, A, n, 1, B, n, 1, 0.0, C, n, 1)
def cublas(): torch.matmul(Ab, Bb)
def bcast():
    with torch.cuda.stream(cs): dist.broadcast(X, src=0)
def run(name, first, second, reps=5):
    ts = []
    for _ in range(reps + 2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        cs.wait_stream(s)
        if first == "bcast":
            with torch.cuda.stream(cs): e[0].record(cs)
            bcast()
            with torch.cuda.stream(cs): e[1].record(cs)
            e[2].record(s); second(); e[3].record(s)
        else:
            e[2].record(s); first(); e[3].record(s)
            with torch.cuda.stream(cs): e[0].record(cs)
            bcast()
            with torch.cuda.stream(cs): e[1].record(cs)
        torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]), max(e[2].elapsed_time(e[1]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[3]), e[0].elapsed_time(e[1]))))
    ts = ts[2:]
    t = torch.tensor([sum(x[i] for x in ts) / len(ts) for i in range(3)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("%-34s bcast %.3f ms | compute %.3f ms | both done after %.3f ms" % (name, *t.tolist()), flush=True)
noop = lambda: None
for _ in range(3): gemm(); cublas(); bcast()
run("bcast alone", "bcast", noop)
run("gemm alone", gemm, noop) if False else None
run("bcast first, then library GEMM", "bcast", gemm)
run("library GEMM first, then bcast", gemm, None)
run("bcast first, then cuBLAS bf16 x3", "bcast", lambda: (cublas(), cublas(), cublas()))
run("cuBLAS bf16 x3 first, then bcast", lambda: (cublas(), cublas(), cublas()), None)
dist.barrier(); dist.destroy_process_group()
