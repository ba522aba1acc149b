"""Run under torchrun (one rank per GPU): row-sharded SGEMM with NCCL broadcast of B, every
rank checks sampled rows of its C panel against the CPU oracle.  Exit code != 0 on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import laser_b200 as L
import oracle as O
from laser_b200.rowshard import gemm_rowsharded, partition_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M, N, K = int(os.environ.get("RS_M", 4096)), 2048, 3000
lo, hi = partition_rows(M, world)[rank]
Ml = hi - lo
A = torch.empty(M * K, dtype=torch.float32, device="cuda"); L.fill_uniform_f32(A, M * K, 7, 0, 1)
A = A.view(M, K)[lo:hi]
Bref = torch.empty(K * N, dtype=torch.float32, device="cuda"); L.fill_uniform_f32(Bref, K * N, 8, 0, 1)   # what the root holds
Bref = Bref.view(K, N)
colmajor = os.environ.get("RS_COLMAJOR", "0") == "1"       # B stored column-major (= [N][K] row-major), passed with strides (1, K)
stored = Bref.t().contiguous() if colmajor else Bref.clone()
if rank != 0:
    stored.fill_(float("nan"))                             # B is an input that lives on the root only
B = stored.t() if colmajor else stored
C = torch.full((Ml, N), 3.0, dtype=torch.float32, device="cuda")
gemm_rowsharded(Ml, N, K, 0.5, A, B, -1.25, C, src=0)      # the C ABI: B travels over NCCL (prepared panels or raw) + this rank's rows
torch.cuda.synchronize()
rows = np.unique(np.random.default_rng(rank).integers(0, Ml, 24))
a = A[rows].cpu().numpy(); b = Bref.cpu().numpy()
want = np.full((len(rows), N), 3.0, np.float32)
O.gemm_strided(len(rows), N, K, 0.5, a, K, 1, b, N, 1, -1.25, want, N, 1)
err = O.max_relative_error(C[rows].cpu().numpy(), want)
print("rank %d rows [%d,%d) max_rel_err %.3e" % (rank, lo, hi, err), flush=True)
ok = torch.tensor([1 if err < 1e-4 else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok.item() == 1 else 1)
