"""CPU-only, world_size 2 over gloo: the host logic of the row-sharded multi-GPU path (partition, B valid on rank 0
only and delivered to every rank, every rank its own rows).  The oracle stands in for the CUDA kernels as gemm_fn and
torch.distributed for the NCCL broadcast the library issues itself on a GPU box -- this tests plumbing, not compute;
the C entry with a stand-in NCCL is covered by tests/test_emulated_library.py (rowsharded)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_gemm(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC):
    import oracle as O
    import ctypes
    f32, i64, vp = ctypes.c_float, ctypes.c_int64, ctypes.c_void_p
    O.lib().oracle_gemm_strided_f32(M, N, K, alpha, vp(A.data_ptr()), rsA, csA, vp(B.data_ptr()), rsB, csB,
                                    beta, vp(C.data_ptr()), rsC, csC)


def _worker(rank, world, port, M, N, K, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from laser_b200.rowshard import gemm_rowsharded, partition_rows
    import oracle as O
    A = torch.from_numpy(O.fill_uniform_f32(M * K, 1, -1, 1).reshape(M, K))
    Bfull = torch.from_numpy(O.fill_uniform_f32(K * N, 2, -1, 1).reshape(K, N))
    C0 = torch.from_numpy(O.fill_uniform_f32(M * N, 3, -1, 1).reshape(M, N))
    lo, hi = partition_rows(M, world, align=16)[rank]
    B = Bfull.clone() if rank == 0 else torch.full((K, N), float("nan"))   # only rank 0 holds B
    C_local = C0[lo:hi].clone()
    gemm_rowsharded(hi - lo, N, K, 0.5, A[lo:hi], B, -1.25, C_local, src=0, gemm_fn=_oracle_gemm)
    assert torch.equal(B, Bfull)                                            # the broadcast delivered B
    want = C0.numpy().copy()
    O.gemm_strided(M, N, K, 0.5, A.numpy(), K, 1, Bfull.numpy(), N, 1, -1.25, want, N, 1)
    err = O.normwise_relative_error(C_local.numpy(), want[lo:hi]) if hi > lo else 0.0
    ret[rank] = (lo, hi, float(err))
    dist.barrier(); dist.destroy_process_group()


def test_partition_rows():
    from laser_b200.rowshard import partition_rows, partition_rows_c
    assert partition_rows(32768, 4) == [(0, 8192), (8192, 16384), (16384, 24576), (24576, 32768)]
    p = partition_rows(1000, 8, align=128)
    assert p[0] == (0, 128) and p[-1] == (896, 1000) and sum(b - a for a, b in p) == 1000
    p = partition_rows(100, 4)                      # fewer tile rows than ranks: trailing ranks idle
    assert p == [(0, 100), (100, 100), (100, 100), (100, 100)]
    for (M, w) in ((32768, 4), (32768, 8), (1000, 8), (100, 4), (8193, 2), (1, 3)):      # the library's own rule
        assert [partition_rows_c(M, w, r) for r in range(w)] == partition_rows(M, w)


@pytest.mark.timeout(300)
def test_rowsharded_gloo_world2():
    world, M, N, K = 2, 100, 48, 200
    port = _free_port()
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, M, N, K, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == [0, 1]
    assert ret[0][:2] == (0, 64) and ret[1][:2] == (64, 100)
    assert ret[0][2] == 0.0 and ret[1][2] == 0.0      # same kernel (the oracle), same rows: identical to the single call
