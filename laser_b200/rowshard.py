"""Row-panel sharding of a large GEMM across the GPUs of one box (one process per GPU).

The reference parallelises exactly this way on a CPU: its `ic` loop hands each worker a
row block of A and C while all workers share one packed panel of B per `pc` iteration, and
beta is applied on the first K-panel only (gemm.nim:150-176).  Here:
  * rank r owns rows [start_r, stop_r) of A and C (never moved);
  * B lives on `src` and is broadcast once over NCCL/NVLink (torch.distributed) -- no
    collective inside the MMA loop, no reduction.  With n_panels > 1 the broadcast is cut in
    K-panels and panel i+1 is in flight while the GEMM consumes panel i with beta' = beta
    (i == 0) or 1 (i > 0), the reference's own rule for its K blocks.  Measured on 2 and 4
    B200s (tools/rowshard_probe.py): one panel is fastest -- a 256 MB broadcast costs 0.45 ms
    against a 3.6 ms GEMM, while every extra K-panel costs a C read-modify-write pass and a
    kernel ramp -- so n_panels defaults to 1.
  * overlap_prepack (LASER_B200_ROWSHARD_OVERLAP=1; off by default until measured): the
    rank's own rows of A are prepared (gemm_prepackA) on a side stream while the broadcast of
    B is in flight, so only B's preparation and the product follow the collective.
The host logic is backend-agnostic (tested on CPU with gloo + the oracle as gemm_fn).
"""
import os

import torch
import torch.distributed as dist

__all__ = ["partition_rows", "k_panels", "gemm_rowsharded"]


def partition_rows(M, world_size, align=128):
    """[(start, stop)] per rank: ceil(M / world) rounded up to the CTA tile height; the
    last ranks take what is left (possibly nothing)."""
    per = -(-M // world_size)
    per = -(-per // align) * align
    out = []
    for r in range(world_size):
        lo = min(M, r * per)
        out.append((lo, min(M, lo + per)))
    return out


def k_panels(K, n_panels, align=32):
    """Split K in at most n_panels panels whose sizes are multiples of `align` (one TMA
    k-block) except possibly the last."""
    n_panels = max(1, min(int(n_panels), -(-K // align)))
    per = -(-K // n_panels)
    per = -(-per // align) * align
    out, k0 = [], 0
    while k0 < K:
        out.append((k0, min(K, k0 + per)))
        k0 += per
    return out


_packed_cache = {}   # (device, M, N, K) -> (packedA, packedB, side stream)


def _packed_buffers(device, M, N, K):
    from . import prepacked as pp
    key = (str(device), M, N, K)
    if key not in _packed_cache:
        _packed_cache.clear()          # one shape at a time: the buffers are as large as the operands
        _packed_cache[key] = (pp.alloc_packed(pp.gemm_prepackA_mem_required(M, N, K)),
                              pp.alloc_packed(pp.gemm_prepackB_mem_required(M, N, K)),
                              torch.cuda.Stream(device=device))
    return _packed_cache[key]


def _overlap_applicable(A_local, B, C_local, n_panels, M_local):
    from .gemm import get_f32_mode
    return (n_panels == 1 and M_local > 0 and A_local.is_cuda and A_local.dtype == torch.float32
            and B.dtype == torch.float32 and C_local.dtype == torch.float32 and get_f32_mode() in (0, 5))


def gemm_rowsharded(M_local, N, K, alpha, A_local, B, beta, C_local, src=0, group=None,
                    n_panels=None, gemm_fn=None, broadcast=True, overlap_prepack=None):
    """C_local <- alpha * A_local @ B + beta * C_local on every rank.

    A_local: (M_local, K) tensor view (any strides), C_local: (M_local, N) view,
    B: (K, N) row-major contiguous tensor on every rank, valid on `src` only (unless
    broadcast=False).  gemm_fn has the gemm_strided signature (default: the CUDA library)."""
    if n_panels is None:     # default 1 (measured best in round 1); env override for A/B runs of bench.py
        n_panels = int(os.environ.get("LASER_B200_ROWSHARD_PANELS", "1"))
    if overlap_prepack is None:
        overlap_prepack = os.environ.get("LASER_B200_ROWSHARD_OVERLAP", "0") == "1"
    overlap = bool(overlap_prepack) and gemm_fn is None and _overlap_applicable(A_local, B, C_local, n_panels, M_local)
    if gemm_fn is None:
        from .gemm import gemm_strided as gemm_fn
    assert B.dim() == 2 and B.shape[0] == K and B.shape[1] == N and B.is_contiguous()
    panels = k_panels(K, n_panels)
    works = []
    if broadcast and dist.is_initialized() and dist.get_world_size(group) > 1:
        for (k0, k1) in panels:  # all panels are queued now; NCCL streams them in order
            works.append(dist.broadcast(B[k0:k1], src=src, group=group, async_op=True))
    rsA, csA = (A_local.stride(0), A_local.stride(1)) if M_local > 0 else (K, 1)
    rsC, csC = (C_local.stride(0), C_local.stride(1)) if M_local > 0 else (N, 1)
    if overlap:
        from . import prepacked as pp
        pa, pb, side = _packed_buffers(A_local.device, M_local, N, K)
        cur = torch.cuda.current_stream(A_local.device)
        side.wait_stream(cur)                      # A (and the packed buffers' last readers) are ordered on cur
        with torch.cuda.stream(side):
            pp.gemm_prepackA(pa, M_local, N, K, A_local, rsA, csA)
        if works:
            works[0].wait()                        # cur waits for B
        pp.gemm_prepackB(pb, M_local, N, K, B, N, 1)
        cur.wait_stream(side)
        pp.gemm_packed(M_local, N, K, alpha, pa, pb, beta, C_local, rsC, csC)
        return C_local
    for i, (k0, k1) in enumerate(panels):
        if works:
            works[i].wait()  # NCCL: the current stream waits for panel i; gloo: host wait
        if M_local == 0:
            continue
        gemm_fn(M_local, N, k1 - k0, alpha, A_local[:, k0:k1], rsA, csA, B[k0:k1], N, 1,
                beta if i == 0 else 1.0, C_local, rsC, csC)
    return C_local
