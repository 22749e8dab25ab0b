"""Batched descriptor kNN: query x gallery tcgen05 GEMM with a fused per-row top-k, gallery
sharded across ranks with a per-rank top-k followed by an allgather merge.
The reference has no retrieval code (SURVEY.md section 0 finding 4); this replaces
``torch.topk(Q @ G.T, k)``."""
from __future__ import annotations

import torch

from . import _lib as L


def knn_topk(q: torch.Tensor, g: torch.Tensor, k: int = 25, g_offset: int = 0):
    """q [Q, D] bf16, g [G, D] bf16 (this rank's shard).  Returns (score fp32 [Q,k] descending,
    idx int64 [Q,k] = g_offset + row; ties -> lower index; missing entries idx = -1)."""
    L.require_cuda(q, "queries")
    L.require_cuda(g, "gallery")
    if q.dtype != torch.bfloat16 or g.dtype != torch.bfloat16:
        raise RuntimeError("knn_topk takes bf16 descriptors (fp32 accumulation on the tensor cores)")
    q = q.contiguous()
    g = g.contiguous()
    Q, D = q.shape
    G = g.shape[0]
    if g.shape[1] != D:
        raise ValueError("query/gallery width mismatch")
    lib = L.lib()
    ws_bytes = lib.vpr_knn_workspace_bytes(Q, G, D, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    score = torch.empty(Q, k, dtype=torch.float32, device=q.device)
    idx = torch.empty(Q, k, dtype=torch.int64, device=q.device)
    with L.on_device(q):
        L.check(lib.vpr_knn_topk(L.ptr(q), L.ptr(g), Q, G, D, k, int(g_offset), L.ptr(score), L.ptr(idx), L.ptr(ws),
                                 ws_bytes, L.stream_ptr(q.device)), "vpr_knn_topk")
    return score, idx


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: int):
    """scores/idx [R, Q, k'] (per-rank lists) -> global (score [Q,k], idx [Q,k])."""
    R, Q, kk = scores.shape
    scores = scores.contiguous().float()
    idx = idx.contiguous().to(torch.int64)
    if kk != k:
        raise ValueError("per-rank lists must hold k entries")
    out_s = torch.empty(Q, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=scores.device)
    with L.on_device(scores):
        L.check(L.lib().vpr_knn_merge(L.ptr(scores), L.ptr(idx), R, Q, k, L.ptr(out_s), L.ptr(out_i),
                                      L.stream_ptr(scores.device)), "vpr_knn_merge")
    return out_s, out_i


def knn_topk_sharded(q: torch.Tensor, g_shard: torch.Tensor, k: int, g_offset: int, group=None):
    """Every rank holds all queries and a contiguous gallery shard starting at global row
    ``g_offset``.  Per-rank top-k, NCCL allgather of the [Q,k] (score, idx) lists, merge."""
    import torch.distributed as dist
    s, i = knn_topk(q, g_shard, k, g_offset)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return s, i
    R = dist.get_world_size(group)
    all_s = torch.empty(R, *s.shape, dtype=s.dtype, device=s.device)
    all_i = torch.empty(R, *i.shape, dtype=i.dtype, device=i.device)
    dist.all_gather_into_tensor(all_s, s, group=group)
    dist.all_gather_into_tensor(all_i, i, group=group)
    return merge_topk(all_s, all_i, k)
