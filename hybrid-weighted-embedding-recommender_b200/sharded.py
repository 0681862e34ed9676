"""Item-sharded top-k across the GPUs of one node (no reference counterpart -- the
reference is single-process).

GPU g owns the contiguous catalogue rows [g*N/G, (g+1)*N/G); queries are replicated; every rank computes its
local exact top-k (rows offset to global ids, fp64 scores), one all-gather exchanges the [B, k] results, and the
merge kernel applies the same (score desc, row asc) rule -- so 1, 2, 4 and 8-GPU runs return identical results.
"""
from typing import Tuple

import torch
import torch.distributed as dist

from . import ops


def partition(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range owned by `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def pad_local_result(idx, s64, k):
    """Local shards smaller than k return fewer columns; pad with (row -1, score -inf)."""
    B, kl = idx.shape
    if kl == k:
        return idx, s64
    pi = torch.full((B, k), -1, dtype=idx.dtype, device=idx.device)
    ps = torch.full((B, k), float("-inf"), dtype=s64.dtype, device=s64.device)
    pi[:, :kl] = idx
    ps[:, :kl] = s64
    return pi, ps


def gather_shard_results(idx, s64, group=None):
    """One collective for both arrays: [B, k] per rank -> ([G, B, k] fp64 scores, [G, B, k] int64 rows)."""
    world = dist.get_world_size(group)
    packed = torch.stack([s64.view(torch.int64), idx], dim=0).contiguous()      # [2, B, k] int64
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    allp = torch.stack(out, dim=0)                                               # [G, 2, B, k]
    return allp[:, 0].contiguous().view(torch.float64), allp[:, 1].contiguous()


class ShardedTopK:
    def __init__(self, local_table, row_offset: int, shadow=None, group=None, max_norm=None):
        self.group = group
        self.row_offset = int(row_offset)
        self.index = ops.TopKIndex(local_table, shadow, max_norm=max_norm)

    def local_topk(self, queries, k, mode="exact"):
        kl = min(int(k), self.index.n)
        idx, _, s64 = self.index.topk(queries, kl, mode, idx_offset=self.row_offset, want_f64=True)
        return pad_local_result(idx, s64, int(k))

    def topk(self, queries, k, mode="exact"):
        idx, s64 = self.local_topk(queries, k, mode)
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return idx, s64.float()
        gs, gi = gather_shard_results(idx, s64, self.group)
        return ops.merge_topk(gs, gi)
