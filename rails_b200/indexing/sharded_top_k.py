"""Corpus-sharded exact MoL top-k over several GPUs (one process per GPU).

The reference evaluates on a single GPU only (eval_from_checkpoint.py:555 asserts world_size == 1); the
north star adds corpus sharding: rank r owns the contiguous item range [r*N/R, (r+1)*N/R), runs the
same `MoLBruteForceTopK` search on its shard, and ONE all-gather of the per-shard (B, k) scores + ids
followed by a (R*k -> k) merge gives the global answer on every rank.  The union of exact per-shard
top-k lists contains the global top-k, so the result equals the unsharded one (up to exact ties).

The exchange is the only collective on the data path.  `local_search` and `merge` are injectable so the
host logic (ranges, gather layout, merge call) is covered by world_size-2 `gloo` tests on CPU; in
production they default to the CUDA engine (no CPU fallback exists for them).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous item range [lo, hi) of `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return rank * num_items // world_size, (rank + 1) * num_items // world_size


def pack_partials(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """(B, k) fp32 scores + (B, k) int64 ids -> one (B, k, 3) int32 buffer, so the exchange is ONE collective."""
    if scores.shape != ids.shape or scores.dim() != 2:
        raise ValueError("scores / ids must both be (B, k)")
    buf = torch.empty(scores.shape + (3,), dtype=torch.int32, device=scores.device)
    buf[..., 0] = scores.to(torch.float32).contiguous().view(torch.int32)
    buf[..., 1:] = ids.to(torch.int64).contiguous().view(torch.int32).view(scores.shape + (2,))
    return buf


def unpack_partials(buf: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse of pack_partials for a (..., B, k, 3) int32 buffer."""
    scores = buf[..., 0].contiguous().view(torch.float32)
    ids = buf[..., 1:].contiguous().view(torch.int64).squeeze(-1)
    return scores, ids


def _engine_merge(part_scores: torch.Tensor, part_ids: torch.Tensor, k: int):
    from rails_b200 import engine

    return engine.merge_topk(part_scores, part_ids, k)


class ShardedMoLBruteForceTopK(torch.nn.Module):
    """Wraps the rank-local top-k module of one corpus shard.

    Args:
        local_top_k: callable (query_embeddings, k, **kwargs) -> (scores (B, k'), ids (B, k') int64 GLOBAL ids)
            over this rank's shard — normally `MoLBruteForceTopK(mol, items[lo:hi], ids[lo:hi])`.
        shard_items: number of items in this rank's shard (k is clamped to it locally; shards with fewer
            than k items pad with (-inf, -1)).
        group: process group (default: WORLD).  With world_size 1 no collective is issued.
        merge: (R, B, k) scores, (R, B, k) ids, k -> (B, k) scores, (B, k) ids; defaults to the CUDA merge.
    """

    def __init__(
        self,
        local_top_k: Callable[..., Tuple[torch.Tensor, torch.Tensor]],
        shard_items: int,
        group: Optional[dist.ProcessGroup] = None,
        merge: Optional[Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]]] = None,
    ) -> None:
        super().__init__()
        self._local = local_top_k
        self._shard_items = int(shard_items)
        self._group = group
        self._merge = merge or _engine_merge

    def world_size(self) -> int:
        return dist.get_world_size(self._group) if dist.is_available() and dist.is_initialized() else 1

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs):
        R = self.world_size()
        k_local = min(int(k), self._shard_items)
        B = query_embeddings.size(0)
        dev = query_embeddings.device
        if k_local > 0:
            s, i = self._local(query_embeddings, k_local, sorted=True, **kwargs)
            s = s.to(torch.float32)
        else:
            s = torch.empty((B, 0), dtype=torch.float32, device=dev)
            i = torch.empty((B, 0), dtype=torch.int64, device=dev)
        if k_local < k:  # pad so every rank contributes the same (B, k) block
            pad = k - k_local
            s = torch.cat([s, torch.full((B, pad), float("-inf"), dtype=torch.float32, device=dev)], dim=1)
            i = torch.cat([i, torch.full((B, pad), -1, dtype=torch.int64, device=dev)], dim=1)
        if R == 1:
            if k_local < k:
                raise RuntimeError(f"selected index k out of range (k={k} > {self._shard_items} items)")
            return s.to(query_embeddings.dtype), i
        mine = pack_partials(s, i)
        gathered = torch.empty((R * B,) + tuple(mine.shape[1:]), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=self._group)  # the single exchange step
        ps, pi = unpack_partials(gathered.view((R,) + tuple(mine.shape)))
        total = int((pi[:, 0, :] >= 0).sum().item()) if k_local < k else R * k
        if total < k:
            raise RuntimeError(f"selected index k out of range (k={k} > {total} items in all shards)")
        ms, mi = self._merge(ps, pi, int(k))
        return ms.to(query_embeddings.dtype), mi
