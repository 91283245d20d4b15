"""Exact MoL top-k over several GPUs (one process per GPU): corpus-sharded and query-split (replicated) modes.

The reference evaluates on a single GPU only (eval_from_checkpoint.py:555 asserts world_size == 1); the
north star adds multi-GPU serving with ONE collective on the data path:

* `ShardedMoLBruteForceTopK` — the corpus is sharded: rank r owns the contiguous item range
  [r*N/R, (r+1)*N/R), runs the same `MoLBruteForceTopK` search on its shard, packs its (B, k) partial list
  into one buffer (`mol_pack_topk`), ONE all-gather moves the R lists to every rank and `mol_merge_topk_packed`
  reduces R*k candidates per query to the global top-k.  The union of exact per-shard top-k lists contains the
  global top-k, so the result equals the unsharded one (up to exact ties).  Use it when the index does not fit one
  GPU (BASELINE.json configs 4 and 5) — its fixed per-step work (query prologue, selections, rescoring of K'
  candidates per query) is repeated on every rank.
* `ReplicatedMoLBruteForceTopK` — every rank holds the whole corpus and searches its slice of the QUERY batch;
  ONE all-gather of the packed (B/R, k) results gives every rank the full (B, k) answer.  No merge, and the fixed
  per-step work shrinks with the slice: the better choice whenever the index fits one GPU (the 1M-item north-star
  corpus is 2 GB).

`local_top_k`, `pack` and `merge_packed` are injectable so the host logic (ranges, padding, gather layout) is
covered by world_size-2 `gloo` tests on CPU; in production they default to the CUDA engine (no CPU fallback).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

PACKED_ENTRY_BYTES = 16  # include/mol_b200.h: MOL_PACKED_ENTRY_BYTES {int64 id, float score, int32 valid}

LocalTopK = Callable[..., Tuple[torch.Tensor, torch.Tensor]]
PackFn = Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor]
MergeFn = Callable[[torch.Tensor, int, int, int], Tuple[torch.Tensor, torch.Tensor]]


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of `rank` out of `num_items` units (items or queries)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return rank * num_items // world_size, (rank + 1) * num_items // world_size


def _engine_pack(scores: torch.Tensor, ids: torch.Tensor, k: int) -> torch.Tensor:
    from rails_b200 import engine

    return engine.pack_topk(scores, ids, k)


def _engine_merge_packed(gathered: torch.Tensor, R: int, B: int, k: int):
    from rails_b200 import engine

    return engine.merge_topk_packed(gathered, R, B, k)


class _MultiGpuTopK(torch.nn.Module):
    def __init__(self, local_top_k: LocalTopK, group, pack: Optional[PackFn], merge_packed: Optional[MergeFn]) -> None:
        super().__init__()
        self._local = local_top_k
        self._group = group
        self._pack = pack or _engine_pack
        self._merge_packed = merge_packed or _engine_merge_packed

    def world_size(self) -> int:
        return dist.get_world_size(self._group) if dist.is_available() and dist.is_initialized() else 1

    def rank(self) -> int:
        return dist.get_rank(self._group) if dist.is_available() and dist.is_initialized() else 0

    def _all_gather(self, mine: torch.Tensor) -> torch.Tensor:
        """The single exchange step: (rows, k, 16) uint8 per rank -> (R, rows, k, 16)."""
        R = self.world_size()
        gathered = torch.empty((R,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(gathered.view(-1), mine.reshape(-1), group=self._group)
        return gathered


class ShardedMoLBruteForceTopK(_MultiGpuTopK):
    """Wraps the rank-local top-k module of one corpus shard.

    Args:
        local_top_k: callable (query_embeddings, k, **kwargs) -> (scores (B, k'), ids (B, k') int64 GLOBAL ids)
            over this rank's shard — normally `MoLBruteForceTopK(mol, items[lo:hi], ids[lo:hi])`.
        shard_items: number of items in this rank's shard.  The shard sizes are exchanged ONCE (first forward), so
            every rank validates k against the whole corpus identically; a shard with fewer than k items contributes
            its whole list and marks the rest of its (B, k) block invalid.
        group: process group (default: WORLD).  With world_size 1 no collective is issued.
        pack / merge_packed: default to the CUDA engine (mol_pack_topk / mol_merge_topk_packed).
    """

    def __init__(
        self,
        local_top_k: LocalTopK,
        shard_items: int,
        group: Optional[dist.ProcessGroup] = None,
        pack: Optional[PackFn] = None,
        merge_packed: Optional[MergeFn] = None,
    ) -> None:
        super().__init__(local_top_k, group, pack, merge_packed)
        self._shard_items = int(shard_items)
        self._total_items: Optional[int] = None

    def total_items(self, device: torch.device) -> int:
        if self._total_items is None:
            if self.world_size() == 1:
                self._total_items = self._shard_items
            else:
                t = torch.tensor([self._shard_items], dtype=torch.int64, device=device)
                dist.all_reduce(t, group=self._group)
                self._total_items = int(t.item())  # (once, not on the per-call path)
        return self._total_items

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs):
        R = self.world_size()
        k = int(k)
        dev = query_embeddings.device
        total = self.total_items(dev)
        if k > total:  # torch.topk's error, raised on EVERY rank before any work
            raise RuntimeError(f"selected index k out of range (k={k} > {total} items in all shards)")
        k_local = min(k, self._shard_items)
        B = query_embeddings.size(0)
        if k_local > 0:
            s, i = self._local(query_embeddings, k_local, sorted=True, **kwargs)
            s = s.to(torch.float32)
        else:
            s = torch.empty((B, 0), dtype=torch.float32, device=dev)
            i = torch.empty((B, 0), dtype=torch.int64, device=dev)
        if R == 1:
            return s.to(query_embeddings.dtype), i
        gathered = self._all_gather(self._pack(s, i, k))
        ms, mi = self._merge_packed(gathered, R, B, k)
        return ms.to(query_embeddings.dtype), mi


class ReplicatedMoLBruteForceTopK(_MultiGpuTopK):
    """Query-split search over a corpus replicated on every rank.

    Args:
        local_top_k: callable (query_embeddings, k, **kwargs) -> (scores, ids) over the WHOLE corpus, e.g.
            `MoLBruteForceTopK(mol, items, ids)` built on every rank from the same items.
        Every rank passes the same (B, D) query batch (and the same per-query kwargs such as user_ids); rank r
        searches rows shard_range(B, r, R) and all ranks return the full (B, k) result.
    """

    def __init__(
        self,
        local_top_k: LocalTopK,
        group: Optional[dist.ProcessGroup] = None,
        pack: Optional[PackFn] = None,
        merge_packed: Optional[MergeFn] = None,
    ) -> None:
        super().__init__(local_top_k, group, pack, merge_packed)

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs):
        R, r = self.world_size(), self.rank()
        k = int(k)
        if R == 1:
            return self._local(query_embeddings, k, sorted=True, **kwargs)
        B = query_embeddings.size(0)
        dev = query_embeddings.device
        lo, hi = shard_range(B, r, R)
        rows = (B + R - 1) // R  # every rank contributes a (rows, k) block; the last row of a short slice is padding
        kw = {n: (v[lo:hi] if torch.is_tensor(v) and v.dim() >= 1 and v.size(0) == B else v) for n, v in kwargs.items()}
        if hi > lo:
            s, i = self._local(query_embeddings[lo:hi], k, sorted=True, **kw)
            s = s.to(torch.float32)
        else:
            s = torch.empty((0, k), dtype=torch.float32, device=dev)
            i = torch.empty((0, k), dtype=torch.int64, device=dev)
        if hi - lo < rows:
            pad = rows - (hi - lo)
            s = torch.cat([s, torch.zeros((pad, k), dtype=torch.float32, device=dev)])
            i = torch.cat([i, torch.zeros((pad, k), dtype=torch.int64, device=dev)])
        gathered = self._all_gather(self._pack(s, i, k))  # (R, rows, k, 16)
        # every row is already a sorted top-k list: a one-part "merge" unpacks it
        us, ui = self._merge_packed(gathered.view(1, R * rows, k, PACKED_ENTRY_BYTES), 1, R * rows, k)
        if rows * R != B:
            keep = torch.cat([torch.arange(q * rows, q * rows + (shard_range(B, q, R)[1] - shard_range(B, q, R)[0]))
                              for q in range(R)]).to(dev)
            us, ui = us[keep], ui[keep]
        return us.to(query_embeddings.dtype), ui
