"""Exact MIPS top-k, B200-native drop-in for rails/indexing/mips_top_k.py (`MIPSTopKModule` :24-39,
`MIPSBruteForceTopK` :41-81): fp32 q . items^T over the whole corpus + row-wise top-k + id gather, all in
libmol_b200.so (`mol_mips_search`).  Large corpora (>= 64k items, D a multiple of 32) take the streaming path: a tcgen05
tf32 pass with a fused threshold filter, fp32 rescoring of the survivors and a per-query completeness proof - the fused
GEMM + top-k of SURVEY.md §8 row f4; the returned scores and ids are those of the fp32 computation."""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch

from rails_b200 import _lib, engine
from rails_b200.indexing.candidate_index import TopKModule


class MIPSTopKModule(TopKModule):
    def __init__(self, item_embeddings: torch.Tensor, item_ids: torch.Tensor) -> None:
        """
        Args:
            item_embeddings: (1, X, D)
            item_ids: (1, X,)
        """
        super().__init__()
        self._item_embeddings: torch.Tensor = item_embeddings
        self._item_ids: torch.Tensor = item_ids


class MIPSBruteForceTopK(MIPSTopKModule):
    def __init__(self, item_embeddings: torch.Tensor, item_ids: torch.Tensor) -> None:
        super().__init__(item_embeddings=item_embeddings, item_ids=item_ids)
        # the C ABI wants (X, D) row-major fp32 (the reference keeps the (D, X) transpose for torch.mm)
        self._items = None
        self._ids = None
        self._key = None
        self._ws = None
        self._norm_cache = None  # {bound on the item norms or -1, accumulator}: mol_mips_search_cached
        if item_embeddings.is_cuda:
            self._sync()

    def _sync(self) -> None:
        """(X, D) fp32 items and int64 ids on the items' device, rebuilt when the source tensors change in place (the
        reference reads `self._item_embeddings` on every call).  k is limited to MOL_MAX_K = 8192 by the select kernels
        (torch.topk accepts any k <= X)."""
        e, i = self._item_embeddings, self._item_ids
        key = (e.data_ptr(), e._version, tuple(e.shape), i.data_ptr(), i._version)
        if self._key != key:
            self._items = e.squeeze(0).detach().to(torch.float32).contiguous()
            self._ids = i.reshape(-1).detach().to(device=self._items.device, dtype=torch.int64).contiguous()
            self._key = key
            self._norm_cache = torch.tensor([-1.0, 0.0], dtype=torch.float32, device=self._items.device)

    def last_search_stats(self) -> dict:
        """Counters of the last forward(): `filter_strategy` = 1 when the streaming tcgen05 path ran (no (B, N) matrix),
        `fallback_queries` = queries re-done by the plain fp32 pass, `filter_overflows`, `max_survivors`."""
        if self._ws is None:
            raise RuntimeError("forward() has not been called yet")
        lib = _lib.load()
        out = (ctypes.c_int32 * _lib.NUM_STATS)()
        dev = self._ws.device
        with torch.cuda.device(dev):
            _lib.check(lib.mol_search_stats(engine._ptr(self._ws), out, engine._stream_ptr(dev)))
        return dict(zip(_lib.STAT_NAMES, (int(v) for v in out)))

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B, D) queries -> (top_k_scores (B, k), top_k_ids (B, k) int64); results are always sorted."""
        lib = _lib.load()
        engine._require_cuda(query_embeddings, "query_embeddings")
        engine._require_cuda(self._item_embeddings, "item_embeddings")
        self._sync()
        dev = self._items.device
        q = query_embeddings.detach().to(device=dev, dtype=torch.float32).contiguous()
        B, D = int(q.size(0)), int(q.size(1))
        if D != self._items.size(1):
            raise ValueError(f"query dim {D} != item dim {self._items.size(1)}")
        N = int(self._items.size(0))
        out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
        nbytes = ctypes.c_size_t()
        _lib.check(lib.mol_mips_workspace_bytes(N, B, k, ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(
                lib.mol_mips_search_cached(
                    engine._ptr(self._items), engine._ptr(self._ids), engine._ptr(q), N, D, B, int(k),
                    engine._ptr(self._norm_cache), engine._ptr(out_s), engine._ptr(out_i), engine._ptr(self._ws), self._ws.numel(),
                    engine._stream_ptr(dev),
                )
            )
        return out_s.to(query_embeddings.dtype), out_i
