"""Exact brute-force MoL top-k, B200-native drop-in for the reference's
rails/indexing/mol_top_k.py: `MoLTopKModule` (:29-81) and `MoLBruteForceTopK` (:84-130).

Same constructor, same `forward(query_embeddings, k, sorted=True, **kwargs) -> (scores (B,k),
ids (B,k) int64)`.  What differs is where the work happens: the item side (projection, l2-norm,
item-only gating MLP) is computed ONCE into an on-device index at construction (the reference
recomputes it over the whole corpus on every call), and each forward() is one `mol_search` call of
libmol_b200.so: CUDA query prologue -> tcgen05 coarse scoring pass (fp16 operands, fp32 accumulate)
with the candidate filter fused into its epilogue -> the K' best candidates per query -> exact fp32
rescoring -> final sorted top-k.  A query whose candidate set fails the acceptance test (an empirical
margin between the coarse and the exact scores) is re-done by the exact fp32 kernel;
`mode=MODE_EXACT` scores every pair in fp32 and is the path with an unconditional guarantee.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from rails_b200 import _lib, engine
from rails_b200.indexing.candidate_index import TopKModule
from rails_b200.similarities.mol.similarity_fn import MoLSimilarity


class MoLTopKModule(TopKModule):
    def __init__(
        self,
        mol_module: MoLSimilarity,
        item_embeddings: torch.Tensor,
        item_ids: torch.Tensor,
        flatten_item_ids_and_embeddings: bool,
        keep_component_level_item_embeddings: bool,
        component_level_item_embeddings_dtype: torch.dtype = torch.bfloat16,
    ) -> None:
        """
        Args (as in the reference, mol_top_k.py:39-52):
            mol_module: MoLSimilarity.
            item_embeddings: (1, X, D) raw item embeddings on a CUDA device.
            item_ids: (1, X,) item ids.
        """
        super().__init__()
        self._mol_module: MoLSimilarity = mol_module
        self._item_embeddings: torch.Tensor = (
            item_embeddings if not flatten_item_ids_and_embeddings else item_embeddings.squeeze(0)
        )
        self._item_ids: torch.Tensor = item_ids if not flatten_item_ids_and_embeddings else item_ids.squeeze(0)
        self._index: Optional[engine.IndexHandle] = None
        self._index_key = None
        if keep_component_level_item_embeddings:
            self._mol_item_embeddings: torch.Tensor = self._ensure_index().xsub_f32().to(
                component_level_item_embeddings_dtype
            )

    @property
    def mol_module(self) -> MoLSimilarity:
        return self._mol_module

    def _flat_items(self) -> torch.Tensor:
        e = self._item_embeddings
        return e.squeeze(0) if e.dim() == 3 else e

    def _ensure_index(self) -> engine.IndexHandle:
        """(Re)builds the item-side cache when the weights or the item tensor changed."""
        items = self._flat_items()
        engine._require_cuda(items, "item_embeddings")
        weights = self._mol_module.packed_weights(items.device)
        key = (self._mol_module._packed_key, items.data_ptr(), items._version, tuple(items.shape))
        if self._index is None or self._index_key != key:
            self._index = engine.IndexHandle(weights, items, self._item_ids.reshape(-1))
            self._index_key = key
        return self._index


class MoLBruteForceTopK(MoLTopKModule):
    supports_invalid_ids = True  # forward(..., invalid_ids=(B, N0)) excludes ids inside the search

    def __init__(
        self,
        mol_module: MoLSimilarity,
        item_embeddings: torch.Tensor,
        item_ids: torch.Tensor,
        mode: int = _lib.MODE_AUTO,
        cuda_graph: bool = False,
    ) -> None:
        """mode / cuda_graph are extensions of the reference's constructor (mol_top_k.py:85-97).  cuda_graph=True replays
        each (B, k) signature of forward() as one captured CUDA graph (engine.GraphedSearch): the launch overhead of the
        ~25 kernels of a search disappears, which is most of the time of a small-batch call."""
        super().__init__(
            mol_module=mol_module,
            item_embeddings=item_embeddings,
            item_ids=item_ids,
            flatten_item_ids_and_embeddings=False,
            keep_component_level_item_embeddings=False,
        )
        self._mode = mode
        self._cuda_graph = bool(cuda_graph)
        self._graphs = {}
        if item_embeddings.is_cuda:
            self._ensure_index()

    def last_search_stats(self) -> dict:
        """Counters of the last forward() (engine.search_stats: exact-fallback queries, filter overflows, ...)."""
        ws = getattr(self, "_last_workspace", None)
        if ws is None:
            raise RuntimeError("forward() has not been called yet")
        return engine.search_stats(ws)

    @torch.no_grad()
    def forward(
        self,
        query_embeddings: torch.Tensor,
        k: int,
        sorted: bool = True,
        **kwargs,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """
        Args:
            query_embeddings: (B, D) x float on the index's CUDA device.
            k: int. final top-k to return (k <= X, as torch.topk requires; the caller clamps,
                indexing/candidate_index.py:149 of the reference).
            sorted: bool. Results are always returned sorted (descending), which satisfies both values.
            **kwargs: "user_ids" is consumed when uid embeddings are configured; "timestamps" / "ratings"
                (data/eval.py:148) are accepted and ignored.  "invalid_ids" ((B, N0) int64, an extension used by
                CandidateIndex.get_top_k_outputs): ids excluded per query INSIDE the search - the result is the top-k
                over the other items, which is what the reference gets from its k + N0 over-fetch and masking
                (indexing/candidate_index.py:144-178) whenever k + N0 <= X.
        Returns:
            Tuple of (top_k_scores x float, top_k_ids x int64), both of shape (B, K,)
        """
        index = self._ensure_index()
        dev = index.device
        weights = self._mol_module.packed_weights(dev)
        invalid_ids = kwargs.get("invalid_ids")
        if self._cuda_graph and query_embeddings.size(0) > 0 and invalid_ids is None:
            uid = kwargs.get("user_ids")
            key = (int(query_embeddings.size(0)), int(k), id(index), id(weights))
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) >= 8:  # (signatures of a serving loop are few; stale weights / indexes age out)
                    self._graphs.pop(next(iter(self._graphs)))
                g = engine.GraphedSearch(weights, index, key[0], key[1], self._mode, weights.shape.num_uid_tables > 0)
                self._graphs[key] = g
            engine._require_cuda(query_embeddings, "query_embeddings")
            scores, ids = g(query_embeddings.detach().to(device=dev, dtype=torch.float32), uid)
            self._last_workspace = g.workspace
            return scores.to(query_embeddings.dtype), ids
        self._last_workspace = self._mol_module.workspace(dev)
        scores, ids = engine.search(
            weights, index, self._mol_module.workspace(dev), query_embeddings, kwargs.get("user_ids"), int(k), sorted,
            self._mode, invalid_ids,
        )
        return scores.to(query_embeddings.dtype), ids


class MoLAvgTopK(MoLTopKModule):
    """Approximate MoL top-k of the reference (mol_top_k.py:296-429): a dot-product prefilter on the group-averaged
    sub-embeddings keeps `avg_top_k` items per query, exact MoL is evaluated on those and the best k are returned.
    Everything runs in `mol_search_avg`; the prefilter operand is kept in fp32 (the reference defaults to bf16
    component embeddings, mol_top_k.py:37), so the candidate set can only be more faithful to the fp32 dot products.
    SURVEY.md §8 row f3."""

    def __init__(self, mol_module: MoLSimilarity, item_embeddings: torch.Tensor, item_ids: torch.Tensor, avg_top_k: int) -> None:
        super().__init__(
            mol_module=mol_module,
            item_embeddings=item_embeddings,
            item_ids=item_ids,
            flatten_item_ids_and_embeddings=True,
            keep_component_level_item_embeddings=False,
        )
        self._avg_top_k: int = int(avg_top_k)
        self._avg_items = None
        self._avg_key = None

    def last_search_stats(self) -> dict:
        """Counters of the last forward(): `filter_strategy` = 1 when the streaming tcgen05 prefilter ran (mol_dotfilter),
        `fallback_queries` = prefilter rows re-done by the plain fp32 pass, `filter_overflows`, `max_survivors`."""
        return engine.search_stats(self._mol_module.workspace(self._ensure_index().device))

    def _ensure_avg(self):
        index = self._ensure_index()
        if self._avg_items is None or self._avg_key is not self._index:
            self._avg_items = engine.avg_item_embeddings(self._mol_module.packed_weights(index.device), index)
            self._avg_key = self._index
        return index, self._avg_items

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        if k > self._avg_top_k:
            raise ValueError(f"avg_top_k ({self._avg_top_k}) must be larger than k ({k})")
        index, avg = self._ensure_avg()
        dev = index.device
        scores, ids = engine.search_avg(
            self._mol_module.packed_weights(dev), index, avg, self._mol_module.workspace(dev), query_embeddings,
            kwargs.get("user_ids"), int(k), self._avg_top_k,
        )
        return scores.to(query_embeddings.dtype), ids


class MoLNaiveTopK(MoLTopKModule):
    """Greedy per-group top-k of the reference (mol_top_k.py:133-293): for each of the P_Q x P_X embedding-set pairs the
    `k_per_group` items with the largest dot product, the union scored with exact MoL.  As in the reference, `k` is
    ignored: all P_Q * P_X * k_per_group candidates come back sorted by score, duplicates carrying the -32767 sentinel
    (:256, :271-292).  One `mol_search_groups` call; the dot products are fp32 (the reference's default is bf16
    component embeddings, :37).  SURVEY.md §8 row f3."""

    def __init__(
        self,
        mol_module: MoLSimilarity,
        item_embeddings: torch.Tensor,
        item_ids: torch.Tensor,
        k_per_group: int,
        use_faiss: bool = False,
    ) -> None:
        if use_faiss:
            raise NotImplementedError("use_faiss=True: FAISS is outside the B200 hot path (DESIGN.md §7)")
        super().__init__(
            mol_module=mol_module,
            item_embeddings=item_embeddings,
            item_ids=item_ids,
            flatten_item_ids_and_embeddings=True,
            keep_component_level_item_embeddings=False,
        )
        self._k_per_group: int = int(k_per_group)
        self._use_faiss: bool = False

    def last_search_stats(self) -> dict:
        """Counters of the last forward() (see MoLAvgTopK.last_search_stats); rows = (query, query group, item group)."""
        return engine.search_stats(self._mol_module.workspace(self._ensure_index().device))

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        index = self._ensure_index()
        dev = index.device
        scores, ids = engine.search_groups(
            self._mol_module.packed_weights(dev), index, None, self._mol_module.workspace(dev), query_embeddings,
            kwargs.get("user_ids"), self._k_per_group, 0,
        )
        return scores.to(query_embeddings.dtype), ids


class MoLCombTopK(MoLAvgTopK):
    """MoLNaiveTopK's per-group candidates plus MoLAvgTopK's prefilter candidates (mol_top_k.py:432-551); returns all
    P_Q * P_X * k_per_group + avg_top_k candidates sorted by exact MoL score (:518, :533-551)."""

    def __init__(
        self,
        mol_module: MoLSimilarity,
        item_embeddings: torch.Tensor,
        item_ids: torch.Tensor,
        avg_top_k: int,
        k_per_group: int,
    ) -> None:
        super().__init__(mol_module, item_embeddings, item_ids, avg_top_k)
        self._k_per_group: int = int(k_per_group)

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, k: int, sorted: bool = True, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        index, avg = self._ensure_avg()
        dev = index.device
        scores, ids = engine.search_groups(
            self._mol_module.packed_weights(dev), index, avg, self._mol_module.workspace(dev), query_embeddings,
            kwargs.get("user_ids"), self._k_per_group, self._avg_top_k,
        )
        return scores.to(query_embeddings.dtype), ids
