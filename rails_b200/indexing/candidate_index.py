"""TopKModule ABC, mirroring rails/indexing/candidate_index.py:24-42 of the reference."""
import abc
from typing import Tuple

import torch


class TopKModule(torch.nn.Module):
    @abc.abstractmethod
    def forward(
        self,
        query_embeddings: torch.Tensor,
        k: int,
        sorted: bool = True,
        **kwargs,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """
        Args:
            query_embeddings: (B, X, ...). Implementation-specific.
            k: int. top k to return.
            sorted: bool.
        Returns:
            Tuple of (top_k_scores, top_k_ids), both of shape (B, K,)
        """
        pass


class CandidateIndex(object):
    """Mirror of the reference's eval-side `CandidateIndex` (indexing/candidate_index.py:31-185), restricted to what
    the brute-force eval path uses: `ids`, `embeddings`, `num_objects` and `get_top_k_outputs`.  The seen-item
    masking + back-fill (:155-178) never materialises a boolean (B, k', N0) tensor and never syncs the host
    (`torch.nonzero`).  With a top-k module that takes the exclusion list itself (`supports_invalid_ids`,
    MoLBruteForceTopK) and k + N0 <= X, the seen ids are excluded INSIDE the search (`mol_search_excluding`): no k' = k + N0
    over-fetch, the result is the top-k over the unseen items directly.  Otherwise the over-fetched list goes through
    one CUDA kernel (`mol_select_valid`).  SURVEY.md §8 row f2."""

    def __init__(self, ids: torch.Tensor, embeddings: torch.Tensor, invalid_ids=None, debug_path=None) -> None:
        super().__init__()
        self._ids = ids
        self._embeddings = embeddings
        self._invalid_ids = invalid_ids
        self._debug_path = debug_path

    @property
    def ids(self) -> torch.Tensor:
        return self._ids

    @property
    def num_objects(self) -> int:
        return self._ids.size(1)

    @property
    def embeddings(self) -> torch.Tensor:
        return self._embeddings

    @torch.no_grad()
    def get_top_k_outputs(
        self,
        query_embeddings: torch.Tensor,
        k: int,
        aux_payloads,
        top_k_module: TopKModule,
        invalid_ids,
        r: int = 1,
        return_embeddings: bool = False,
        truncate_k_prime_to=None,
    ):
        """Returns (top_k_ids (B, k), top_k_scores (B, k), None) — argument meaning as in the reference (:116-147)."""
        import ctypes

        from rails_b200 import _lib

        if return_embeddings:
            raise NotImplementedError("return_embeddings=True is broken in the reference (:182) and not implemented")
        max_num_invalid_ids = 0 if invalid_ids is None else invalid_ids.size(1)
        if (
            max_num_invalid_ids > 0
            and truncate_k_prime_to is None
            and getattr(top_k_module, "supports_invalid_ids", False)
            and k + max_num_invalid_ids <= min(self.num_objects, _lib.MOL_MAX_K)
            and query_embeddings.is_cuda
        ):
            # at most N0 of the best k + N0 items are excluded, so "the first k unseen entries of the top-(k + N0)" (the
            # reference's result, :155-178, no back-fill needed) IS the top-k over the unseen items
            scores, ids = top_k_module(query_embeddings=query_embeddings, k=k, invalid_ids=invalid_ids, **aux_payloads)
            return ids, scores, None
        k_prime = min(k + max_num_invalid_ids, self.num_objects)
        if truncate_k_prime_to is not None:
            k_prime = min(k_prime, truncate_k_prime_to)
        scores, ids = top_k_module(query_embeddings=query_embeddings, k=k_prime, **aux_payloads)
        if invalid_ids is None:
            return ids, scores, None
        if k_prime < k:
            raise RuntimeError(f"cannot select {k} of {k_prime} candidates")
        if not scores.is_cuda:
            raise RuntimeError("CandidateIndex.get_top_k_outputs needs CUDA tensors (no CPU fallback)")
        dev = scores.device
        s32 = scores.detach().to(torch.float32).contiguous()
        i64 = ids.detach().to(torch.int64).contiguous()
        inv = invalid_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
        B = s32.size(0)
        out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(
                _lib.load().mol_select_valid(
                    ctypes.c_void_p(s32.data_ptr()), ctypes.c_void_p(i64.data_ptr()), ctypes.c_void_p(inv.data_ptr()),
                    B, k_prime, inv.size(1), k, ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream),
                )
            )
        return out_i, out_s.to(scores.dtype), None
