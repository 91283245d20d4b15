"""DotProductSimilarity, mirroring rails/similarities/dot_product_similarity_fn.py:24-68 ((1, X, D) branch on the
GPU through `mol_dot_scores`; the per-row (B, X, D) branches are training-time paths, out of scope)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from rails_b200 import _lib, engine
from rails_b200.similarities.module import SimilarityModule


class DotProductSimilarity(SimilarityModule):
    def __init__(self) -> None:
        super().__init__()

    def debug_str(self) -> str:
        return "dp"

    @torch.no_grad()
    def forward(self, query_embeddings: torch.Tensor, item_embeddings: torch.Tensor, **kwargs) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """query_embeddings (B, D), item_embeddings (1, X, D) -> ((B, X), {})."""
        if item_embeddings.dim() != 3 or item_embeddings.size(0) != 1:
            raise NotImplementedError("only the (1, X, D) shared-candidates branch is implemented")
        engine._require_cuda(query_embeddings, "query_embeddings")
        dev = query_embeddings.device
        items = item_embeddings.squeeze(0).detach().to(device=dev, dtype=torch.float32).contiguous()
        q = query_embeddings.detach().to(torch.float32).contiguous()
        out = torch.empty((q.size(0), items.size(0)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(
                _lib.load().mol_dot_scores(
                    engine._ptr(items), engine._ptr(q), items.size(0), items.size(1), q.size(0), engine._ptr(out),
                    engine._stream_ptr(dev),
                )
            )
        return out.to(query_embeddings.dtype), {}
