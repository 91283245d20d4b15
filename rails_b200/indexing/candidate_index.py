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
