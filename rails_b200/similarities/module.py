"""Plugin boundary ABC, mirroring rails/similarities/module.py:21-42 of the reference."""
import abc
from typing import Dict, Tuple

import torch


class SimilarityModule(torch.nn.Module):
    @abc.abstractmethod
    def forward(
        self,
        query_embeddings: torch.Tensor,
        item_embeddings: torch.Tensor,
        **kwargs,
    ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """
        Args:
            query_embeddings: (B, input_embedding_dim) x float.
            item_embeddings: (1/B, X, item_embedding_dim) x float.
        Returns:
            ((B, X) similarity values, keyed aux losses (empty at inference)).
        """
        pass
