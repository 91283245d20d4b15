"""ABC for MoL embedding functions (reference: rails/similarities/mol/embeddings_fn.py:56-78)."""
import abc
from typing import Dict, Tuple

import torch


class MoLEmbeddingsFn(torch.nn.Module):
    @abc.abstractmethod
    def forward(self, input_embeddings: torch.Tensor, **kwargs) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """(B, ...) -> ((B, P, d) l2-normalised component embeddings, aux losses)."""
        pass
