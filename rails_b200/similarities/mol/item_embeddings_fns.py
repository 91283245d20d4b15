"""Item-side MoL embedding function (reference: rails/similarities/mol/item_embeddings_fns.py:122-183).

Parameter container; X_sub = l2norm(reshape(Linear(e_x), (P_X, d))) is computed once per corpus by
the CUDA index build (csrc/mol_api.cu: mol_index_build) instead of on every call.
"""
from typing import Callable

import torch

from rails_b200.similarities.mol.embeddings_fn import MoLEmbeddingsFn


class RecoMoLItemEmbeddingsFn(MoLEmbeddingsFn):
    def __init__(
        self,
        item_embedding_dim: int,
        item_dot_product_groups: int,
        dot_product_dimension: int,
        dot_product_l2_norm: bool,
        proj_fn: Callable[[int, int], torch.nn.Module],
        eps: float,
    ) -> None:
        super().__init__()
        self._item_emb_based_dot_product_groups: int = item_dot_product_groups
        self._item_emb_proj_module: torch.nn.Module = proj_fn(
            item_embedding_dim, dot_product_dimension * self._item_emb_based_dot_product_groups
        )
        self._dot_product_dimension: int = dot_product_dimension
        self._dot_product_l2_norm: bool = dot_product_l2_norm
        self._eps: float = eps

    def forward(self, input_embeddings: torch.Tensor, **kwargs):  # pragma: no cover
        raise RuntimeError("call MoLSimilarity.get_item_component_embeddings (CUDA index build)")
