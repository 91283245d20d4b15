"""Query-side MoL embedding function (reference: rails/similarities/mol/query_embeddings_fns.py:129-254).

Parameter container with the reference's attribute names; the computation
    Q_sub = l2norm(cat(reshape(Linear(GLU(q))), UidEmb[(uid % hash) + 1]))
is executed by the CUDA query prologue (csrc/mol_prologue.cu) through the owning MoLSimilarity.
"""
from typing import Callable, List

import torch

from rails_b200.similarities.mol.embeddings_fn import MoLEmbeddingsFn


class RecoMoLQueryEmbeddingsFn(MoLEmbeddingsFn):
    def __init__(
        self,
        query_embedding_dim: int,
        query_dot_product_groups: int,
        dot_product_dimension: int,
        dot_product_l2_norm: bool,
        proj_fn: Callable[[int, int], torch.nn.Module],
        eps: float,
        uid_embedding_hash_sizes: List[int],
        uid_dropout_rate: float,
        uid_embedding_level_dropout: bool = False,
    ) -> None:
        super().__init__()
        self._uid_embedding_hash_sizes: List[int] = list(uid_embedding_hash_sizes)
        self._query_emb_based_dot_product_groups: int = query_dot_product_groups - len(self._uid_embedding_hash_sizes)
        self._query_emb_proj_module: torch.nn.Module = proj_fn(
            query_embedding_dim, dot_product_dimension * self._query_emb_based_dot_product_groups
        )
        self._dot_product_dimension: int = dot_product_dimension
        self._dot_product_l2_norm: bool = dot_product_l2_norm
        for i, hash_size in enumerate(self._uid_embedding_hash_sizes):
            setattr(self, f"_uid_embeddings_{i}", torch.nn.Embedding(hash_size + 1, dot_product_dimension, padding_idx=0))
        self._uid_dropout_rate: float = uid_dropout_rate
        self._uid_embedding_level_dropout: bool = uid_embedding_level_dropout
        self._eps: float = eps

    def forward(self, input_embeddings: torch.Tensor, **kwargs):  # pragma: no cover
        raise RuntimeError("call MoLSimilarity.get_query_component_embeddings (CUDA query prologue)")
