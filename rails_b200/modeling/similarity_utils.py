"""Factory with the reference's keyword surface (modeling/similarity_utils.py:41-245).

`create_mol_interaction_module(**gin_bindings)` returns `(MoLSimilarity, debug_str)`; the debug
string is the one the reference uses to name checkpoints, and the module's state-dict keys equal the
reference's, so `load_state_dict(ckpt, strict=True)` works on reference checkpoints.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from rails_b200.similarities.layers import GeGLU, SwiGLU
from rails_b200.similarities.mol.item_embeddings_fns import RecoMoLItemEmbeddingsFn
from rails_b200.similarities.mol.query_embeddings_fns import RecoMoLQueryEmbeddingsFn
from rails_b200.similarities.mol.similarity_fn import MoLSimilarity, SoftmaxDropoutCombiner


def init_mlp_xavier_weights_zero_bias(m) -> None:
    if isinstance(m, torch.nn.Linear):
        torch.nn.init.xavier_uniform_(m.weight)
        if getattr(m, "bias", None) is not None:
            m.bias.data.fill_(0.0)


def create_mol_interaction_module(
    query_embedding_dim: int,
    item_embedding_dim: int,
    dot_product_dimension: int,
    query_dot_product_groups: int,
    item_dot_product_groups: int,
    temperature: float,
    query_dropout_rate: float,
    query_hidden_dim: int,
    item_dropout_rate: float,
    item_hidden_dim: int,
    gating_query_hidden_dim: int,
    gating_qi_hidden_dim: int,
    gating_item_hidden_dim: int,
    softmax_dropout_rate: float,
    bf16_training: bool,
    gating_query_fn: bool = True,
    gating_item_fn: bool = True,
    dot_product_l2_norm: bool = True,
    query_nonlinearity: str = "geglu",
    item_nonlinearity: str = "geglu",
    uid_dropout_rate: float = 0.5,
    uid_embedding_hash_sizes: Optional[List[int]] = None,
    uid_embedding_level_dropout: bool = False,
    gating_combination_type: str = "glu_silu",
    gating_item_dropout_rate: float = 0.0,
    gating_qi_dropout_rate: float = 0.0,
    eps: float = 1e-6,
) -> Tuple[MoLSimilarity, str]:
    if query_hidden_dim <= 0:
        raise ValueError("rails_b200 supports query_hidden_dim > 0 (GLU query projection) only")
    if item_hidden_dim > 0:
        raise ValueError("rails_b200 supports item_hidden_dim = -1 (single Linear item projection) only")
    if gating_qi_hidden_dim <= 0 or not gating_query_fn or not gating_item_fn:
        raise ValueError("rails_b200 requires the query-only, item-only and qi gating MLPs")
    if query_nonlinearity not in ("geglu", "swiglu"):
        raise ValueError(f"Unknown query_nonlinearity {query_nonlinearity}")

    def query_proj(input_dim: int, output_dim: int) -> torch.nn.Module:
        glu = GeGLU if query_nonlinearity == "geglu" else SwiGLU
        return torch.nn.Sequential(
            torch.nn.Dropout(p=query_dropout_rate),
            glu(in_features=input_dim, out_features=query_hidden_dim),
            torch.nn.Linear(in_features=query_hidden_dim, out_features=output_dim),
        )

    def item_proj(input_dim: int, output_dim: int) -> torch.nn.Module:
        return torch.nn.Sequential(
            torch.nn.Dropout(p=item_dropout_rate),
            torch.nn.Linear(in_features=input_dim, out_features=output_dim),
        ).apply(init_mlp_xavier_weights_zero_bias)

    def gq(input_dim: int, output_dim: int) -> torch.nn.Module:
        return torch.nn.Sequential(
            torch.nn.Linear(in_features=input_dim, out_features=gating_query_hidden_dim),
            torch.nn.SiLU(),
            torch.nn.Linear(in_features=gating_query_hidden_dim, out_features=output_dim, bias=False),
        ).apply(init_mlp_xavier_weights_zero_bias)

    def gi(input_dim: int, output_dim: int) -> torch.nn.Module:
        return torch.nn.Sequential(
            torch.nn.Dropout(p=gating_item_dropout_rate),
            torch.nn.Linear(in_features=input_dim, out_features=gating_item_hidden_dim),
            torch.nn.SiLU(),
            torch.nn.Linear(in_features=gating_item_hidden_dim, out_features=output_dim, bias=False),
        ).apply(init_mlp_xavier_weights_zero_bias)

    def gqi(input_dim: int, output_dim: int) -> torch.nn.Module:
        return torch.nn.Sequential(
            torch.nn.Dropout(p=gating_qi_dropout_rate),
            torch.nn.Linear(in_features=input_dim, out_features=gating_qi_hidden_dim),
            torch.nn.SiLU(),
            torch.nn.Linear(in_features=gating_qi_hidden_dim, out_features=output_dim),
        ).apply(init_mlp_xavier_weights_zero_bias)

    mol_module = MoLSimilarity(
        query_embedding_dim=query_embedding_dim,
        item_embedding_dim=item_embedding_dim,
        dot_product_dimension=dot_product_dimension,
        query_dot_product_groups=query_dot_product_groups,
        item_dot_product_groups=item_dot_product_groups,
        temperature=temperature,
        dot_product_l2_norm=dot_product_l2_norm,
        query_embeddings_fn=RecoMoLQueryEmbeddingsFn(
            query_embedding_dim=query_embedding_dim,
            query_dot_product_groups=query_dot_product_groups,
            dot_product_dimension=dot_product_dimension,
            dot_product_l2_norm=dot_product_l2_norm,
            proj_fn=query_proj,
            uid_embedding_hash_sizes=uid_embedding_hash_sizes or [],
            uid_dropout_rate=uid_dropout_rate,
            uid_embedding_level_dropout=uid_embedding_level_dropout,
            eps=eps,
        ),
        item_embeddings_fn=RecoMoLItemEmbeddingsFn(
            item_embedding_dim=item_embedding_dim,
            item_dot_product_groups=item_dot_product_groups,
            dot_product_dimension=dot_product_dimension,
            dot_product_l2_norm=dot_product_l2_norm,
            proj_fn=item_proj,
            eps=eps,
        ),
        item_proj_fn=None,
        gating_query_only_partial_fn=gq,
        gating_item_only_partial_fn=gi,
        gating_qi_partial_fn=gqi,
        gating_combination_type=gating_combination_type,
        gating_normalization_fn=lambda _: SoftmaxDropoutCombiner(dropout_rate=softmax_dropout_rate, eps=1e-6),
        eps=eps,
        autocast_bf16=bf16_training,
    )
    debug_str = (
        f"MoL-{query_dot_product_groups}x{item_dot_product_groups}x{dot_product_dimension}"
        + f"-t{temperature}-d{softmax_dropout_rate}"
        + f"{'-l2' if dot_product_l2_norm else ''}"
        + f"-q{query_hidden_dim}d{query_dropout_rate}{query_nonlinearity}"
        + f"-id{item_dropout_rate}"
        + f"-gq{gating_query_hidden_dim}"
        + f"-gi{gating_item_hidden_dim}d{gating_item_dropout_rate}"
        + f"-gqi{gating_qi_hidden_dim}d{gating_qi_dropout_rate}-x-{gating_combination_type}"
    )
    if uid_embedding_hash_sizes is not None:
        debug_str += f"-uids{'-'.join([str(x) for x in uid_embedding_hash_sizes])}"
        if uid_dropout_rate > 0.0:
            debug_str += f"d{uid_dropout_rate}"
        if uid_embedding_level_dropout:
            debug_str += "-el"
    return mol_module, debug_str
