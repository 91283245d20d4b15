"""Shapes and synthetic workloads of the MoL path: the dataclass that names a MoL head, the
reference's gin bindings as factory kwargs, and the seeded synthetic corpora / queries bench.py and the tests use.
Pure host-side setup (no oracle, no CUDA library)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn.functional as F


@dataclass
class MoLConfig:
    """Shape/hyper-parameters of one MoL head (names follow the reference's ctor kwargs)."""

    query_embedding_dim: int
    item_embedding_dim: int
    dot_product_dimension: int
    query_dot_product_groups: int
    item_dot_product_groups: int
    temperature: float = 0.05
    query_nonlinearity: str = "geglu"  # "geglu" | "swiglu"
    uid_embedding_hash_sizes: Tuple[int, ...] = ()
    softmax_dropout_rate: float = 0.2
    eps: float = 1e-6

    @property
    def num_logits(self) -> int:
        return self.query_dot_product_groups * self.item_dot_product_groups

    def to_json(self) -> dict:
        d = dict(self.__dict__)
        d["uid_embedding_hash_sizes"] = list(self.uid_embedding_hash_sizes)
        return d

    @staticmethod
    def from_json(d: dict) -> "MoLConfig":
        d = dict(d)
        d["uid_embedding_hash_sizes"] = tuple(d.get("uid_embedding_hash_sizes", ()))
        return MoLConfig(**d)


def factory_kwargs(cfg: MoLConfig) -> dict:
    """The gin bindings of the reference's MoL configs (configs/*/hstu-mol-*.gin), as kwargs."""
    return dict(
        query_embedding_dim=cfg.query_embedding_dim,
        item_embedding_dim=cfg.item_embedding_dim,
        dot_product_dimension=cfg.dot_product_dimension,
        query_dot_product_groups=cfg.query_dot_product_groups,
        item_dot_product_groups=cfg.item_dot_product_groups,
        temperature=cfg.temperature,
        query_dropout_rate=0.0,
        query_hidden_dim=512,
        item_dropout_rate=0.1,
        item_hidden_dim=-1,
        gating_query_hidden_dim=128,
        gating_qi_hidden_dim=128,
        gating_item_hidden_dim=128,
        softmax_dropout_rate=cfg.softmax_dropout_rate,
        bf16_training=False,
        query_nonlinearity=cfg.query_nonlinearity,
        item_nonlinearity=cfg.query_nonlinearity,
        uid_dropout_rate=0.5,
        uid_embedding_hash_sizes=list(cfg.uid_embedding_hash_sizes) or None,
        gating_combination_type="glu_silu",
        eps=cfg.eps,
    )


def build_module(cfg: MoLConfig, sd=None, device="cpu", seed=0):
    """MoLSimilarity of `cfg` with the reference's default initialisers (seeded) or a given state dict, in eval mode."""
    from rails_b200.modeling.similarity_utils import create_mol_interaction_module

    torch.manual_seed(seed)
    mol, debug_str = create_mol_interaction_module(**factory_kwargs(cfg))
    if sd is not None:
        mol.load_state_dict(sd, strict=True)
    mol.eval()
    return mol.to(device), debug_str


def synthetic_inputs(cfg: MoLConfig, N: int, B: int, seed: int, device="cpu"):
    """items 0.02*N(0,1), queries layer_norm(N(0,1)), ids 1..N, user_ids U[1,1e5)."""
    g = torch.Generator().manual_seed(seed + 1)
    items = 0.02 * torch.randn(N, cfg.item_embedding_dim, generator=g)
    item_ids = torch.arange(1, N + 1, dtype=torch.int64)
    g = torch.Generator().manual_seed(seed + 100)
    queries = F.layer_norm(torch.randn(B, cfg.query_embedding_dim, generator=g), (cfg.query_embedding_dim,))
    user_ids = None
    if cfg.uid_embedding_hash_sizes:
        user_ids = torch.randint(1, 100000, (B,), generator=g, dtype=torch.int64)
        user_ids = user_ids.to(device)
    return items.to(device), item_ids.to(device), queries.to(device), user_ids


# the shapes of BASELINE.json's configs
CFG_8x8x32 = MoLConfig(64, 64, 32, 8, 8, 0.05, "geglu", ())            # north star, configs 3 and 4
CFG_8x4x64 = MoLConfig(50, 50, 64, 8, 4, 0.05, "swiglu", (6040,))      # config 1 (ML-1M)
CFG_8x4x128 = MoLConfig(256, 256, 128, 8, 4, 0.05, "swiglu", (512,))   # config 2 shape (test-sized uid table)
CFG_16x16x64 = MoLConfig(64, 64, 64, 16, 16, 0.05, "geglu", ())        # config 5
