"""Imports the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by ``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by
``oracle/gen_golden_large.py`` / ``oracle/gen_golden_next.py`` (build container only: /root/reference is absent on the GPU box).
The reference's ``modeling/similarity_utils.py`` imports ``gin`` (not installed): a stub whose
``configurable`` is the identity decorator is put on ``sys.modules`` first, then
``create_mol_interaction_module`` is called with explicit kwargs taken from the .gin files
(configs/ml-1m/hstu-mol-…gin:52-79 etc.).  ``@torch.compile`` at similarity_fn.py:31 is
disabled (TORCHDYNAMO_DISABLE=1 -> eager), as in SURVEY.md §8c.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, Optional, Tuple

REFERENCE_ROOT = os.environ.get("RAILS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rails", "similarities"))


def _install_gin_stub() -> None:
    if "gin" in sys.modules:
        return
    gin = types.ModuleType("gin")

    def configurable(fn=None, **_kw):
        if fn is None or isinstance(fn, str):
            return lambda f: f
        return fn

    gin.configurable = configurable
    gin.REQUIRED = object()
    sys.modules["gin"] = gin


def import_reference():
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_gin_stub()
    # this repo ships a `rails` ALIAS package (rails/__init__.py -> rails_b200): it must never stand in for the reference
    # here.  Drop any alias modules already imported, put the reference first on sys.path, and verify what was imported.
    for name in [m for m in sys.modules if m == "rails" or m.startswith("rails.")]:
        if not (getattr(sys.modules[name], "__file__", None) or "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    if sys.path[0] != REFERENCE_ROOT:
        sys.path.insert(0, REFERENCE_ROOT)
    import modeling.similarity_utils as su  # noqa: E402
    import rails.indexing.mol_top_k as ref_top_k  # noqa: E402

    for mod in (su, ref_top_k):
        if not os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)):
            raise RuntimeError(f"{mod.__name__} was imported from {mod.__file__}, not from the reference")
    return su, ref_top_k.MoLBruteForceTopK


def build_reference_mol(cfg) -> "torch.nn.Module":
    """cfg: oracle.mol_oracle.MoLConfig.  Mirrors the gin bindings of the MoL configs."""
    su, _ = import_reference()
    mol, debug_str = su.create_mol_interaction_module(
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
    mol.eval()
    return mol


ML1M_CKPT = os.path.join(
    REFERENCE_ROOT,
    "ckpts/ml-1m-l200/HSTU-b8-h2-dqk25-dv25-lsilud0.2-ad0.0_MoL-8x4x64-t0.05-d0.2-l2-q512d0.0swiglu-id0.1-gq128-"
    "gi128d0.0-gqi128d0.0-x-glu_silu-uids6040d0.5_local_ssl-n128-lwuid_embedding_l2_norm:0.1-mi_loss:0.001-b128-"
    "lr0.001-wu0-wd0.001-2024-11-06_ep72",
)


def load_ml1m_checkpoint() -> Tuple[Dict[str, "torch.Tensor"], "torch.Tensor"]:
    """Returns (MoL state dict with the `module._ndp_module.` prefix stripped, item table (3953, 50))."""
    import torch

    ck = torch.load(ML1M_CKPT, map_location="cpu", weights_only=False)
    sd = ck["model_state_dict"]
    pref = "module._ndp_module."
    mol_sd = {k[len(pref):]: v.float() for k, v in sd.items() if k.startswith(pref)}
    # legacy key rename, eval_from_checkpoint.py:366-376
    mol_sd = {
        k.replace("_item_proj_module.", "_item_embeddings_fn._item_emb_proj_module."): v
        for k, v in mol_sd.items()
    }
    item_table = sd["module._embedding_module._item_emb.weight"].float()
    return mol_sd, item_table


def ml1m_item_ids():
    """Item ids as the reference builds them for ML-1M (data/reco_dataset.py: all rows of movies.csv)."""
    import pandas as pd
    import torch

    path = os.path.join(REFERENCE_ROOT, "tmp/processed/ml-1m/movies.csv")
    df = pd.read_csv(path)
    col = "movie_id" if "movie_id" in df.columns else df.columns[0]
    return torch.tensor(sorted(df[col].astype(int).tolist()), dtype=torch.int64)
