"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python -m oracle.gen_golden            # from the repo root; needs /root/reference

Each fixture holds: the MoLConfig (json), the reference module's state dict, the inputs
(queries, user_ids, raw item embeddings, item ids), and the reference's own outputs on CPU fp32:
  ref_scores (B, N)  = MoLSimilarity.forward(queries, items[None])[0]
  ref_top_scores / ref_top_ids (B, k) = MoLBruteForceTopK(mol, items[None], ids[None]).forward(queries, k)
Synthetic inputs follow SURVEY.md §8(d): weights = reference default initialisers under
torch.manual_seed(seed); items 0.02·N(0,1); queries layer_norm(N(0,1)); user_ids ~ U[1, hash].
Fixture `cfg1_ml1m_ckpt` uses the real ML-1M checkpoint weights + item table (known-answer test).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.mol_oracle import MoLConfig  # noqa: E402
from oracle import reference_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, cfg, sd, queries, user_ids, items, item_ids, k, ref_scores, ref_top_s, ref_top_i):
    arrays = {f"sd::{k_}": v.detach().cpu().numpy() for k_, v in sd.items()}
    arrays.update(
        cfg=np.frombuffer(json.dumps(cfg.to_json()).encode(), dtype=np.uint8),
        queries=queries.numpy(),
        items=items.numpy(),
        item_ids=item_ids.numpy(),
        k=np.int64(k),
        ref_scores=ref_scores.numpy(),
        ref_top_scores=ref_top_s.numpy(),
        ref_top_ids=ref_top_i.numpy(),
    )
    if user_ids is not None:
        arrays["user_ids"] = user_ids.numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: N={items.shape[0]} B={queries.shape[0]} k={k} -> {os.path.getsize(path)/1e6:.2f} MB")


@torch.inference_mode()
def _run_reference(mol, queries, items, item_ids, k, user_ids):
    _, MoLBruteForceTopK = rl.import_reference()
    kwargs = {}
    if user_ids is not None:
        # the caller passes all three payloads (data/eval.py:148); only user_ids is consumed
        kwargs = dict(user_ids=user_ids, timestamps=torch.zeros_like(user_ids), ratings=torch.zeros_like(user_ids))
    scores, _ = mol(queries, items.unsqueeze(0), **kwargs)
    topk = MoLBruteForceTopK(mol, items.unsqueeze(0), item_ids.unsqueeze(0))
    top_s, top_i = topk(queries, k=k, sorted=True, **kwargs)
    return scores, top_s, top_i


def synthetic(name, cfg: MoLConfig, N, B, k, seed):
    torch.manual_seed(seed)
    mol = rl.build_reference_mol(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    items = 0.02 * torch.randn(N, cfg.item_embedding_dim, generator=g)
    item_ids = torch.arange(1, N + 1, dtype=torch.int64)
    g = torch.Generator().manual_seed(seed + 100)
    queries = F.layer_norm(torch.randn(B, cfg.query_embedding_dim, generator=g), (cfg.query_embedding_dim,))
    user_ids = None
    if cfg.uid_embedding_hash_sizes:
        user_ids = torch.randint(1, 100000, (B,), generator=g, dtype=torch.int64)
    scores, top_s, top_i = _run_reference(mol, queries, items, item_ids, k, user_ids)
    _save(name, cfg, mol.state_dict(), queries, user_ids, items, item_ids, k, scores, top_s, top_i)


def ml1m(name, B, k, seed):
    cfg = MoLConfig(50, 50, 64, 8, 4, 0.05, "swiglu", (6040,))
    mol = rl.build_reference_mol(cfg)
    sd, table = rl.load_ml1m_checkpoint()
    missing = mol.load_state_dict(sd, strict=True)
    print("ml-1m checkpoint:", missing)
    item_ids = rl.ml1m_item_ids()
    items = table[item_ids]
    g = torch.Generator().manual_seed(seed)
    queries = F.layer_norm(torch.randn(B, 50, generator=g), (50,))
    user_ids = torch.randint(1, 6041, (B,), generator=g, dtype=torch.int64)
    scores, top_s, top_i = _run_reference(mol, queries, items, item_ids, k, user_ids)
    _save(name, cfg, mol.state_dict(), queries, user_ids, items, item_ids, k, scores, top_s, top_i)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ml1m("cfg1_ml1m_ckpt", B=4, k=10, seed=7)
    synthetic("cfg2_8x4x128", MoLConfig(256, 256, 128, 8, 4, 0.05, "swiglu", (512,)), N=1500, B=16, k=100, seed=11)
    synthetic("cfg3_8x8x32", MoLConfig(64, 64, 32, 8, 8, 0.05, "geglu", ()), N=4096, B=16, k=200, seed=13)
    synthetic("cfg5_16x16x64", MoLConfig(64, 64, 64, 16, 16, 0.05, "geglu", ()), N=1024, B=8, k=100, seed=17)
    # edge cases: ragged N (not a tile multiple), k == N, single query, tiny corpus
    synthetic("edge_ragged_kmax", MoLConfig(64, 64, 32, 8, 8, 0.05, "geglu", ()), N=333, B=3, k=333, seed=19)
    synthetic("edge_tiny", MoLConfig(64, 64, 32, 8, 8, 0.05, "geglu", ()), N=5, B=1, k=1, seed=23)


if __name__ == "__main__":
    main()
