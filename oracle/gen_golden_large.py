"""Generates tests/golden/large_8x8x32_150k.npz by RUNNING THE UNMODIFIED REFERENCE on a corpus large enough for the
fused candidate-filter strategy of the CUDA path (>= 65 536 items; build container only).

    python -m oracle.gen_golden_large      # from the repo root; needs /root/reference

The raw item embeddings (150 000 x 64 fp32 = 38 MB) are NOT stored: they are regenerated from the seed by
`large_inputs()` below (torch's CPU generator is deterministic for a given torch version; the fixture carries a digest
of the generated tensor, and the loader refuses to compare when it does not match).  Stored: the reference module's
state dict, the queries, the reference's own top-k (scores, ids) from MoLBruteForceTopK.forward and every 61st column
of the reference's full (B, N) score matrix from MoLSimilarity.forward.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rails_b200.workloads import MoLConfig  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NAME = "large_8x8x32_150k"
N, B, K, SEED, COL_STRIDE = 150_000, 8, 200, 29, 61
CFG = MoLConfig(64, 64, 32, 8, 8, 0.05, "geglu", ())


def large_inputs():
    """(items (N, 64), item_ids (N), queries (B, 64)) of the fixture, from the seed."""
    g = torch.Generator().manual_seed(SEED + 1)
    items = 0.02 * torch.randn(N, CFG.item_embedding_dim, generator=g)
    item_ids = torch.randperm(N, generator=g).to(torch.int64) + 1000  # ids are not positions
    g = torch.Generator().manual_seed(SEED + 100)
    queries = F.layer_norm(torch.randn(B, CFG.query_embedding_dim, generator=g), (CFG.query_embedding_dim,))
    return items, item_ids, queries


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    from oracle import reference_loader as rl

    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(SEED)
    mol = rl.build_reference_mol(CFG)
    _, MoLBruteForceTopK = rl.import_reference()
    items, item_ids, queries = large_inputs()
    with torch.inference_mode():
        scores = torch.cat([mol(queries[b : b + 2], items.unsqueeze(0))[0] for b in range(0, B, 2)])
        top = MoLBruteForceTopK(mol, items.unsqueeze(0), item_ids.unsqueeze(0))
        parts = [top(queries[b : b + 2], k=K, sorted=True) for b in range(0, B, 2)]
    top_s = torch.cat([p[0] for p in parts])
    top_i = torch.cat([p[1] for p in parts])
    arrays = {f"sd::{k_}": v.detach().cpu().numpy() for k_, v in mol.state_dict().items()}
    arrays.update(
        cfg=np.frombuffer(json.dumps(CFG.to_json()).encode(), dtype=np.uint8),
        queries=queries.numpy(),
        k=np.int64(K),
        ref_top_scores=top_s.numpy(),
        ref_top_ids=top_i.numpy(),
        ref_scores_strided=scores[:, ::COL_STRIDE].contiguous().numpy(),
        items_sha256=np.frombuffer(digest(items).encode(), dtype=np.uint8),
        ids_sha256=np.frombuffer(digest(item_ids).encode(), dtype=np.uint8),
    )
    path = os.path.join(OUT, NAME + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{NAME}: N={N} B={B} k={K} -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
