"""Generates tests/golden/next_*.npz from the UNMODIFIED reference (build container only):
   CandidateIndex.get_top_k_outputs (indexing/candidate_index.py:116-185) driven by the reference's
   MIPSBruteForceTopK (rails/indexing/mips_top_k.py:41-81).

    python oracle/gen_golden_next.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.reference_loader import REFERENCE_ROOT, _install_gin_stub  # noqa: E402

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
_install_gin_stub()
sys.path.insert(0, REFERENCE_ROOT)
from indexing.candidate_index import CandidateIndex  # noqa: E402
from rails.indexing.mips_top_k import MIPSBruteForceTopK  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case(name, N, D, B, k, n0, seed, force_short_rows=False):
    g = torch.Generator().manual_seed(seed)
    items = torch.randn(N, D, generator=g)
    ids = torch.randperm(N, generator=g) + 1
    q = torch.randn(B, D, generator=g)
    top = MIPSBruteForceTopK(items.unsqueeze(0), ids.unsqueeze(0))
    index = CandidateIndex(ids=ids.unsqueeze(0), embeddings=items.unsqueeze(0))
    # invalid ids: a mix of each row's own best items (so the mask bites) and random ids; 0 = padding
    full_s, full_i = top(q, k=min(N, k + n0))
    inv = torch.randint(1, N + 1, (B, n0), generator=g)
    take = min(n0, full_i.size(1))
    hit = torch.rand(B, take, generator=g) < (0.95 if force_short_rows else 0.4)
    inv[:, :take] = torch.where(hit, full_i[:, :take], inv[:, :take])
    inv[:, -1] = 0
    out_ids, out_scores, _ = index.get_top_k_outputs(
        query_embeddings=q, k=k, aux_payloads={}, top_k_module=top, invalid_ids=inv
    )
    kp = min(k + n0, N)
    ps, pi = top(q, k=kp)
    np.savez_compressed(
        os.path.join(OUT, f"next_{name}.npz"),
        items=items.numpy(), ids=ids.numpy(), queries=q.numpy(), invalid_ids=inv.numpy(), k=np.int64(k),
        ref_prime_scores=ps.numpy(), ref_prime_ids=pi.numpy(), ref_ids=out_ids.numpy(), ref_scores=out_scores.numpy(),
    )
    print(name, "k'", kp, "->", tuple(out_ids.shape))


if __name__ == "__main__":
    case("mask_basic", N=500, D=16, B=6, k=10, n0=12, seed=1)
    case("mask_short_rows", N=40, D=8, B=5, k=30, n0=25, seed=2, force_short_rows=True)  # k' clamps to N, rows back-fill
    case("mask_wide", N=3000, D=32, B=4, k=100, n0=211, seed=3)
