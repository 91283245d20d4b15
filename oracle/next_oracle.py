"""CPU oracle for the callers / siblings of the MoL top-k path.  TEST INFRASTRUCTURE ONLY.

Restates, with torch CPU ops,
  * the seen-item masking + back-fill of ``CandidateIndex.get_top_k_outputs``
    (reference indexing/candidate_index.py:155-178), and
  * ``MIPSBruteForceTopK.forward`` (reference rails/indexing/mips_top_k.py:74-81).
Pinned against outputs of the unmodified reference classes by ``oracle/gen_golden_next.py`` ->
``tests/golden/next_*.npz`` (``tests/test_next_oracle_golden.py``).
"""
from __future__ import annotations

from typing import Tuple

import torch


def select_valid(
    top_k_prime_scores: torch.Tensor, top_k_prime_ids: torch.Tensor, invalid_ids: torch.Tensor, k: int
) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B, k') score-sorted candidates + (B, N0) invalid ids -> (scores (B, k), ids (B, k)).  candidate_index.py:155-178."""
    is_seen = (top_k_prime_ids.unsqueeze(2) == invalid_ids.unsqueeze(1)).max(2)[0]  # :156
    valid = ~is_seen
    valid = torch.logical_and(valid, torch.cumsum(valid.int(), dim=1) <= k)  # :158
    inv = ~valid  # :163 (seen ids and valid ids beyond the first k)
    gap = k - valid.int().sum(1, keepdim=True)  # :164
    valid = torch.logical_or(valid, torch.logical_and(inv, torch.cumsum(inv.int(), dim=1) <= gap))  # :165-171
    offsets = torch.nonzero(valid, as_tuple=True)[1].view(-1, k)  # :174
    return torch.gather(top_k_prime_scores, 1, offsets), torch.gather(top_k_prime_ids, 1, offsets)


def mips_top_k(
    query_embeddings: torch.Tensor, item_embeddings: torch.Tensor, item_ids: torch.Tensor, k: int
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(B, D), (N, D), (N,) -> (top scores (B, k), top ids (B, k), all logits (B, N)).  mips_top_k.py:74-81."""
    all_logits = torch.mm(query_embeddings, item_embeddings.t())
    s, i = torch.topk(all_logits, dim=1, k=k, sorted=True, largest=True)
    return s, item_ids.reshape(-1)[i], all_logits


def mol_avg_top_k(cfg, sd, query_embeddings, item_embeddings, item_ids, k: int, avg_top_k: int, user_ids=None):
    """MoLAvgTopK.forward with fp32 component embeddings (mol_top_k.py:331-385).
    Returns (top scores (B, k), top ids (B, k), prefilter positions (B, avg_top_k))."""
    from oracle import mol_oracle as O

    qs = O.query_sub_embeddings(cfg, sd, query_embeddings, user_ids)  # (B, P_Q, d)
    xs = O.item_sub_embeddings(cfg, sd, item_embeddings)  # (N, P_X, d)
    avg_items = xs.sum(1) / xs.size(1)  # :322-324
    avg_sim = torch.mm(qs.sum(1), avg_items.t())  # :352-356
    _, pos = torch.topk(avg_sim, k=avg_top_k, dim=1)  # :357-360
    scores = torch.stack(
        [
            O.similarity(cfg, sd, query_embeddings[b : b + 1], item_embeddings[pos[b]],
                         None if user_ids is None else user_ids[b : b + 1])[0]
            for b in range(query_embeddings.size(0))
        ]
    )  # :368-373 (B'==B branch == per-query candidate lists)
    s, j = torch.topk(scores, k=min(k, avg_top_k), dim=1, largest=True, sorted=True)  # :374-380
    return s, item_ids.reshape(-1)[torch.gather(pos, 1, j)], pos
