"""CPU oracle for the callers / siblings of the MoL top-k path.  TEST INFRASTRUCTURE ONLY.

Restates, with torch CPU ops,
  * the seen-item masking + back-fill of ``CandidateIndex.get_top_k_outputs``
    (reference indexing/candidate_index.py:155-178), and
  * ``MIPSBruteForceTopK.forward`` (reference rails/indexing/mips_top_k.py:74-81),
  * ``MoLAvgTopK.forward`` (rails/indexing/mol_top_k.py:331-385), ``MoLNaiveTopK.forward`` (:205-293) and
    ``MoLCombTopK.forward`` (:469-551) with fp32 component embeddings.
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


def _per_group_positions(qs: torch.Tensor, xs: torch.Tensor, k_per_group: int) -> torch.Tensor:
    """(B, P_Q, d), (N, P_X, d) -> (B, P_Q * P_X * k_per_group) item positions.  mol_top_k.py:160-162, :239-249."""
    B, P_Q, d = qs.shape
    N, P_X, _ = xs.shape
    items_t = xs.permute(1, 0, 2).reshape(-1, d).transpose(0, 1)  # (d, P_X * N)  :160-162
    out = []
    for i in range(P_Q):
        sim = torch.mm(qs[:, i, :], items_t).view(B * P_X, N)  # :241-244
        _, idx = torch.topk(sim, k=k_per_group, dim=1, sorted=False)  # :245-247
        out.append(idx.view(B, P_X * k_per_group))  # :249
    return torch.cat(out, dim=1)


def _score_union(cfg, sd, query_embeddings, item_embeddings, item_ids, all_indices, user_ids):
    """Sort the candidate positions, exact MoL on them, mask duplicates, return ALL candidates sorted by score.
    mol_top_k.py:252-292 (Naive) == :514-551 (Comb)."""
    from oracle import mol_oracle as O

    sorted_idx, _ = torch.sort(all_indices, dim=1)  # :252
    scores = torch.stack(
        [
            O.similarity(cfg, sd, query_embeddings[b : b + 1], item_embeddings[sorted_idx[b]],
                         None if user_ids is None else user_ids[b : b + 1])[0]
            for b in range(query_embeddings.size(0))
        ]
    )  # :257-270
    valid = torch.cat(
        [torch.ones_like(sorted_idx[:, 0:1], dtype=torch.bool), sorted_idx[:, 1:] != sorted_idx[:, :-1]], dim=1
    )  # :273-279
    scores = torch.where(valid, scores, torch.tensor(-32767.0))  # :280
    s, j = torch.topk(scores, k=sorted_idx.size(1), dim=1, largest=True, sorted=True)  # :281-283 (k overwritten at :256)
    return s, item_ids.reshape(-1)[torch.gather(sorted_idx, 1, j)]


def mol_naive_top_k(cfg, sd, query_embeddings, item_embeddings, item_ids, k_per_group: int, user_ids=None):
    """MoLNaiveTopK.forward (use_faiss=False) with fp32 component embeddings.  Returns (scores (B, C), ids (B, C)),
    C = P_Q * P_X * k_per_group; duplicates carry the score -32767."""
    from oracle import mol_oracle as O

    qs = O.query_sub_embeddings(cfg, sd, query_embeddings, user_ids)
    xs = O.item_sub_embeddings(cfg, sd, item_embeddings)
    return _score_union(cfg, sd, query_embeddings, item_embeddings, item_ids, _per_group_positions(qs, xs, k_per_group), user_ids)


def mol_comb_top_k(cfg, sd, query_embeddings, item_embeddings, item_ids, avg_top_k: int, k_per_group: int, user_ids=None):
    """MoLCombTopK.forward with fp32 component embeddings: per-group candidates + MoLAvgTopK.topk_ids (:387-429)."""
    from oracle import mol_oracle as O

    qs = O.query_sub_embeddings(cfg, sd, query_embeddings, user_ids)
    xs = O.item_sub_embeddings(cfg, sd, item_embeddings)
    groups = _per_group_positions(qs, xs, k_per_group)
    avg_items = xs.sum(1) / xs.size(1)  # :322-324
    avg_sim = torch.mm(qs.sum(1) / qs.size(1), avg_items.t())  # :412-422
    _, avg_pos = torch.topk(avg_sim, k=avg_top_k, dim=1, sorted=False)  # :423-426
    return _score_union(cfg, sd, query_embeddings, item_embeddings, item_ids, torch.cat([groups, avg_pos], dim=1), user_ids)
