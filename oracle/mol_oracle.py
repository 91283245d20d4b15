"""CPU oracle for the MoL brute-force top-k path.  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement (torch CPU ops, fp32 or fp64) of the reference's
eval-mode arithmetic for ``MoLBruteForceTopK.forward``.  It is the CHECKER: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``rails_b200/`` imports it and the
product path raises when the CUDA library is missing.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container from ``/root/reference`` by ``oracle/gen_golden.py``; the resulting
fixtures live in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this
restatement against every one of them (max |Δscore| == 0.0 on the fp32 path, because
the restatement issues the same ATen ops in the same order).

Reference lines followed (all relative to the reference root):
  rails/similarities/mol/query_embeddings_fns.py:175-254   query sub-embeddings (+uid hash)
  rails/similarities/layers.py:19-74                       GeGLU / SwiGLU
  rails/similarities/mol/item_embeddings_fns.py:149-183    item sub-embeddings
  rails/similarities/mol/similarity_fn.py:341-413          einsum + /temperature
  rails/similarities/mol/similarity_fn.py:148-201          gating (glu_silu branch)
  rails/similarities/mol/similarity_fn.py:31-46            softmax / renorm / weighted sum
  rails/indexing/mol_top_k.py:99-130                       topk + id map
State-dict key names are those of ``MoLSimilarity`` built by
``modeling/similarity_utils.py:41-245``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


from rails_b200.workloads import MoLConfig  # noqa: E402,F401  (shape dataclass; pure Python, no CUDA library)


# state-dict keys (reference MoLSimilarity; SURVEY.md §8a "Weights of the path")
K_Q_GLU_W = "_query_embeddings_fn._query_emb_proj_module.1._w"
K_Q_GLU_B = "_query_embeddings_fn._query_emb_proj_module.1._b"
K_Q_OUT_W = "_query_embeddings_fn._query_emb_proj_module.2.weight"
K_Q_OUT_B = "_query_embeddings_fn._query_emb_proj_module.2.bias"
K_UID = "_query_embeddings_fn._uid_embeddings_{}.weight"
K_X_W = "_item_embeddings_fn._item_emb_proj_module.1.weight"
K_X_B = "_item_embeddings_fn._item_emb_proj_module.1.bias"
K_GQ_W1 = "_gating_fn._query_only_partial_module.0.weight"
K_GQ_B1 = "_gating_fn._query_only_partial_module.0.bias"
K_GQ_W2 = "_gating_fn._query_only_partial_module.2.weight"
K_GI_W1 = "_gating_fn._item_only_partial_module.1.weight"
K_GI_B1 = "_gating_fn._item_only_partial_module.1.bias"
K_GI_W2 = "_gating_fn._item_only_partial_module.3.weight"
K_QI_W1 = "_gating_fn._qi_partial_module.1.weight"
K_QI_B1 = "_gating_fn._qi_partial_module.1.bias"
K_QI_W2 = "_gating_fn._qi_partial_module.3.weight"
K_QI_B2 = "_gating_fn._qi_partial_module.3.bias"


def _cast(sd: Dict[str, torch.Tensor], dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def _l2norm(x: torch.Tensor, eps: float) -> torch.Tensor:
    # query_embeddings_fns.py:244-253 / item_embeddings_fns.py:173-182
    return x / torch.clamp(torch.linalg.norm(x, ord=None, dim=-1, keepdim=True), min=eps)


def query_sub_embeddings(
    cfg: MoLConfig, sd: Dict[str, torch.Tensor], q: torch.Tensor, user_ids: Optional[torch.Tensor]
) -> torch.Tensor:
    """(B, D) -> (B, P_Q, d).  query_embeddings_fns.py:191-253, layers.py:36-43/67-74."""
    B = q.size(0)
    u = len(cfg.uid_embedding_hash_sizes)
    hidden = sd[K_Q_GLU_W].size(1) // 2
    pre = torch.mm(q.reshape(-1, cfg.query_embedding_dim), sd[K_Q_GLU_W]) + sd[K_Q_GLU_B]
    lhs, rhs = torch.split(pre, [hidden, hidden], dim=-1)
    act = F.gelu(lhs) if cfg.query_nonlinearity == "geglu" else F.silu(lhs)
    h = act * rhs
    out = F.linear(h, sd[K_Q_OUT_W], sd[K_Q_OUT_B]).reshape(
        B, cfg.query_dot_product_groups - u, cfg.dot_product_dimension
    )
    if u > 0:
        parts = [out]
        for i, hash_size in enumerate(cfg.uid_embedding_hash_sizes):
            idx = (user_ids % hash_size) + 1  # :205-207
            parts.append(F.embedding(idx, sd[K_UID.format(i)]).unsqueeze(1))
        out = torch.cat(parts, dim=1)
    return _l2norm(out, cfg.eps)


def item_sub_embeddings(cfg: MoLConfig, sd: Dict[str, torch.Tensor], items: torch.Tensor) -> torch.Tensor:
    """(..., D) -> (..., P_X, d).  item_embeddings_fns.py:165-182 (single Linear, item_hidden_dim=-1)."""
    out = F.linear(items, sd[K_X_W], sd[K_X_B]).reshape(
        items.size()[:-1] + (cfg.item_dot_product_groups, cfg.dot_product_dimension)
    )
    return _l2norm(out, cfg.eps)


def _mlp_silu(x: torch.Tensor, w1, b1, w2, b2=None) -> torch.Tensor:
    return F.linear(F.silu(F.linear(x, w1, b1)), w2, b2)


def similarity(
    cfg: MoLConfig,
    sd: Dict[str, torch.Tensor],
    query_embeddings: torch.Tensor,
    item_embeddings: torch.Tensor,
    user_ids: Optional[torch.Tensor] = None,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """Scores (B, N) of every query against every item; eval mode.

    query_embeddings (B, D); item_embeddings (1, N, D) or (N, D).
    Follows MoLSimilarity.forward (B'==1 branch) -> MoLGatingFn.forward (glu_silu) ->
    _softmax_dropout_combiner_fn with training=False.
    """
    sd = _cast(sd, dtype)
    q = query_embeddings.to(dtype)
    items = item_embeddings.to(dtype)
    if items.dim() == 2:
        items = items.unsqueeze(0)
    B, N, L = q.size(0), items.size(1), cfg.num_logits

    qs = query_sub_embeddings(cfg, sd, q, user_ids)  # (B, P_Q, d)
    xs = item_sub_embeddings(cfg, sd, items)  # (1, N, P_X, d)
    logits = torch.einsum("bnd,xmd->bxnm", qs, xs.squeeze(0)).reshape(B, N, L)  # :389-396
    logits = logits / cfg.temperature  # :405

    gq = _mlp_silu(q, sd[K_GQ_W1], sd[K_GQ_B1], sd[K_GQ_W2]).unsqueeze(1)  # (B,1,L)  :166-169
    gi = _mlp_silu(items, sd[K_GI_W1], sd[K_GI_B1], sd[K_GI_W2])  # (1,N,L)  :170-171
    gqi = _mlp_silu(logits, sd[K_QI_W1], sd[K_QI_B1], sd[K_QI_W2], sd[K_QI_B2])  # (B,N,L) :172-173
    g = gq * gi + gqi  # :175-178
    w = g * torch.sigmoid(g)  # :179

    p = F.softmax(w, dim=-1)  # :42
    if cfg.softmax_dropout_rate > 0.0:  # :43-45 (dropout is identity in eval; renorm still runs)
        p = p / torch.clamp(p.sum(-1, keepdim=True), min=cfg.eps)
    return (p * logits).sum(-1)  # :46


def brute_force_top_k(
    cfg: MoLConfig,
    sd: Dict[str, torch.Tensor],
    query_embeddings: torch.Tensor,
    item_embeddings: torch.Tensor,
    item_ids: torch.Tensor,
    k: int,
    user_ids: Optional[torch.Tensor] = None,
    dtype: torch.dtype = torch.float32,
    chunk: int = 0,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """mol_top_k.py:117-130.  Returns (top scores (B,k), top ids (B,k) int64, all scores (B,N)).

    ``chunk`` > 0 evaluates queries in micro-batches (the reference's
    ``user_max_batch_size`` semantics, data/eval.py:131-138) so (chunk, N, 128) fits RAM.
    """
    B = query_embeddings.size(0)
    if chunk and chunk < B:
        outs = [
            similarity(
                cfg, sd, query_embeddings[s : s + chunk], item_embeddings,
                None if user_ids is None else user_ids[s : s + chunk], dtype,
            )
            for s in range(0, B, chunk)
        ]
        scores = torch.cat(outs, dim=0)
    else:
        scores = similarity(cfg, sd, query_embeddings, item_embeddings, user_ids, dtype)
    top_s, top_i = torch.topk(scores, dim=1, k=k, sorted=True, largest=True)
    return top_s, item_ids.reshape(-1)[top_i], scores


def compare_top_k(
    got_scores: torch.Tensor,
    got_ids: torch.Tensor,
    all_scores: torch.Tensor,
    item_ids: torch.Tensor,
    k: int,
    score_tol: float = 1e-3,
    tie_tol: float = 1e-4,
) -> Dict[str, float]:
    """Tie-aware comparator (SURVEY.md §7 hard part 2).

    Strict: got ids == oracle ids position by position.
    Tie-aware pass: every returned item's ORACLE score is within ``tie_tol`` of the oracle's
    score at that rank (i.e. only near-tied neighbours may swap), the returned score is
    within ``score_tol`` of that item's oracle score, and no id repeats within a row.
    """
    all_scores = all_scores.double().cpu()
    got_scores = got_scores.double().cpu()
    got_ids = got_ids.cpu()
    ids_flat = item_ids.reshape(-1).cpu()
    ref_s, ref_i = torch.topk(all_scores, k=k, dim=1, sorted=True, largest=True)
    ref_ids = ids_flat[ref_i]
    # map returned ids -> column index
    order = torch.argsort(ids_flat)
    pos = order[torch.searchsorted(ids_flat[order], got_ids.reshape(-1))].reshape(got_ids.shape)
    assert bool((ids_flat[pos] == got_ids).all()), "returned an id that is not in the corpus"
    oracle_of_got = torch.gather(all_scores, 1, pos)
    strict = (got_ids == ref_ids).all(dim=1).double().mean().item()
    max_score_err = (got_scores - oracle_of_got).abs().max().item()
    rank_gap = (oracle_of_got - ref_s).abs().max().item()
    dup = max(int(k - torch.unique(r).numel()) for r in got_ids)
    return {
        "strict_row_match": strict,
        "max_score_err": max_score_err,
        "max_rank_gap": rank_gap,
        "duplicates": float(dup),
        "ok": float(max_score_err <= score_tol and rank_gap <= tie_tol and dup == 0),
    }
