"""CPU model of the NUMERICS of the tcgen05 coarse pass (TEST INFRASTRUCTURE, not product code).

Mirrors rails_b200/csrc/mol_coarse_sm100.cu rounding points with torch ops:
  * X_sub, Q_sub/tau, GI, 0.5*gq, 0.5*W1, 0.5*b1, 0.5*W2, 0.5*b2 rounded to fp16 (tensor-core operands / smem images;
    the biases ride in a "ones" K-block of the GEMMs, gq*gi is a GEMM against diag(0.5 gq)),
  * logits accumulated in fp32, re-rounded to fp16 as the A operand of the hidden GEMM,
  * hidden activations in packed half2: u -> fp16, tanh -> fp16, h = fma(u, tanh u, u) -> fp16 (A operand of the gate GEMM),
  * gate / softmax / weighted sum in fp32 using the fp32 logits.
Used to validate the candidate-set policy (K') on CPU before spending GPU time, and by GPU tests to
bound the coarse-vs-exact error the kernel is expected to show.
"""
import torch

from oracle import mol_oracle as O


def _h(x):
    return x.to(torch.float16).to(torch.float32)


def coarse_scores(cfg, sd, queries, items, user_ids=None):
    sd = {k: v.float() for k, v in sd.items()}
    q = queries.float()
    qs = O.query_sub_embeddings(cfg, sd, q, user_ids)          # (B, Pq, d) fp32 (exact prologue)
    xs = O.item_sub_embeddings(cfg, sd, items.float())         # (N, Px, d)
    gq = O._mlp_silu(q, sd[O.K_GQ_W1], sd[O.K_GQ_B1], sd[O.K_GQ_W2])          # (B, L)
    gi = O._mlp_silu(items.float(), sd[O.K_GI_W1], sd[O.K_GI_B1], sd[O.K_GI_W2])  # (N, L)
    B, N, L = q.size(0), items.size(0), cfg.num_logits
    logits = torch.einsum("bnd,xmd->bxnm", _h(qs / cfg.temperature), _h(xs)).reshape(B, N, L)
    w1h, w2h = _h(0.5 * sd[O.K_QI_W1]), _h(0.5 * sd[O.K_QI_W2])
    u = _h(torch.nn.functional.linear(_h(logits), w1h) + _h(0.5 * sd[O.K_QI_B1]))
    t = _h(torch.tanh(u))
    h = _h(u + u * t)
    ug = (
        torch.nn.functional.linear(h, w2h)
        + _h(0.5 * gq).unsqueeze(1) * _h(gi).unsqueeze(0)
        + _h(0.5 * sd[O.K_QI_B2])
    )
    w = ug + ug * torch.tanh(ug)
    p = torch.softmax(w, dim=-1)
    return (p * logits).sum(-1)


def containment(coarse, exact, k, kp):
    """Fraction of queries whose exact top-k is contained in the coarse top-kp, plus the safety-check outcome."""
    B = exact.size(0)
    ex_i = torch.topk(exact, k, dim=1).indices
    co_s, co_i = torch.topk(coarse, kp, dim=1)
    ok, flagged = 0, 0
    for b in range(B):
        ok += int(set(ex_i[b].tolist()) <= set(co_i[b].tolist()))
        err = (co_s[b] - exact[b, co_i[b]]).abs().max().item()
        s_k = torch.topk(exact[b, co_i[b]], k).values[-1].item()
        flagged += int(co_s[b, -1].item() + 1.5 * err + 1e-3 >= s_k)
    return ok / B, flagged / B
