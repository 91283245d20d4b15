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


# fp16 constants of the MUFU-free half2 silu(2u) (tools/fit_silu_h2.py; csrc/mol_coarse_sm100.cu kH2*)
_H2 = {
    False: dict(inv_a=-0.16666667, nc=(-0.052520752, -0.35986328, -1.6826172, 1.0927734)),  # A = 6, Q degree 3
    True: dict(inv_a=-0.18181818, nc=(-0.060974121, -1.0429688)),                            # MOL_H2_LITE
}


def _fma16(a, b, c):
    return _h(a * b + c)  # fp16 operands: the fp32 product is exact; one fp32 + one fp16 rounding of the sum


def _h2_parts(u16, lite=False):
    """a = |u|, aw = a * relu(1 - a/A)^4, nq = -Q(w): the kernel's h2_bump_parts, one fp16 rounding per instruction."""
    k = _H2[lite]
    a = u16.abs()
    y = torch.relu(_fma16(a, _h(torch.tensor(k["inv_a"])), 1.0))
    y = _h(y * y)
    w = _h(y * y)
    aw = _h(a * w)
    nc = [_h(torch.tensor(v)) for v in k["nc"]]
    nq = nc[-1].expand_as(w)
    for c in reversed(nc[:-1]):
        nq = _fma16(nq, w, c)
    return a, aw, nq


def silu2_h2(u, lite=False):
    """silu(2u) by the half2 form: (u + |u|) - s(|u|), everything in fp16 (E2)."""
    u16 = _h(u)
    a, aw, nq = _h2_parts(u16, lite)
    return _fma16(aw, nq, _h(u16 + a))


def silu2_h2_f32(u, lite=False):
    """E3 form: the bump -s(|u|) in half2, the large part u + |u| and the sum in fp32."""
    a, aw, nq = _h2_parts(_h(u), lite)
    return (u + u.abs()) + _h(aw * nq)


def coarse_scores(cfg, sd, queries, items, user_ids=None, e2_h2_mask=0, e3_h2_of4=0, lite=False):
    """e2_h2_mask / e3_h2_of4 / lite model the kernel's MOL_E2_H2_MASK / MOL_E3_H2_OF4 / MOL_H2_LITE build knobs
    (default 0: the shipped kernel; the fp32 polynomial chunks of MOL_E2_POLY_MASK are modelled as exact tanh)."""
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
    if e2_h2_mask:
        unit = torch.arange(u.size(-1))
        sel = ((e2_h2_mask >> (unit // 16)) & 1).bool()  # chunk c = hidden units [16c, 16c + 16)
        h = torch.where(sel, silu2_h2(u, lite), h)
    ug = (
        torch.nn.functional.linear(h, w2h)
        + _h(0.5 * gq).unsqueeze(1) * _h(gi).unsqueeze(0)
        + _h(0.5 * sd[O.K_QI_B2])
    )
    w = ug + ug * torch.tanh(ug)
    if e3_h2_of4:
        # the kernel walks the logits in its own order l' = m * P_Q + n, two per packed pair; pair j2 of a 16-column
        # chunk is selected when (j2 & 3) >= 4 - of4
        l = torch.arange(w.size(-1))
        PQ, PX = cfg.query_dot_product_groups, cfg.item_dot_product_groups
        lk = (l % PX) * PQ + (l // PX)
        sel = ((lk // 2) % 4) >= 4 - e3_h2_of4
        w = torch.where(sel, silu2_h2_f32(ug, lite), w)
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
