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


def avg_case(name, cfg, N, B, k, avg_top_k, seed):
    """Reference MoLAvgTopK (rails/indexing/mol_top_k.py:296-429).  Stored twice: with the reference's default bf16
    component embeddings, and with them replaced by fp32 (what rails_b200 computes)."""
    import json

    from oracle.reference_loader import build_reference_mol
    from rails.indexing.mol_top_k import MoLAvgTopK
    from tests.helpers import synthetic_inputs

    torch.manual_seed(seed)
    mol = build_reference_mol(cfg).eval()
    items, ids, q, uid = synthetic_inputs(cfg, N, B, seed)
    kw = {} if uid is None else {"user_ids": uid}
    with torch.inference_mode():
        top = MoLAvgTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), avg_top_k)
        s_bf16, i_bf16 = top(q, k=k, **kw)
        comp, _ = mol.get_item_component_embeddings(items, decoupled_inference=True)
        top._mol_item_embeddings = comp.float()
        top._avg_mol_item_embeddings_t = (comp.float().sum(1) / comp.size(1)).transpose(0, 1)
        s_f32, i_f32 = top(q, k=k, **kw)
    out = {f"sd::{k_}": v.detach().numpy() for k_, v in mol.state_dict().items()}
    out.update(
        cfg=np.frombuffer(json.dumps(cfg.to_json()).encode(), dtype=np.uint8), items=items.numpy(), item_ids=ids.numpy(),
        queries=q.numpy(), k=np.int64(k), avg_top_k=np.int64(avg_top_k),
        ref_scores_f32=s_f32.numpy(), ref_ids_f32=i_f32.numpy(), ref_scores_bf16=s_bf16.float().numpy(), ref_ids_bf16=i_bf16.numpy(),
    )
    if uid is not None:
        out["user_ids"] = uid.numpy()
    np.savez_compressed(os.path.join(OUT, f"next_{name}.npz"), **out)
    print(name, tuple(i_f32.shape), "bf16-vs-fp32 prefilter id overlap:",
          float(np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(i_f32, i_bf16)])))


def groups_case(name, cfg, N, B, k_per_group, avg_top_k, seed):
    """Reference MoLNaiveTopK (avg_top_k == 0; mol_top_k.py:133-293) / MoLCombTopK (:432-551).  Stored twice, as avg_case."""
    import json

    from oracle.reference_loader import build_reference_mol
    from rails.indexing.mol_top_k import MoLCombTopK, MoLNaiveTopK
    from tests.helpers import synthetic_inputs

    torch.manual_seed(seed)
    mol = build_reference_mol(cfg).eval()
    items, ids, q, uid = synthetic_inputs(cfg, N, B, seed)
    kw = {} if uid is None else {"user_ids": uid}
    with torch.inference_mode():
        if avg_top_k > 0:
            top = MoLCombTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), avg_top_k, k_per_group)
        else:
            top = MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), k_per_group)
        s_bf16, i_bf16 = top(q, k=10, **kw)
        comp, _ = mol.get_item_component_embeddings(items, decoupled_inference=True)
        comp = comp.float()
        P_X, D_P = comp.size(1), comp.size(2)
        top._mol_item_embeddings = comp
        top._mol_item_embeddings_t = comp.permute(1, 0, 2).reshape(-1, D_P).transpose(0, 1)
        if avg_top_k > 0:
            top._avg_top_k_module._mol_item_embeddings = comp
            top._avg_top_k_module._avg_mol_item_embeddings_t = (comp.sum(1) / P_X).transpose(0, 1)
        s_f32, i_f32 = top(q, k=10, **kw)
    out = {f"sd::{k_}": v.detach().numpy() for k_, v in mol.state_dict().items()}
    out.update(
        cfg=np.frombuffer(json.dumps(cfg.to_json()).encode(), dtype=np.uint8), items=items.numpy(), item_ids=ids.numpy(),
        queries=q.numpy(), k_per_group=np.int64(k_per_group), avg_top_k=np.int64(avg_top_k),
        ref_scores_f32=s_f32.numpy(), ref_ids_f32=i_f32.numpy(), ref_scores_bf16=s_bf16.float().numpy(), ref_ids_bf16=i_bf16.numpy(),
    )
    if uid is not None:
        out["user_ids"] = uid.numpy()
    np.savez_compressed(os.path.join(OUT, f"next_{name}.npz"), **out)
    n_valid = (s_f32 > -32767.0).sum(1)
    print(name, tuple(i_f32.shape), "distinct candidates per query:", n_valid.tolist())


if __name__ == "__main__":
    from tests.helpers import CFG_8x8x32, CFG_8x4x64

    groups_case("naive_8x8x32", CFG_8x8x32, N=3000, B=4, k_per_group=5, avg_top_k=0, seed=21)
    groups_case("naive_8x4x64_uid", CFG_8x4x64, N=1200, B=3, k_per_group=10, avg_top_k=0, seed=22)
    groups_case("comb_8x8x32", CFG_8x8x32, N=3000, B=4, k_per_group=5, avg_top_k=100, seed=23)
    groups_case("comb_8x4x64_uid", CFG_8x4x64, N=1200, B=3, k_per_group=4, avg_top_k=50, seed=24)
    avg_case("avg_8x8x32", CFG_8x8x32, N=4000, B=5, k=20, avg_top_k=200, seed=11)
    avg_case("avg_8x4x64_uid", CFG_8x4x64, N=1500, B=4, k=10, avg_top_k=64, seed=12)
    case("mask_basic", N=500, D=16, B=6, k=10, n0=12, seed=1)
    case("mask_short_rows", N=40, D=8, B=5, k=30, n0=25, seed=2, force_short_rows=True)  # k' clamps to N, rows back-fill
    case("mask_wide", N=3000, D=32, B=4, k=100, n0=211, seed=3)
