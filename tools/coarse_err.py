"""Prints the coarse (tcgen05) pass' max |score - numerics model| and max |score - exact| on two seeded workloads
(the quantities tests/test_gpu_parity.py::test_coarse_pass_matches_its_numerics_model bounds by 0.03 / 0.08)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import mol_oracle as O
from rails_b200 import engine
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs
from tests.sim_coarse import coarse_scores

out = {}
for N, B, seed in ((5000, 7, 2), (40000, 33, 3)):
    cfg = CFG_8x8x32
    mol, _ = build_module(cfg, None, "cuda", seed=seed)
    items, ids, q, uid = synthetic_inputs(cfg, N, B, seed, "cuda")
    w = mol.packed_weights(torch.device("cuda"))
    idx = mol.build_index(items, ids)
    got = engine.score_all(w, idx, mol.workspace(torch.device("cuda")), q, uid, coarse=True).cpu()
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    sim = coarse_scores(cfg, sd, q.cpu(), items.cpu(), None)
    exact = O.similarity(cfg, sd, q.cpu(), items.cpu(), None)
    out[f"N{N}_B{B}"] = {"err_model": round((got - sim).abs().max().item(), 5), "err_exact": round((got - exact).abs().max().item(), 5),
                         "rms_exact": round((got - exact).pow(2).mean().sqrt().item(), 6)}
print(json.dumps(out))
