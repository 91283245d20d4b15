"""One MIPS (B = 512) and one MoLNaiveTopK (B = 64) call over 1M items, for an ncu launch list of the streaming path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLNaiveTopK
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs

dev = torch.device("cuda:0")
cfg = CFG_8x8x32
mol, _ = build_module(cfg, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(cfg, 1_000_000, 512, 0, dev)
which = sys.argv[1] if len(sys.argv) > 1 else "mips"
if which == "mips":
    top = MIPSBruteForceTopK(items.unsqueeze(0), ids.unsqueeze(0))
    top(q, k=100)
elif which == "naive":
    top = MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 5)
    top(q[:64], k=100)
else:
    top = MoLAvgTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 2000)
    top(q[:64], k=100)
torch.cuda.synchronize()
