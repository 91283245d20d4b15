"""One B = 1 (and B = 8) MoL search over 1M items for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs
dev = torch.device("cuda:0")
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, 64, 0, dev)
top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
top(q[:b], k=100)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
top(q[:b], k=100)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
