import sys, os
sys.path.insert(0, "/root/repo")
import torch
from rails_b200 import _lib, engine
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from rails_b200.workloads import CFG_8x8x32, build_module, synthetic_inputs
N, B = int(sys.argv[1]), int(sys.argv[2])
mol, _ = build_module(CFG_8x8x32, None, "cuda:0", seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, N, B, 0, "cuda:0")
top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
s, i = top(q, k=100)
torch.cuda.synchronize()
print("ok", N, B, engine.search_stats(mol.workspace(torch.device("cuda:0"))))
