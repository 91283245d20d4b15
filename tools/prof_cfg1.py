"""One cfg1 search (ML-1M checkpoint, 3883 items, one query) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.golden_util import load_golden
from tests.helpers import build_module
dev = "cuda:0"
g = load_golden("cfg1_ml1m_ckpt")
mol, _ = build_module(g["cfg"], g["sd"], dev)
top = MoLBruteForceTopK(mol, g["items"].to(dev).unsqueeze(0), g["item_ids"].to(dev).unsqueeze(0))
q, uid = g["queries"][:1].to(dev), g["user_ids"][:1].to(dev)
top(q, k=10, user_ids=uid)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
top(q, k=10, user_ids=uid)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
