"""Small-batch searches for a launch-list profile: cfg1 (ML-1M checkpoint, B=1, 3883 items) and B=1 over 1M items.
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_small.csv python tools/small_batch.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from rails_b200.workloads import CFG_8x8x32, build_module, synthetic_inputs
from tests.golden_util import load_golden

dev = "cuda:0"
g = load_golden("cfg1_ml1m_ckpt")
mol1, _ = build_module(g["cfg"], g["sd"], dev)
top1 = MoLBruteForceTopK(mol1, g["items"].to(dev).unsqueeze(0), g["item_ids"].to(dev).unsqueeze(0))
u, q1 = g["user_ids"][:1].to(dev), g["queries"][:1].to(dev)
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, 1, 0, dev)
top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
for _ in range(2):
    top1(q1, k=10, user_ids=u)
    top(q, k=100)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("MARK_cfg1")
top1(q1, k=10, user_ids=u)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
torch.cuda.nvtx.range_push("MARK_b1")
top(q, k=100)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
