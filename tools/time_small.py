"""Eager and CUDA-graph timings of the two small-batch searches: cfg1 (ML-1M checkpoint, 3883 items, one query) and one query
over 1M items (8x8x32).  Prints one JSON object."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from rails_b200.workloads import CFG_8x8x32, build_module, synthetic_inputs
from tests.golden_util import load_golden


def timed(fn, warm=5, reps=50):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


dev = "cuda:0"
out = {}
g = load_golden("cfg1_ml1m_ckpt")
mol1, _ = build_module(g["cfg"], g["sd"], dev)
u, q1 = g["user_ids"][:1].to(dev), g["queries"][:1].to(dev)
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, 8, 0, dev)
for graph in (False, True):
    tag = "graph" if graph else "eager"
    top1 = MoLBruteForceTopK(mol1, g["items"].to(dev).unsqueeze(0), g["item_ids"].to(dev).unsqueeze(0), cuda_graph=graph)
    out[f"cfg1_us_{tag}"] = timed(lambda: top1(q1, k=10, user_ids=u))
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), cuda_graph=graph)
    for b in (1, 8):
        out[f"b{b}_1m_us_{tag}"] = timed(lambda: top(q[:b], k=100))
print(json.dumps(out))
