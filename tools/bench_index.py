"""Index build timing (1M items, 8x8x32): tensor-core tf32 x 3 build vs the CUDA-core build."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200 import engine
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs

dev = torch.device("cuda:0")
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, 8, 0, dev)
w = mol.packed_weights(dev)
out = {}
for x3 in ("1", "0"):
    os.environ["MOL_B200_INDEX_X3"] = x3
    for _ in range(2):
        engine.IndexHandle(w, items, ids)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        engine.IndexHandle(w, items, ids)
    e1.record()
    torch.cuda.synchronize()
    out["index_build_1M_ms_" + ("tensor" if x3 == "1" else "cuda_core")] = e0.elapsed_time(e1) / 5
print(json.dumps(out))
