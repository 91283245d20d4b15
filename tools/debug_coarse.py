"""Debug aid: raw coarse-pass scores of the loaded library (MOL_B200_LIB) against the exact fp32 scores, error pattern
by query / tile / row.   python tools/debug_coarse.py [N] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from rails_b200 import engine
from rails_b200.workloads import CFG_8x8x32, CFG_8x4x64, CFG_16x16x64, build_module, synthetic_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = {"8x4x64": CFG_8x4x64, "16x16x64": CFG_16x16x64}.get(os.environ.get("DBG_CFG", ""), CFG_8x8x32)
dev = torch.device("cuda:0")
mol, _ = build_module(cfg, None, dev, seed=3)
items, ids, q, uid = synthetic_inputs(cfg, N, B, 3, dev)
w = mol.packed_weights(dev)
idx = mol.build_index(items, ids)
a = engine.score_all(w, idx, mol.workspace(dev), q, uid, coarse=True)
e = engine.score_all(w, idx, mol.workspace(dev), q, uid)
torch.cuda.synchronize()
d = (a - e).abs()
print("max err", float(d.max()), "finite", bool(torch.isfinite(a).all()))
print("per query max:", [round(float(x), 3) for x in d.max(dim=1).values])
T = (N + 127) // 128
pad = T * 128 - N
dd = torch.nn.functional.pad(d, (0, pad)).view(B, T, 128)
print("per tile max:", [round(float(x), 3) for x in dd.amax(dim=(0, 2))][:32])
print("per (query, tile) max:")
for b in range(min(B, 8)):
    print("  q", b, [round(float(x), 2) for x in dd[b].amax(dim=1)][:16])
print("row pattern (q0, tile0) first 16 rows:", [round(float(x), 2) for x in dd[0, 0, :16]])
print("coarse q0[:8]", [round(float(x), 3) for x in a[0, :8]], "\nexact  q0[:8]", [round(float(x), 3) for x in e[0, :8]])
