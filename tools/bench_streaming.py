"""Timings of the streaming dot-product top-k (mol_dotfilter) against the materialised-matrix path it replaces:
MIPS top-100 and the approximate MoL modules over 1M items.  Prints one JSON object (CUDA events, 3 warm-up + 10 runs)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLCombTopK, MoLNaiveTopK
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dev = torch.device("cuda:0")
cfg = CFG_8x8x32
N, k = 1_000_000, 100
mol, _ = build_module(cfg, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(cfg, N, 512, 0, dev)
out = {}
for stream in ("1", "0"):
    os.environ["MOL_B200_DOTFILTER"] = stream
    tag = "stream" if stream == "1" else "matrix"
    mips = MIPSBruteForceTopK(items.unsqueeze(0), ids.unsqueeze(0))
    for b in (1, 64, 512):
        out[f"mips_top100_B{b}_1M_{tag}_ms"] = timed(lambda: mips(q[:b], k=k))
    out[f"mips_{tag}_stats"] = mips.last_search_stats()
    mods = (("mol_avg_top2000", MoLAvgTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 2000)),
            ("mol_naive_kpg5", MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 5)),
            ("mol_comb_kpg5_avg200", MoLCombTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 200, 5)))
    for name, mod in mods:
        for b in ((64, 512) if stream == "1" else (64,)):
            out[f"{name}_B{b}_1M_{tag}_ms"] = timed(lambda: mod(q[:b], k=k), 1, 3)
        out[f"{name}_{tag}_stats"] = mod.last_search_stats()
print(json.dumps(out))
