"""Secondary timings on one GPU (not the headline metric): index build, query prologue, exact mode, MIPS top-k,
seen-item masking.  Prints one JSON object; CUDA-event timing, 3 warm-up + 10 timed runs each."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200 import _lib, engine
from rails_b200.indexing.candidate_index import CandidateIndex
from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.helpers import CFG_8x8x32, CFG_8x4x64, CFG_8x4x128, build_module, synthetic_inputs


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
out = {}
cfg = CFG_8x8x32
N, B, k = 1_000_000, 512, 100
mol, _ = build_module(cfg, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(cfg, N, B, 0, dev)
w = mol.packed_weights(dev)
out["index_build_1M_items_ms"] = timed(lambda: engine.IndexHandle(w, items, ids), 1, 3)
out["query_prologue_B512_ms"] = timed(lambda: engine.query_prologue(w, mol.workspace(dev), q, None))
top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
out["mol_top100_B512_1M_ms"] = timed(lambda: top(q, k=k))
for b in (1, 8, 32, 128):
    out[f"mol_top100_B{b}_1M_ms"] = timed(lambda: top(q[:b], k=k))
ex = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)
out["mol_exact_mode_B16_1M_ms"] = timed(lambda: ex(q[:16], k=k), 1, 2)
mips = MIPSBruteForceTopK(items.unsqueeze(0), ids.unsqueeze(0))
out["mips_top100_B512_1M_ms"] = timed(lambda: mips(q, k=k))
inv = torch.randint(1, N + 1, (B, 211), device=dev)
index = CandidateIndex(ids=ids.unsqueeze(0), embeddings=items.unsqueeze(0))
out["candidate_index_mol_k100_n0_211_B512_1M_ms"] = timed(lambda: index.get_top_k_outputs(q, k, {}, top, inv))
# approximate modules of the reference, same corpus; B = 64 (their score matrices are chunked)
from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLCombTopK, MoLNaiveTopK
for name, mod in (("mol_avg_top2000", MoLAvgTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 2000)),
                  ("mol_naive_kpg5", MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 5)),
                  ("mol_comb_kpg5_avg200", MoLCombTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 200, 5))):
    out[f"{name}_B64_1M_ms"] = timed(lambda: mod(q[:64], k=k), 1, 3)
# the reference's own configs (BASELINE.json configs[0..2])
for name, c, n, b, kk in (("cfg1_8x4x64_N3883_B1_k10", CFG_8x4x64, 3883, 1, 10), ("cfg2_8x4x128_N27278_B128_k100", CFG_8x4x128, 27278, 128, 100),
                          ("cfg3_8x8x32_N695762_B256_k200", CFG_8x8x32, 695762, 256, 200)):
    m2, _ = build_module(c, None, dev, seed=1)
    it2, id2, q2, u2 = synthetic_inputs(c, n, b, 1, dev)
    t2 = MoLBruteForceTopK(m2, it2.unsqueeze(0), id2.unsqueeze(0))
    kw = {} if u2 is None else {"user_ids": u2}
    ms = timed(lambda: t2(q2, k=kk, **kw))
    out[name + "_ms"] = ms
    out[name + "_qps"] = b / (ms * 1e-3)
print(json.dumps(out))
