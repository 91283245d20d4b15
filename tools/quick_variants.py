"""Times several tuning builds of the library in ONE process (one torch import / CUDA init for all of them) on the bench
workload (8x8x32, 1M items, 512 queries, top-100) and checks each against the default build:

    python tools/quick_variants.py default h2_ff h2_0e ...          # names = rails_b200/lib/libmol_b200_<name>.so

Per variant: ms per search step (CUDA events, 3 warm-ups + STEPS timed, device-resident inputs), whether the final
(scores, ids) equal the default build's (they must: final scores come from the exact fp32 rescoring), and the max
|coarse score - default coarse score| on a 20k-item corpus.  Lines are flushed one by one (and appended to
gpurun_out/quick_variants.log) so a call that runs out of time still leaves the variants it finished.
Not a bench: the number to quote is bench.py's.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = time.time()
import torch
import torch.nn.functional as F

from rails_b200 import _lib, engine
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs

STEPS = int(os.environ.get("QV_STEPS", "5"))
N = int(os.environ.get("QV_ITEMS", "1000000"))
B = int(os.environ.get("QV_BATCH", "512"))
K = 100


def use_lib(name):
    path = os.path.join(ROOT, "rails_b200", "lib", "libmol_b200.so" if name == "default" else f"libmol_b200_{name}.so")
    _lib._lib = None
    _lib.LIB_PATH = path
    return _lib.load()


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quick_variants.log"), "a") as f:
        f.write(line + "\n")


def main():
    names = sys.argv[1:] or ["default"]
    dev = torch.device("cuda", 0)
    cfg = CFG_8x8x32
    g = torch.Generator(device=dev).manual_seed(1)
    items = 0.02 * torch.randn(N, cfg.item_embedding_dim, device=dev, generator=g)
    ids = torch.arange(1, N + 1, dtype=torch.int64, device=dev)
    gq = torch.Generator().manual_seed(100)
    q = F.layer_norm(torch.randn(B, cfg.query_embedding_dim, generator=gq), (cfg.query_embedding_dim,)).to(dev)
    small_items, small_ids, small_q, _ = synthetic_inputs(cfg, 20000, 16, 5, dev)
    ref = None
    for name in names:
        rec = {"variant": name, "t_start_s": round(time.time() - T0, 1)}
        try:
            lib = use_lib(name)
            rec["knobs"] = lib.mol_version().decode()
            mol, _ = build_module(cfg, None, dev, seed=0)
            top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
            index = top._ensure_index()
            w, wsp = mol.packed_weights(dev), mol.workspace(dev)
            for _ in range(3):
                s, i = engine.search(w, index, wsp, q, None, K, True, _lib.MODE_AUTO)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(STEPS):
                s, i = engine.search(w, index, wsp, q, None, K, True, _lib.MODE_AUTO)
            e1.record()
            torch.cuda.synchronize()
            rec["ms_per_step"] = round(e0.elapsed_time(e1) / STEPS, 3)
            rec["queries_per_s"] = round(B / (rec["ms_per_step"] * 1e-3))
            sidx = mol.build_index(small_items, small_ids)
            coarse = engine.score_all(w, sidx, wsp, small_q, None, coarse=True)
            torch.cuda.synchronize()
            if ref is None:
                ref = (s.clone(), i.clone(), coarse.clone())
                rec["is_reference"] = True
            else:
                rec["ids_equal_default"] = bool(torch.equal(i, ref[1]))
                rec["max_score_diff_default"] = float((s - ref[0]).abs().max())
                rec["max_coarse_diff_default"] = float((coarse - ref[2]).abs().max())
            rec["coarse_finite"] = bool(torch.isfinite(coarse).all())
        except Exception as e:  # keep going: the next variant may be fine
            rec["error"] = f"{type(e).__name__}: {e}"
        emit(rec)


if __name__ == "__main__":
    main()
