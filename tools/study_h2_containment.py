"""CPU study: does the MUFU-free half2 silu (MOL_E2_H2_MASK / MOL_E3_H2_OF4 build knobs of the coarse kernel) keep the
coarse pass good enough as a candidate filter?  Uses the numerics model of the kernel (tests/sim_coarse.py) against the
fp32 oracle on the north-star workload shape (8x8x32, top-100, K' = 256), item-chunked so it fits host memory.

    python tools/study_h2_containment.py [N=1000000] [B=8] [seed=0]

Per variant: rms / max |coarse - exact|, the fraction of queries whose exact top-k lies inside the coarse top-K', and the
fraction the kernel's safety check (c_min + 1.5 err + 1e-3 >= exact score at rank k) would send to the exact fallback.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import mol_oracle as O
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs
from tests.sim_coarse import coarse_scores, containment

VARIANTS = [
    ("shipped (MUFU tanh, fp16)", dict()),
    ("E2 h2 chunks 1-3 (0x0E)", dict(e2_h2_mask=0x0E)),
    ("E2 h2 all (0xFF)", dict(e2_h2_mask=0xFF)),
    ("E2 h2 all + E3 h2 2/4", dict(e2_h2_mask=0xFF, e3_h2_of4=2)),
    ("E2 h2 all + E3 h2 4/4", dict(e2_h2_mask=0xFF, e3_h2_of4=4)),
    ("LITE: E2 h2 all", dict(e2_h2_mask=0xFF, lite=True)),
    ("LITE: E2 h2 all + E3 h2 4/4", dict(e2_h2_mask=0xFF, e3_h2_of4=4, lite=True)),
]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    k, kp, chunk = 100, 256, 25_000
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = CFG_8x8x32
    mol, _ = build_module(cfg, None, "cpu", seed=seed)
    sd = {k_: v.detach() for k_, v in mol.state_dict().items()}
    items, _, q, _ = synthetic_inputs(cfg, N, B, seed, "cpu")
    with torch.inference_mode():
        t0 = time.time()
        exact = torch.cat([O.similarity(cfg, sd, q, items[i : i + chunk], None) for i in range(0, N, chunk)], dim=1)
        print(f"N={N} B={B} k={k} K'={kp}; exact scores in {time.time() - t0:.1f} s; "
              f"score at rank k / rank K' (mean over queries): {torch.topk(exact, kp, dim=1).values[:, [k - 1, kp - 1]].mean(0).tolist()}")
        for name, kw in VARIANTS:
            t0 = time.time()
            co = torch.cat([coarse_scores(cfg, sd, q, items[i : i + chunk], None, **kw) for i in range(0, N, chunk)], dim=1)
            d = co - exact
            ok, flagged = containment(co, exact, k, kp)
            # error on the items that matter (the coarse top-K'): what the safety check sees
            top = torch.topk(co, kp, dim=1).indices
            dt = torch.gather(d, 1, top)
            print(f"{name:32s} rms {d.pow(2).mean().sqrt().item():.2e} max {d.abs().max().item():.2e} | top-K' max {dt.abs().max().item():.2e} "
                  f"| exact top-k inside coarse top-K': {ok:.3f} | flagged for fallback: {flagged:.3f}  ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
