"""Stage timeline of the coarse kernel (MOL_TRACE build): CTA 0, slot 0.  Prints clk offsets per query."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOL_B200_LIB"] = os.path.join(ROOT, "rails_b200", "lib", "libmol_b200_trace.so")
import torch
from rails_b200 import _lib, engine
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _lib.load()
dev = "cuda:0"
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, B, 0, dev)
top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
top(q, k=100)
torch.cuda.synchronize()
buf = torch.zeros(3 * 256 * 8, dtype=torch.int64, device=dev)
lib.mol_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
top(q, k=100)
torch.cuda.synchronize()
t = buf.cpu().view(3, 256, 8)
base = t[0, 0, 0].item()
A, Bq, I = t[0] - base, t[1] - base, t[2] - base
print("A: E1start logfull E1done | E3start gatefull E3end ;  B: wait hidfull e2a e2done ;  I: w_e1 e1 g2iss e2a+gf g3a_iss e2d g3b_iss")
for j in list(range(0, 6)) + list(range(100, 112)):
    print(j, "A", A[j, :6].tolist(), "B", Bq[j, :4].tolist(), "I", I[j, :7].tolist())
import numpy as np
a = A.numpy(); b = Bq.numpy(); i = I.numpy()
s = slice(40, 200)
print("per-query period (A E1start delta):", np.diff(a[s, 0]).mean())
print("A: wait log_full", (a[s, 1] - a[s, 0]).mean(), "E1 work", (a[s, 2] - a[s, 1]).mean(), "wait gate_full", (a[s, 4] - a[s, 3]).mean(), "E3 work", (a[s, 5] - a[s, 4]).mean())
print("B: wait hid_full", (b[s, 1] - b[s, 0]).mean(), "E2 first half", (b[s, 2] - b[s, 1]).mean(), "E2 second half", (b[s, 3] - b[s, 2]).mean())
print("I: wait e1", (i[s, 1] - i[s, 0]).mean(), "issue G2+G1", (i[s, 2] - i[s, 1]).mean(), "wait e2a/gate_free", (i[s, 3] - i[s, 2]).mean(), "issue G3a", (i[s, 4] - i[s, 3]).mean(), "wait e2d", (i[s, 5] - i[s, 4]).mean(), "issue G3b", (i[s, 6] - i[s, 5]).mean())
print("latency e1_done arrive -> issuer acquired:", (i[s, 1] - a[s, 2]).mean(), "; G2 issued -> hid_full seen by B:", (b[s, 1] - i[s, 2]).mean(),
      "; e2_done arrive -> issuer:", (i[s, 5] - b[s, 3]).mean(), "; G3b issued -> gate_full seen by A:", (a[s, 4] - i[s, 6]).mean())
