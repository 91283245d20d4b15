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
import numpy as np
a = A.numpy(); b = Bq.numpy(); i = I.numpy()
print("rows: A[p=0] k-th own query: [.,.,.,E3start,gate_full,E3end]; B0 per query: [hid_wait,hid_full,-,e2done,log_wait,log_full]; I per query: [w_e1,e1,g2g1_iss,e2a+gf,g3a_iss,e2b,g3b_iss]")
for j in list(range(100, 108)):
    print(j, "A", a[j, 3:6].tolist(), "B", b[j, [4, 5, 0, 1, 3]].tolist(), "I", i[j, :7].tolist())
s = slice(40, 200)
print("issuer period per query:", np.diff(i[s, 0]).mean())
print("I: wait e1", (i[s, 1] - i[s, 0]).mean(), "issue G2+G1", (i[s, 2] - i[s, 1]).mean(), "wait e2a/gate_free", (i[s, 3] - i[s, 2]).mean(), "issue G3a", (i[s, 4] - i[s, 3]).mean(), "wait e2b", (i[s, 5] - i[s, 4]).mean(), "issue G3b", (i[s, 6] - i[s, 5]).mean())
print("B0: wait log_full", (b[s, 5] - b[s, 4]).mean(), "E1 work", (b[s, 0] - b[s, 5]).mean(), "wait hid_full", (b[s, 1] - b[s, 0]).mean(), "E2 work", (b[s, 3] - b[s, 1]).mean())
sa = slice(20, 100)
print("A0 (every other query): period", np.diff(a[sa, 3]).mean(), "wait gate_full", (a[sa, 4] - a[sa, 3]).mean(), "E3 work", (a[sa, 5] - a[sa, 4]).mean())
