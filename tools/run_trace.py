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
# trace events (mol_coarse_sm100.cu TR(role, ev, idx)), slot 0 of CTA 0, idx = the slot's query counter:
#   issuer  I: 0 before wait e1_done, 1 after, 2 G2 (+ next G1) issued, 3 e2a_done + gate_free seen, 4 G3a issued,
#              5 e2_done seen, 6 G3b issued
#   E1/E2   B: 4 before wait log_full, 5 after, 0 E1 done, 1 hid_full seen, 2 first half of A3 written, 3 E2 done
#   E3      A: 0 before wait e1_done, 1 after, 2 logits copied (a2_read), 3 before wait gate_full (previous query),
#              4 after, 5 E3 done
s = slice(40, 200)
def m(x):
    return round(float(x[s].mean()), 1)
print("per-query period of the slot (clk):", m(np.diff(i[:, 0])[39:199]), " -> per (query, tile) unit of the CTA:", m(np.diff(i[:, 0])[39:199]) / 2)
print("issuer : wait e1_done", m(i[:, 1] - i[:, 0]), "| issue G2+G1", m(i[:, 2] - i[:, 1]), "| wait e2a/gate_free", m(i[:, 3] - i[:, 2]),
      "| issue G3a", m(i[:, 4] - i[:, 3]), "| wait e2_done", m(i[:, 5] - i[:, 4]), "| issue G3b", m(i[:, 6] - i[:, 5]))
print("E1/E2  : wait log_full", m(b[:, 5] - b[:, 4]), "| E1", m(b[:, 0] - b[:, 5]), "| wait hid_full", m(b[:, 1] - b[:, 0]),
      "| E2 first half", m(b[:, 2] - b[:, 1]), "| E2 second half", m(b[:, 3] - b[:, 2]))
print("E3     : wait e1_done", m(a[:, 1] - a[:, 0]), "| copy logits", m(a[:, 2] - a[:, 1]), "| wait gate_full", m(a[:, 4] - a[:, 3]),
      "| E3", m(a[:, 5] - a[:, 4]), "| tail (store, next wait)", m(np.roll(a[:, 0], -1) - a[:, 5]))
for j in range(100, 104):
    print(j, "I", (i[j, :7] - i[j, 0]).tolist(), "B", (b[j, [4, 5, 0, 1, 2, 3]] - i[j, 0]).tolist(), "A", (a[j, :6] - i[j, 0]).tolist())
