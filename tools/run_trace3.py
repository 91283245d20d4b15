"""Stage timeline of mol_coarse3_kernel (build: python -m rails_b200.build --variant trace3 MOL_COARSE_V3=1 MOL_TRACE=1):
CTA 0, slot 0.  Events (TR(role, ev, idx) in csrc/mol_coarse_v3.cuh), idx = the slot's query counter:
  issuer I (role 2): 0 before wait gate_free, 1 after (G1 issue), 2 e1_done seen (G2 issue), 3 e2a_done + a2_read seen (G3a),
                     4 e2b_done seen (G3b)
  E2 a / b (roles 1 / 3): 0 before wait log_full, 1 after, 2 E1 done, 3 hid_full seen, 4 E2 done
  E3 (role 0): 0 before wait e1_done, 1 after, 2 logits copied, 3 gate_full seen, 4 gate in registers (gate_free), 5 score done
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOL_B200_LIB"] = os.path.join(ROOT, "rails_b200", "lib", "libmol_b200_trace3.so")
import numpy as np
import torch
from rails_b200 import _lib
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from rails_b200.workloads import CFG_8x8x32, build_module, synthetic_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _lib.load()
dev = "cuda:0"
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, B, 0, dev)
top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
top(q, k=100)
torch.cuda.synchronize()
buf = torch.zeros(4 * 256 * 8, dtype=torch.int64, device=dev)
lib.mol_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
top(q, k=100)
torch.cuda.synchronize()
t = buf.cpu().view(4, 256, 8).numpy()
e3, ea, iss, eb = t[0], t[1], t[2], t[3]
s = slice(40, 200)
def m(x):
    return round(float(x[s].mean()), 1)
# events: issuer 0 e1_done seen (G2 issue), 1 G2 + early G1 issued, 2 e2_done / lg_free / e1_done(next) seen (G3 issue), 3 G3 issued
#         E2a/b  0 before wait log_full, 1 after (+ a2_read), 2 E1 done, 3 hid_full seen, 4 E2 done
#         E3     0 before wait gate_full, 1 after, 2 gate + logits in registers (lg_free), 3 score done
print("per-query period of the slot (clk):", m(np.diff(iss[:, 0])[39:199]))
print("issuer: G2 issue + early G1 (incl. its waits)", m(iss[:, 1] - iss[:, 0]), "| wait for G3's conditions", m(iss[:, 2] - iss[:, 1]),
      "| G3 issue", m(iss[:, 3] - iss[:, 2]), "| -> next e1_done seen", m(np.roll(iss[:, 0], -1) - iss[:, 3]))
for name, e in (("E2a", ea), ("E2b", eb)):
    print(name, ": wait log_full/a2_read", m(e[:, 1] - e[:, 0]), "| E1", m(e[:, 2] - e[:, 1]), "| wait hid_full", m(e[:, 3] - e[:, 2]),
          "| E2", m(e[:, 4] - e[:, 3]))
print("E3 : wait gate_full", m(e3[:, 1] - e3[:, 0]), "| loads", m(e3[:, 2] - e3[:, 1]), "| math", m(e3[:, 3] - e3[:, 2]),
      "| tail", m(np.roll(e3[:, 0], -1) - e3[:, 3]))
for j in range(100, 103):
    b0 = iss[j, 0]
    print(j, "I", (iss[j, :4] - b0).tolist(), "Ea", (ea[j, :5] - b0).tolist(), "Eb", (eb[j, :5] - b0).tolist(), "E3", (e3[j, :4] - b0).tolist())
