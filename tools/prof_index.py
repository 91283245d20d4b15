"""One tensor-core index build of 1M items (8x8x32) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rails_b200 import engine
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs
dev = torch.device("cuda:0")
mol, _ = build_module(CFG_8x8x32, None, dev, seed=0)
items, ids, q, _ = synthetic_inputs(CFG_8x8x32, 1_000_000, 8, 0, dev)
engine.IndexHandle(mol.packed_weights(dev), items, ids)
torch.cuda.synchronize()
