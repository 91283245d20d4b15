"""Loads tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np
import torch

from oracle.mol_oracle import MoLConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """MoL fixtures (oracle/gen_golden.py); the next_*.npz fixtures of oracle/gen_golden_next.py have their own loader."""
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith("next_"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {"sd": {}}
    for key in z.files:
        if key.startswith("sd::"):
            out["sd"][key[4:]] = torch.from_numpy(z[key])
        elif key == "cfg":
            out["cfg"] = MoLConfig.from_json(json.loads(bytes(z[key]).decode()))
        elif key == "k":
            out["k"] = int(z[key])
        else:
            out[key] = torch.from_numpy(z[key])
    out.setdefault("user_ids", None)
    return out
