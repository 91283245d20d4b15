"""Loads tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np
import torch

from rails_b200.workloads import MoLConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """MoL fixtures (oracle/gen_golden.py); the next_*.npz fixtures of oracle/gen_golden_next.py have their own loader."""
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith(("next_", "large_")))


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


def load_large(name="large_8x8x32_150k"):
    """The >= 100k-item fixture of oracle/gen_golden_large.py: items / ids are regenerated from the seed and checked
    against the digests stored beside the reference's outputs."""
    from oracle import gen_golden_large as G

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    items, item_ids, queries = G.large_inputs()
    assert G.digest(items) == bytes(z["items_sha256"]).decode(), "seeded item embeddings differ from the generating run"
    assert G.digest(item_ids) == bytes(z["ids_sha256"]).decode(), "seeded item ids differ from the generating run"
    out = {"sd": {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}}
    out["cfg"] = MoLConfig.from_json(json.loads(bytes(z["cfg"]).decode()))
    out["k"] = int(z["k"])
    out["items"], out["item_ids"] = items, item_ids
    assert torch.equal(queries, torch.from_numpy(z["queries"]))
    out["queries"] = queries
    out["ref_top_scores"] = torch.from_numpy(z["ref_top_scores"])
    out["ref_top_ids"] = torch.from_numpy(z["ref_top_ids"])
    out["ref_scores_strided"] = torch.from_numpy(z["ref_scores_strided"])
    out["col_stride"] = G.COL_STRIDE
    return out
