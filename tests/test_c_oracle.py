"""The plain-C restatement (oracle/mol_oracle.c) against the outputs of the unmodified reference (tests/golden/*.npz) and
against the Python oracle on seeded inputs.  CPU only; the C library is compiled on first use (gcc, seconds)."""
import pytest
import torch

from oracle import c_oracle
from oracle import mol_oracle as O
from tests.golden_util import golden_names, load_golden
from tests.helpers import CFG_8x4x64, CFG_8x8x32, build_module, synthetic_inputs

NAMES = golden_names()


def test_c_oracle_builds_and_loads():
    assert b"plain C" in c_oracle.load().molc_version()


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_matches_reference_scores(name):
    g = load_golden(name)
    s = c_oracle.similarity(g["cfg"], g["sd"], g["queries"], g["items"], g.get("user_ids"))
    assert s.shape == g["ref_scores"].shape
    # same arithmetic as the reference, different summation order than ATen's GEMMs
    assert (s - g["ref_scores"]).abs().max().item() <= 5e-5


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_topk_matches_reference(name):
    g = load_golden(name)
    s = c_oracle.similarity(g["cfg"], g["sd"], g["queries"], g["items"], g.get("user_ids"))
    k = g["k"]
    top_s, top_i = torch.topk(s, k=k, dim=1)
    ids = g["item_ids"].reshape(-1)[top_i]
    r = O.compare_top_k(top_s, ids, g["ref_scores"], g["item_ids"], k, 1e-3, 1e-4)
    assert r["ok"] == 1.0, r


@pytest.mark.parametrize("cfg,seed", [(CFG_8x8x32, 4), (CFG_8x4x64, 5)])
def test_c_oracle_matches_python_oracle_on_seeded_inputs(cfg, seed):
    mol, _ = build_module(cfg, None, "cpu", seed=seed)
    sd = {k: v.detach() for k, v in mol.state_dict().items()}
    items, _, q, uid = synthetic_inputs(cfg, 777, 5, seed, "cpu")
    a = c_oracle.similarity(cfg, sd, q, items, uid)
    b = O.similarity(cfg, sd, q, items, uid)
    assert (a - b).abs().max().item() <= 5e-5
    assert torch.equal(torch.topk(a, 10, dim=1).indices, torch.topk(b, 10, dim=1).indices)
