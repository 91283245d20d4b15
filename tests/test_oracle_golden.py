"""Pins the oracle restatement against outputs of the UNMODIFIED reference (tests/golden/*.npz)."""
import pytest
import torch

from oracle import mol_oracle as O
from tests.golden_util import golden_names, load_golden

NAMES = golden_names()


def test_fixtures_present():
    assert {"cfg1_ml1m_ckpt", "cfg2_8x4x128", "cfg3_8x8x32", "cfg5_16x16x64", "edge_ragged_kmax", "edge_tiny"} <= set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_fp32_matches_reference_scores(name):
    g = load_golden(name)
    s = O.similarity(g["cfg"], g["sd"], g["queries"], g["items"], g["user_ids"], torch.float32)
    assert s.shape == g["ref_scores"].shape
    # same ATen ops in the same order -> expected bit-equal; allow 2e-6 for thread-count dependent GEMM splits
    assert (s - g["ref_scores"]).abs().max().item() <= 2e-6


@pytest.mark.parametrize("name", NAMES)
def test_oracle_topk_matches_reference(name):
    g = load_golden(name)
    top_s, top_ids, scores = O.brute_force_top_k(
        g["cfg"], g["sd"], g["queries"], g["items"], g["item_ids"], g["k"], g["user_ids"]
    )
    r = O.compare_top_k(g["ref_top_scores"], g["ref_top_ids"], scores, g["item_ids"], g["k"], 1e-5, 1e-5)
    assert r["ok"] == 1.0, r
    assert top_ids.dtype == torch.int64 and top_s.shape == (g["queries"].size(0), g["k"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_fp64_close_to_fp32(name):
    g = load_golden(name)
    s64 = O.similarity(g["cfg"], g["sd"], g["queries"], g["items"], g["user_ids"], torch.float64)
    assert (s64.float() - g["ref_scores"]).abs().max().item() < 2e-4


def test_chunked_equals_unchunked():
    g = load_golden("cfg3_8x8x32")
    a = O.brute_force_top_k(g["cfg"], g["sd"], g["queries"], g["items"], g["item_ids"], 50)
    b = O.brute_force_top_k(g["cfg"], g["sd"], g["queries"], g["items"], g["item_ids"], 50, chunk=5)
    assert torch.equal(a[1], b[1])


def test_oracle_matches_reference_on_the_large_fixture():
    """150k-item corpus (the size class of the CUDA path's fused candidate filter): the reference's top-200 and a strided
    sample of its score matrix, reproduced by the oracle from seed-regenerated items."""
    from tests.golden_util import load_large

    g = load_large()
    top_s, top_ids, scores = O.brute_force_top_k(g["cfg"], g["sd"], g["queries"], g["items"], g["item_ids"], g["k"], chunk=2)
    assert (scores[:, :: g["col_stride"]] - g["ref_scores_strided"]).abs().max().item() <= 2e-6
    r = O.compare_top_k(g["ref_top_scores"], g["ref_top_ids"], scores, g["item_ids"], g["k"], 1e-5, 1e-5)
    assert r["ok"] == 1.0, r
    assert torch.equal(top_ids, g["ref_top_ids"])
