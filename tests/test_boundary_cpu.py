"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mol_b200.h declares, the ctypes structs match the header, the torch modules keep the
reference's state-dict keys, and argument errors map to the exception types the reference raises.
No compute entry point is called here (no GPU in this container)."""
import ctypes
import os
import re

import pytest
import torch

from rails_b200 import _lib, engine
from tests.golden_util import golden_names, load_golden
from tests.helpers import CFG_8x8x32, build_module, factory_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, "include", "mol_b200.h")) as f:
        return f.read()


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = set(re.findall(r"\b(mol_[a-z0-9_]+)\s*\(", _header()))
    declared -= {"mol_shape", "mol_weights", "mol_index"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.mol_version()


def test_default_build_reports_the_shipped_kernel_knobs():
    """The in-tree library is the shipped configuration: the experimental activation / sharing knobs are off (tuning
    variants are separate files selected with MOL_B200_LIB and report their own knobs)."""
    if os.environ.get("MOL_B200_LIB"):
        pytest.skip("a tuning variant is loaded")
    k = _lib.build_knobs()
    assert k == {"e2poly": 0, "e2h2": 0x3E, "hidf16": 1}


def test_constants_match_header():
    h = _header()
    assert int(re.search(r"#define MOL_MAX_K (\d+)", h).group(1)) == _lib.MOL_MAX_K
    assert int(re.search(r"#define MOL_MAX_UID_TABLES (\d+)", h).group(1)) == _lib.MOL_MAX_UID_TABLES
    for name in ("MOL_ERR_INVALID", "MOL_ERR_CUDA", "MOL_ERR_WORKSPACE", "MOL_ERR_RANGE"):
        assert int(re.search(rf"#define {name} (\d+)", h).group(1)) == getattr(_lib, name)


def test_struct_layouts():
    assert ctypes.sizeof(_lib.MolShape) == 4 * (11 + 4 + 1) + 8
    assert ctypes.sizeof(_lib.MolWeights) == 8 * (16 + 4 + 1)  # + `prepared`
    assert ctypes.sizeof(_lib.MolIndex) == 8 * 8


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_keys_equal_reference(name):
    g = load_golden(name)
    mol, _ = build_module(g["cfg"])
    assert list(mol.state_dict().keys()) == list(g["sd"].keys()) or set(mol.state_dict().keys()) == set(g["sd"].keys())
    mol.load_state_dict(g["sd"], strict=True)
    for k, v in g["sd"].items():
        assert mol.state_dict()[k].shape == v.shape


def test_debug_string_matches_reference_checkpoint_name():
    from tests.helpers import CFG_8x4x64

    _, s = build_module(CFG_8x4x64)
    # the ML-1M checkpoint file name embeds this string
    assert s == (
        "MoL-8x4x64-t0.05-d0.2-l2-q512d0.0swiglu-id0.1-gq128-gi128d0.0-gqi128d0.0-x-glu_silu-uids6040d0.5"
    )


def test_shape_struct_from_module():
    mol, _ = build_module(CFG_8x8x32)
    s = mol.mol_shape()
    assert (s.query_embedding_dim, s.item_embedding_dim, s.dot_product_dimension) == (64, 64, 32)
    assert (s.query_dot_product_groups, s.item_dot_product_groups) == (8, 8)
    assert (s.query_hidden_dim, s.gating_qi_hidden_dim, s.query_nonlinearity, s.softmax_renorm) == (512, 128, 0, 1)
    assert abs(s.temperature - 0.05) < 1e-8


def test_workspace_and_index_bytes_are_pure_host_calls():
    lib = _lib.load()
    mol, _ = build_module(CFG_8x8x32)
    s = mol.mol_shape()
    n = ctypes.c_size_t()
    _lib.check(lib.mol_index_bytes(ctypes.byref(s), 1000, ctypes.byref(n)))
    # fp32 X_sub + GI, bf16 copies padded to 128 rows
    assert n.value >= 1000 * (256 + 64) * 4 + 1024 * (256 + 64) * 2
    _lib.check(lib.mol_search_workspace_bytes(ctypes.byref(s), 1000, 4, 10, _lib.MODE_EXACT, ctypes.byref(n)))
    assert n.value >= 4 * 1000 * 4


def test_unsupported_shapes_raise_value_error():
    lib = _lib.load()
    mol, _ = build_module(CFG_8x8x32)
    s = mol.mol_shape()
    s.gating_qi_hidden_dim = 96
    with pytest.raises(ValueError, match="gating_qi_hidden_dim"):
        _lib.check(lib.mol_shape_check(ctypes.byref(s), None))
    kw = factory_kwargs(CFG_8x8x32)
    kw["gating_combination_type"] = "none"
    from rails_b200.modeling.similarity_utils import create_mol_interaction_module

    m, _ = create_mol_interaction_module(**kw)
    with pytest.raises(ValueError, match="glu_silu"):
        m.mol_shape()
    kw = factory_kwargs(CFG_8x8x32)
    kw["item_hidden_dim"] = 256
    with pytest.raises(ValueError):
        create_mol_interaction_module(**kw)


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: CPU inputs raise instead of being scored somewhere else."""
    from rails_b200.indexing.mol_top_k import MoLBruteForceTopK

    mol, _ = build_module(CFG_8x8x32)
    items = torch.randn(1, 10, 64)
    ids = torch.arange(10).unsqueeze(0)
    top = MoLBruteForceTopK(mol, items, ids)  # construction keeps references only
    with pytest.raises(RuntimeError, match="CUDA"):
        top(torch.randn(2, 64), k=3)
    with pytest.raises(RuntimeError, match="CUDA"):
        mol(torch.randn(2, 64), items)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_product_code_never_imports_the_oracle():
    for pkg in (os.path.join(ROOT, "rails_b200"), os.path.join(ROOT, "rails"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(pkg):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, f
                    assert "tests." not in src or f == "workloads.py", f


def test_rails_alias_package_is_the_same_modules():
    """the replacement classes are importable under the reference's module paths."""
    import rails  # noqa: F401
    from rails.indexing.mol_top_k import MoLBruteForceTopK
    from rails.similarities.mol.similarity_fn import MoLSimilarity
    from rails_b200.indexing import mol_top_k
    from rails_b200.similarities.mol import similarity_fn

    assert MoLBruteForceTopK is mol_top_k.MoLBruteForceTopK and MoLSimilarity is similarity_fn.MoLSimilarity


def test_get_top_k_module_names_match_the_reference():
    """indexing/utils_rails.py:25-233 of the reference: same method names -> same classes / constructor arguments."""
    import types

    import pytest
    import torch

    from rails_b200.indexing import mol_top_k as M
    from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
    from rails_b200.indexing.utils_rails import get_top_k_module
    from tests.helpers import CFG_8x8x32, build_module

    mol, _ = build_module(CFG_8x8x32, None, "cpu")
    model = types.SimpleNamespace(_ndp_module=mol)
    items, ids = torch.randn(1, 300, 64), torch.arange(1, 301).unsqueeze(0)
    assert isinstance(get_top_k_module("MIPSBruteForceTopK", model, items, ids), MIPSBruteForceTopK)
    assert isinstance(get_top_k_module("MoLBruteForceTopK", model, items, ids), M.MoLBruteForceTopK)
    t = get_top_k_module("MoLNaiveTopK25", model, items, ids)
    assert isinstance(t, M.MoLNaiveTopK) and t._k_per_group == 25 and t.mol_module is mol
    t = get_top_k_module("MoLAvgTopK2500", model, items, ids)
    assert isinstance(t, M.MoLAvgTopK) and t._avg_top_k == 2500
    t = get_top_k_module("MoLCombTopK50_1000", model, items, ids)
    assert isinstance(t, M.MoLCombTopK) and (t._k_per_group, t._avg_top_k) == (50, 1000)
    for bad in ("MoLNaiveTopK7", "MoLAvgTopK123", "MoLCombTopK5_1000", "FaissTopK", ""):
        with pytest.raises(ValueError, match="Invalid top-k method"):
            get_top_k_module(bad, model, items, ids)
    with pytest.raises(NotImplementedError):
        get_top_k_module("MoLNaiveFaissTopK5", model, items, ids)


def test_round2_entry_points_host_side_contracts(monkeypatch):
    """Host-only behaviour of the entry points added around the path: workspace sizes are pure host
    calls that account for the streaming buffers, and argument errors come back as status codes before any CUDA work."""
    from ctypes import byref, c_size_t

    lib = _lib.load()
    mol, _ = build_module(CFG_8x8x32, None, "cpu", seed=0)
    shape = mol.mol_shape()
    a, b, c = c_size_t(), c_size_t(), c_size_t()
    # exclusion lists: the workspace grows with the list (internal over-fetch of the fp32 paths), never shrinks
    assert lib.mol_search_workspace_bytes(byref(shape), 1_000_000, 64, 100, _lib.MODE_AUTO, byref(a)) == 0
    assert lib.mol_search_excluding_workspace_bytes(byref(shape), 1_000_000, 64, 100, 211, _lib.MODE_AUTO, byref(b)) == 0
    assert b.value >= a.value
    assert lib.mol_search_excluding_workspace_bytes(byref(shape), 1_000_000, 64, 100, 0, _lib.MODE_AUTO, byref(c)) == 0
    assert c.value == a.value
    # MIPS: large corpora carry the streaming path's buffers on top of the (rows, N) matrix, and MOL_B200_DOTFILTER=0
    # drops them again
    assert lib.mol_mips_workspace_bytes(1_000_000, 64, 100, byref(a)) == 0
    monkeypatch.setenv("MOL_B200_DOTFILTER", "0")
    assert lib.mol_mips_workspace_bytes(1_000_000, 64, 100, byref(b)) == 0
    monkeypatch.delenv("MOL_B200_DOTFILTER")
    assert a.value > b.value >= 64 * 1_000_000 * 4
    assert lib.mol_mips_workspace_bytes(10_000, 64, 100, byref(c)) == 0 and c.value < a.value
    # argument errors: status + message, no CUDA call needed
    assert lib.mol_mips_search_cached(None, None, None, 100, 64, 4, 101, None, None, None, None, 0, None) == _lib.MOL_ERR_RANGE
    assert b"out of range" in lib.mol_last_error()
    assert lib.mol_select_valid(None, None, None, 4, 10, 3, 20, None, None, None) == _lib.MOL_ERR_INVALID  # k' < k
