"""ctypes binding of libmol_b200.so (include/mol_b200.h).

This is the binding a maintainer of the reference would add: the reference path
(rails/indexing/mol_top_k.py:99-130 -> rails/similarities/mol/similarity_fn.py:341-413) is pure
PyTorch, so the FFI is Python -> C ABI.  There is NO fallback: if the shared library is missing or
fails to load, importing any product module that needs it raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional

MOL_OK = 0
MOL_ERR_INVALID = 1
MOL_ERR_CUDA = 2
MOL_ERR_WORKSPACE = 3
MOL_ERR_RANGE = 4
MOL_MAX_UID_TABLES = 4
MOL_MAX_K = 8192
MODE_AUTO, MODE_EXACT, MODE_TENSOR = 0, 1, 2

# MOL_B200_LIB selects another build of the same library (tuning variants made by `python -m rails_b200.build --variant`)
LIB_PATH = os.environ.get("MOL_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmol_b200.so")

# every symbol include/mol_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = (
    "mol_version", "mol_last_error", "mol_shape_check", "mol_index_bytes", "mol_index_layout",
    "mol_index_build_workspace_bytes", "mol_index_build", "mol_search_workspace_bytes", "mol_search",
    "mol_search_host", "mol_score_all", "mol_score_all_coarse", "mol_query_prologue", "mol_merge_topk_workspace_bytes",
    "mol_merge_topk", "mol_topk_workspace_bytes", "mol_topk", "mol_launch_count", "mol_launch_count_reset",
    "mol_profile_enable", "mol_profile_collect", "mol_select_valid", "mol_mips_workspace_bytes", "mol_mips_search", "mol_mips_search_cached", "mol_search_excluding_workspace_bytes", "mol_search_excluding",
    "mol_dot_scores", "mol_index_avg_embeddings", "mol_search_avg_workspace_bytes", "mol_search_avg",
    "mol_search_groups_workspace_bytes", "mol_search_groups",
    "mol_weights_prepared_bytes", "mol_weights_prepare", "mol_search_stats",
    "mol_pack_topk", "mol_merge_topk_packed_workspace_bytes", "mol_merge_topk_packed",
)
MOL_PACKED_ENTRY_BYTES = 16
NUM_STATS = 8
STAT_NAMES = ("fallback_queries", "filter_overflows", "max_survivors", "filter_strategy", "tensor_path",
              "second_chance_queries", "k_prime", "survivor_capacity")


class MolShape(ctypes.Structure):
    _fields_ = [
        ("query_embedding_dim", c_int32),
        ("item_embedding_dim", c_int32),
        ("dot_product_dimension", c_int32),
        ("query_dot_product_groups", c_int32),
        ("item_dot_product_groups", c_int32),
        ("query_hidden_dim", c_int32),
        ("gating_query_hidden_dim", c_int32),
        ("gating_item_hidden_dim", c_int32),
        ("gating_qi_hidden_dim", c_int32),
        ("query_nonlinearity", c_int32),
        ("num_uid_tables", c_int32),
        ("uid_hash_sizes", c_int32 * MOL_MAX_UID_TABLES),
        ("softmax_renorm", c_int32),
        ("temperature", c_float),
        ("eps", c_float),
    ]


class MolWeights(ctypes.Structure):
    _fields_ = [
        ("q_glu_w", c_void_p), ("q_glu_b", c_void_p), ("q_out_w", c_void_p), ("q_out_b", c_void_p),
        ("uid_emb", c_void_p * MOL_MAX_UID_TABLES),
        ("x_w", c_void_p), ("x_b", c_void_p),
        ("gq_w1", c_void_p), ("gq_b1", c_void_p), ("gq_w2", c_void_p),
        ("gi_w1", c_void_p), ("gi_b1", c_void_p), ("gi_w2", c_void_p),
        ("qi_w1", c_void_p), ("qi_b1", c_void_p), ("qi_w2", c_void_p), ("qi_b2", c_void_p),
        ("prepared", c_void_p),
    ]


class MolIndex(ctypes.Structure):
    _fields_ = [
        ("num_items", c_int64),
        ("raw_items", c_void_p),
        ("item_ids", c_void_p),
        ("xsub_f32", c_void_p),
        ("gi_f32", c_void_p),
        ("xsub_half", c_void_p),
        ("gi_half", c_void_p),
        ("half_overflow", c_void_p),
    ]


_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Loads the CUDA library; raises RuntimeError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m rails_b200.build` "
            "(rails_b200 has no CPU / PyTorch fallback for the MoL top-k path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    P = POINTER
    lib.mol_version.restype = c_char_p
    lib.mol_last_error.restype = c_char_p
    lib.mol_launch_count.restype = c_int64
    lib.mol_launch_count_reset.restype = None
    lib.mol_shape_check.argtypes = [P(MolShape), P(c_int32)]
    lib.mol_index_bytes.argtypes = [P(MolShape), c_int64, P(c_size_t)]
    lib.mol_index_layout.argtypes = [P(MolShape), c_int64, c_void_p, c_void_p, c_void_p, c_size_t, P(MolIndex)]
    lib.mol_index_build_workspace_bytes.argtypes = [P(MolShape), c_int64, P(c_size_t)]
    lib.mol_index_build.argtypes = [P(MolShape), P(MolWeights), P(MolIndex), c_void_p, c_size_t, c_void_p]
    lib.mol_search_workspace_bytes.argtypes = [P(MolShape), c_int64, c_int32, c_int32, c_int32, P(c_size_t)]
    lib.mol_search.argtypes = [
        P(MolShape), P(MolWeights), P(MolIndex), c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_search_host.argtypes = lib.mol_search.argtypes
    lib.mol_score_all.argtypes = [
        P(MolShape), P(MolWeights), P(MolIndex), c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_score_all_coarse.argtypes = lib.mol_score_all.argtypes
    lib.mol_query_prologue.argtypes = [
        P(MolShape), P(MolWeights), c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_merge_topk_workspace_bytes.argtypes = [c_int32, c_int32, c_int32, P(c_size_t)]
    lib.mol_merge_topk.argtypes = [
        c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_topk_workspace_bytes.argtypes = [c_int64, c_int32, c_int32, P(c_size_t)]
    lib.mol_topk.argtypes = [
        c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_select_valid.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
    ]
    lib.mol_mips_workspace_bytes.argtypes = [c_int64, c_int32, c_int32, P(c_size_t)]
    lib.mol_mips_search.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_mips_search_cached.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
        c_void_p,
    ]
    lib.mol_dot_scores.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]
    lib.mol_index_avg_embeddings.argtypes = [P(MolShape), P(MolIndex), c_void_p, c_void_p]
    lib.mol_search_avg_workspace_bytes.argtypes = [P(MolShape), c_int64, c_int32, c_int32, c_int32, P(c_size_t)]
    lib.mol_search_avg.argtypes = [
        P(MolShape), P(MolWeights), P(MolIndex), c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_search_groups_workspace_bytes.argtypes = [P(MolShape), c_int64, c_int32, c_int32, c_int32, P(c_size_t)]
    lib.mol_search_groups.argtypes = [
        P(MolShape), P(MolWeights), P(MolIndex), c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_weights_prepared_bytes.argtypes = [P(MolShape), P(c_size_t)]
    lib.mol_weights_prepare.argtypes = [P(MolShape), P(MolWeights), c_void_p, c_size_t, c_void_p]
    lib.mol_search_excluding_workspace_bytes.argtypes = [P(MolShape), c_int64, c_int32, c_int32, c_int32, c_int32, P(c_size_t)]
    lib.mol_search_excluding.argtypes = [
        P(MolShape), P(MolWeights), P(MolIndex), c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32,
        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]
    lib.mol_search_stats.argtypes = [c_void_p, P(c_int32), c_void_p]
    lib.mol_pack_topk.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.mol_merge_topk_packed_workspace_bytes.argtypes = [c_int32, c_int32, c_int32, P(c_size_t)]
    lib.mol_merge_topk_packed.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.mol_profile_enable.argtypes = [c_int32]
    lib.mol_profile_enable.restype = None
    lib.mol_profile_collect.argtypes = [P(ctypes.c_double), P(c_int32)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("mol_version", "mol_last_error", "mol_launch_count", "mol_launch_count_reset", "mol_profile_enable"):
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def build_knobs() -> dict:
    """Compile-time tuning knobs of the loaded library, parsed from mol_version()'s "[name=value ...]" suffix
    (ints; e.g. {"e2poly": 14, "e2h2": 0, ...}).  A tuning variant loaded through MOL_B200_LIB reports its own."""
    v = load().mol_version().decode()
    out = {}
    if "[" in v:
        for tok in v[v.index("[") + 1 : v.rindex("]")].split():
            name, _, val = tok.partition("=")
            out[name] = int(val, 0)
    return out


def check(status: int) -> None:
    """Maps the C status to the exception type the reference raises at this boundary."""
    if status == MOL_OK:
        return
    msg = load().mol_last_error().decode("utf-8", "replace")
    if status == MOL_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


__all__ = [
    "MolShape", "MolWeights", "MolIndex", "load", "check", "build_knobs", "byref", "c_size_t", "c_int32", "c_void_p",
    "MODE_AUTO", "MODE_EXACT", "MODE_TENSOR", "MOL_MAX_K", "EXPORTED_SYMBOLS", "LIB_PATH", "MOL_PACKED_ENTRY_BYTES",
    "NUM_STATS", "STAT_NAMES",
]
