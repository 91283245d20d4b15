"""ctypes wrapper of oracle/mol_oracle.c (plain-C restatement of the reference's MoL scoring).  TEST INFRASTRUCTURE ONLY:
imported by tests/ (and built by __graft_entry__.build()); nothing under rails_b200/ uses it.

Same call shape as oracle.mol_oracle.similarity: (cfg, state dict, queries (B, D), items (N, D), user_ids) -> (B, N) fp32.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np
import torch

from oracle import mol_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_c", "libmol_oracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc via oracle/Makefile (seconds)."""
    src = os.path.join(HERE, "mol_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        r = subprocess.run(["make", "-C", HERE, "-B"], capture_output=True, text=True)
        if r.returncode != 0:  # a gcc without libgomp: single-threaded build
            r = subprocess.run(["make", "-C", HERE, "-B", "OPENMP="], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building the C oracle failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.molc_version.restype = ctypes.c_char_p
    return _lib


def _f(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy())


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def similarity(cfg: O.MoLConfig, sd: Dict[str, torch.Tensor], queries: torch.Tensor, items: torch.Tensor,
               user_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = load()
    q, it = _f(queries), _f(items.reshape(-1, items.size(-1)))
    B, N = q.shape[0], it.shape[0]
    PQ, PX, d, L = cfg.query_dot_product_groups, cfg.item_dot_product_groups, cfg.dot_product_dimension, cfg.num_logits
    w = {k: _f(v) for k, v in sd.items() if v.is_floating_point()}
    Hq = w[O.K_Q_GLU_W].shape[1] // 2
    Hgq, Hgi, H = w[O.K_GQ_W1].shape[0], w[O.K_GI_W1].shape[0], w[O.K_QI_W1].shape[0]
    u = len(cfg.uid_embedding_hash_sizes)
    xsub = np.empty((N, PX, d), np.float32)
    gi = np.empty((N, L), np.float32)
    lib.molc_item_side(
        ctypes.c_int64(N), it.shape[1], PX, d, L, Hgi, _p(it), _p(w[O.K_X_W]), _p(w[O.K_X_B]), _p(w[O.K_GI_W1]),
        _p(w[O.K_GI_B1]), _p(w[O.K_GI_W2]), ctypes.c_float(cfg.eps), _p(xsub), _p(gi),
    )
    qsub = np.empty((B, PQ, d), np.float32)
    gq = np.empty((B, L), np.float32)
    hashes = (ctypes.c_int32 * max(u, 1))(*cfg.uid_embedding_hash_sizes)
    tables = [w[O.K_UID.format(i)] for i in range(u)]
    table_ptrs = (ctypes.c_void_p * max(u, 1))(*[t.ctypes.data for t in tables])
    uid = None
    if u > 0:
        uid = np.ascontiguousarray(user_ids.detach().cpu().to(torch.int64).numpy())
    lib.molc_query_side(
        B, q.shape[1], PQ, d, L, Hq, Hgq, 0 if cfg.query_nonlinearity == "geglu" else 1, u, hashes, table_ptrs, _p(uid),
        _p(q), _p(w[O.K_Q_GLU_W]), _p(w[O.K_Q_GLU_B]), _p(w[O.K_Q_OUT_W]), _p(w[O.K_Q_OUT_B]), _p(w[O.K_GQ_W1]),
        _p(w[O.K_GQ_B1]), _p(w[O.K_GQ_W2]), ctypes.c_float(cfg.eps), _p(qsub), _p(gq),
    )
    scores = np.empty((B, N), np.float32)
    lib.molc_scores(
        B, ctypes.c_int64(N), PQ, PX, d, H, _p(qsub), _p(xsub), _p(gq), _p(gi), _p(w[O.K_QI_W1]), _p(w[O.K_QI_B1]),
        _p(w[O.K_QI_W2]), _p(w[O.K_QI_B2]), ctypes.c_float(cfg.temperature), ctypes.c_float(cfg.eps),
        1 if cfg.softmax_dropout_rate > 0.0 else 0, _p(scores),
    )
    return torch.from_numpy(scores)
