"""`get_top_k_module`, mirroring the reference's indexing/utils_rails.py:25-233: the same method names map to the same
constructor arguments, an unknown name raises the same ValueError.  The reference spells the mapping as an if-chain;
here the names are parsed (family + numeric suffix) and checked against the set of names the reference accepts."""
import re

import torch

from rails_b200.indexing.candidate_index import TopKModule
from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLBruteForceTopK, MoLCombTopK, MoLNaiveTopK

_NAIVE_K = (5, 10, 25, 50, 75, 100)  # utils_rails.py:37-91
_AVG_K = (100, 200, 500, 1000, 2000, 2500, 3000, 4000)  # :92-147
_COMB = ((5, 100), (5, 200), (5, 500), (1, 100), (10, 100), (1, 500), (10, 500), (50, 500), (50, 1000), (100, 1000))  # :148-230


def get_top_k_module(top_k_method: str, model: torch.nn.Module, item_embeddings: torch.Tensor, item_ids: torch.Tensor) -> TopKModule:
    if top_k_method == "MIPSBruteForceTopK":
        return MIPSBruteForceTopK(item_embeddings=item_embeddings, item_ids=item_ids)
    mol = dict(mol_module=model._ndp_module, item_embeddings=item_embeddings, item_ids=item_ids)
    if top_k_method == "MoLBruteForceTopK":
        return MoLBruteForceTopK(**mol)
    if top_k_method == "MoLNaiveFaissTopK5":
        return MoLNaiveTopK(k_per_group=5, use_faiss=True, **mol)  # raises: FAISS is outside the B200 hot path
    m = re.fullmatch(r"MoLNaiveTopK(\d+)", top_k_method)
    if m and int(m.group(1)) in _NAIVE_K:
        return MoLNaiveTopK(k_per_group=int(m.group(1)), **mol)
    m = re.fullmatch(r"MoLAvgTopK(\d+)", top_k_method)
    if m and int(m.group(1)) in _AVG_K:
        return MoLAvgTopK(avg_top_k=int(m.group(1)), **mol)
    m = re.fullmatch(r"MoLCombTopK(\d+)_(\d+)", top_k_method)
    if m and (int(m.group(1)), int(m.group(2))) in _COMB:
        return MoLCombTopK(k_per_group=int(m.group(1)), avg_top_k=int(m.group(2)), **mol)
    raise ValueError(f"Invalid top-k method {top_k_method}")
