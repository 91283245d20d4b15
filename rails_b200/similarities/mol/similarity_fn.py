"""Mixture-of-Logits similarity, B200-native.

Mirrors the public surface of the reference's rails/similarities/mol/similarity_fn.py
(`SoftmaxDropoutCombiner` :66-96, `MoLGatingFn` :99-201, `MoLSimilarity` :204-413): same constructor
arguments, same sub-module attribute names, hence the same state-dict keys — reference checkpoints
load with strict=True.  The modules below hold parameters only; the arithmetic

    score[b,x] = sum_l softmax_l(silu(GQ[b]*GI[x] + W2 silu(W1 l[b,x] + b1) + b2)) * l[b,x,l],
    l[b,x,n*P_X+m] = <Q_sub[b,n], X_sub[x,m]> / tau

runs in libmol_b200.so (csrc/mol_exact.cu for exact fp32 scores, csrc/mol_coarse_sm100.cu for the
tcgen05 pass used by MoLBruteForceTopK).  Inference (eval-mode) semantics only: dropout is identity,
the eval-time renormalisation quirk of similarity_fn.py:43-45 is kept, aux losses are empty.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from rails_b200 import _lib, engine
from rails_b200._lib import MolShape
from rails_b200.similarities.layers import GeGLU, SwiGLU
from rails_b200.similarities.module import SimilarityModule
from rails_b200.similarities.mol.embeddings_fn import MoLEmbeddingsFn


class SoftmaxDropoutCombiner(torch.nn.Module):
    """Holder of (dropout_rate, eps) — reference similarity_fn.py:66-96.  No parameters."""

    def __init__(self, dropout_rate: float, eps: float) -> None:
        super().__init__()
        self._dropout_rate: float = dropout_rate
        self._eps: float = eps

    def forward(self, gating_weights, x):  # pragma: no cover
        raise RuntimeError("the softmax combiner is fused into the CUDA scoring kernels")


class MoLGatingFn(torch.nn.Module):
    """Parameter container of the gating MLPs — reference similarity_fn.py:99-201."""

    def __init__(
        self,
        num_logits: int,
        query_embedding_dim: int,
        item_embedding_dim: int,
        query_only_partial_fn: Optional[Callable[[int, int], torch.nn.Module]],
        item_only_partial_fn: Optional[Callable[[int, int], torch.nn.Module]],
        qi_partial_fn: Optional[Callable[[int, int], torch.nn.Module]],
        combination_type: str,
        normalization_fn: Callable[[int], torch.nn.Module],
    ) -> None:
        super().__init__()
        self._query_only_partial_module = (
            query_only_partial_fn(query_embedding_dim, num_logits) if query_only_partial_fn else None
        )
        self._item_only_partial_module = (
            item_only_partial_fn(item_embedding_dim, num_logits) if item_only_partial_fn else None
        )
        self._qi_partial_module = qi_partial_fn(num_logits, num_logits) if qi_partial_fn is not None else None
        if (
            self._query_only_partial_module is None
            and self._item_only_partial_module is None
            and self._qi_partial_module is None
        ):
            raise ValueError(
                "At least one of query_only_partial_fn, item_only_partial_fn, and qi_partial_fn must not be None."
            )
        self._num_logits: int = num_logits
        self._combination_type: str = combination_type
        self._normalization_fn: torch.nn.Module = normalization_fn(num_logits)

    def forward(self, logits, query_embeddings, item_embeddings):  # pragma: no cover
        raise RuntimeError("the gating function is fused into the CUDA scoring kernels")


class MoLSimilarity(SimilarityModule):
    def __init__(
        self,
        query_embedding_dim: int,
        item_embedding_dim: int,
        dot_product_dimension: int,
        query_dot_product_groups: int,
        item_dot_product_groups: int,
        temperature: float,
        dot_product_l2_norm: bool,
        query_embeddings_fn: MoLEmbeddingsFn,
        item_embeddings_fn: Optional[MoLEmbeddingsFn],
        item_proj_fn: Optional[Callable[[int, int], torch.nn.Module]],
        gating_query_only_partial_fn: Optional[Callable[[int, int], torch.nn.Module]],
        gating_item_only_partial_fn: Optional[Callable[[int, int], torch.nn.Module]],
        gating_qi_partial_fn: Optional[Callable[[int], torch.nn.Module]],
        gating_combination_type: str,
        gating_normalization_fn: Callable[[int], torch.nn.Module],
        eps: float,
        apply_query_embeddings_fn: bool = True,
        apply_item_embeddings_fn: bool = True,
        autocast_bf16: bool = False,
    ) -> None:
        super().__init__()
        self._gating_fn: MoLGatingFn = MoLGatingFn(
            num_logits=query_dot_product_groups * item_dot_product_groups,
            query_embedding_dim=query_embedding_dim,
            item_embedding_dim=item_embedding_dim,
            query_only_partial_fn=gating_query_only_partial_fn,
            item_only_partial_fn=gating_item_only_partial_fn,
            qi_partial_fn=gating_qi_partial_fn,
            combination_type=gating_combination_type,
            normalization_fn=gating_normalization_fn,
        )
        self._query_embeddings_fn: MoLEmbeddingsFn = query_embeddings_fn
        self._item_embeddings_fn: Optional[MoLEmbeddingsFn] = item_embeddings_fn
        if item_embeddings_fn is None:
            raise ValueError(
                "rails_b200 does not implement the reference's deprecated `item_proj_fn` legacy path "
                "(similarity_fn.py:252-259); pass item_embeddings_fn"
            )
        self._item_proj_module = None
        self._apply_query_embeddings_fn: bool = apply_query_embeddings_fn
        self._apply_item_embeddings_fn: bool = apply_item_embeddings_fn
        self._dot_product_l2_norm: bool = dot_product_l2_norm
        self._query_dot_product_groups: int = query_dot_product_groups
        self._item_dot_product_groups: int = item_dot_product_groups
        self._dot_product_dimension: int = dot_product_dimension
        self._temperature: float = temperature
        self._eps: float = eps
        self._autocast_bf16: bool = autocast_bf16
        # engine-side caches (not part of the state dict)
        self._packed: Optional[engine.PackedWeights] = None
        self._packed_key = None
        self._workspaces: Dict[torch.device, engine.Workspace] = {}
        self._index_cache = None  # (item tensor, key, IndexHandle) of the last forward()

    # ------------------------------------------------------------------ engine plumbing
    def mol_shape(self) -> MolShape:
        """The C-ABI shape struct of this head; raises ValueError for configurations outside the path."""
        if not self._apply_query_embeddings_fn or not self._apply_item_embeddings_fn:
            raise ValueError("rails_b200 MoLSimilarity requires apply_{query,item}_embeddings_fn=True")
        if not self._dot_product_l2_norm:
            raise ValueError("rails_b200 MoLSimilarity requires dot_product_l2_norm=True (all reference configs)")
        g = self._gating_fn
        if g._combination_type != "glu_silu":
            raise ValueError(f"gating_combination_type {g._combination_type!r} is not supported (only 'glu_silu')")
        if g._query_only_partial_module is None or g._item_only_partial_module is None or g._qi_partial_module is None:
            raise ValueError("rails_b200 MoLSimilarity requires all three gating partial modules")
        sd = self.state_dict()
        for key in engine._FIELD_OF_KEY:
            if key not in sd:
                raise ValueError(f"unsupported MoL structure: parameter {key} not found")
        glu = self._query_embeddings_fn._query_emb_proj_module[1]
        if not isinstance(glu, (GeGLU, SwiGLU)):
            raise ValueError("query projection must be Sequential(Dropout, GeGLU|SwiGLU, Linear) (query_hidden_dim > 0)")
        hashes = list(getattr(self._query_embeddings_fn, "_uid_embedding_hash_sizes", []))
        s = MolShape()
        s.query_embedding_dim = sd[engine.K_Q_GLU_W].size(0)
        s.item_embedding_dim = sd[engine.K_X_W].size(1)
        s.dot_product_dimension = self._dot_product_dimension
        s.query_dot_product_groups = self._query_dot_product_groups
        s.item_dot_product_groups = self._item_dot_product_groups
        s.query_hidden_dim = sd[engine.K_Q_GLU_W].size(1) // 2
        s.gating_query_hidden_dim = sd[engine.K_GQ_W1].size(0)
        s.gating_item_hidden_dim = sd[engine.K_GI_W1].size(0)
        s.gating_qi_hidden_dim = sd[engine.K_QI_W1].size(0)
        s.query_nonlinearity = glu.kind
        s.num_uid_tables = len(hashes)
        if len(hashes) > _lib.MOL_MAX_UID_TABLES:
            raise ValueError(f"at most {_lib.MOL_MAX_UID_TABLES} uid embedding tables are supported")
        for i, h in enumerate(hashes):
            s.uid_hash_sizes[i] = int(h)
        s.softmax_renorm = 1 if g._normalization_fn._dropout_rate > 0.0 else 0
        s.temperature = float(self._temperature)
        s.eps = float(self._eps)
        L = s.query_dot_product_groups * s.item_dot_product_groups
        expect = {
            engine.K_Q_OUT_W: ((s.query_dot_product_groups - len(hashes)) * s.dot_product_dimension, s.query_hidden_dim),
            engine.K_X_W: (s.item_dot_product_groups * s.dot_product_dimension, s.item_embedding_dim),
            engine.K_GQ_W2: (L, s.gating_query_hidden_dim),
            engine.K_GI_W2: (L, s.gating_item_hidden_dim),
            engine.K_QI_W1: (s.gating_qi_hidden_dim, L),
            engine.K_QI_W2: (L, s.gating_qi_hidden_dim),
        }
        for key, shp in expect.items():
            if tuple(sd[key].shape) != shp:
                raise ValueError(f"parameter {key} has shape {tuple(sd[key].shape)}, expected {shp}")
        _lib.check(_lib.load().mol_shape_check(_lib.byref(s), None))
        return s

    def _weights_key(self, device: torch.device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def packed_weights(self, device: torch.device) -> engine.PackedWeights:
        key = self._weights_key(device)
        if self._packed is None or self._packed_key != key:
            self._packed = engine.PackedWeights(self.state_dict(), self.mol_shape(), device)
            self._packed_key = key
            self._index_cache = None
        return self._packed

    def invalidate_index_cache(self) -> None:
        """Drops the cached weights / item-side index (needed only after edits through `.data`, which `_version` misses)."""
        self._packed = None
        self._packed_key = None
        self._index_cache = None

    def workspace(self, device: torch.device) -> engine.Workspace:
        if device not in self._workspaces:
            self._workspaces[device] = engine.Workspace(device)
        return self._workspaces[device]

    def build_index(self, item_embeddings: torch.Tensor, item_ids: Optional[torch.Tensor]) -> engine.IndexHandle:
        """Item-side cache (X_sub, GI; fp32 + bf16) for an (N, D) corpus on a CUDA device."""
        engine._require_cuda(item_embeddings, "item_embeddings")
        return engine.IndexHandle(self.packed_weights(item_embeddings.device), item_embeddings, item_ids)

    # ------------------------------------------------------------------ reference API
    def get_query_component_embeddings(
        self, input_embeddings: torch.Tensor, decoupled_inference: bool = False, **kwargs
    ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """(B, D) -> (B, P_Q, d) l2-normalised.  Reference: similarity_fn.py:270-292."""
        if decoupled_inference and not self._apply_query_embeddings_fn:
            return input_embeddings, {}
        dev = input_embeddings.device
        qsub, _ = engine.query_prologue(
            self.packed_weights(dev), self.workspace(dev), input_embeddings, kwargs.get("user_ids")
        )
        return qsub.to(input_embeddings.dtype), {}

    def get_item_component_embeddings(
        self, input_embeddings: torch.Tensor, decoupled_inference: bool = False, **kwargs
    ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """(..., D) -> (..., P_X, d) l2-normalised.  Reference: similarity_fn.py:294-339."""
        if decoupled_inference and not self._apply_item_embeddings_fn:
            return input_embeddings, {}
        lead = input_embeddings.size()[:-1]
        flat = input_embeddings.reshape(-1, input_embeddings.size(-1))
        idx = self.build_index(flat, None)
        out = idx.xsub_f32().reshape(lead + (self._item_dot_product_groups, self._dot_product_dimension))
        return out.to(input_embeddings.dtype), {}

    @torch.no_grad()
    def forward(
        self, query_embeddings: torch.Tensor, item_embeddings: torch.Tensor, **kwargs
    ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """
        Args:
            query_embeddings: (B, D) x float on a CUDA device.
            item_embeddings: (1, X, D) x float (the brute-force branch, similarity_fn.py:389-396).
            kwargs: "user_ids" (B,) int64 is consumed when uid embeddings are configured; other keys
                (timestamps, ratings — data/eval.py:148) are accepted and ignored.
        Returns:
            ((B, X) similarity values in query_embeddings.dtype, {}).
        """
        if self.training:
            raise RuntimeError("rails_b200 MoLSimilarity implements inference only; call .eval()")
        if item_embeddings.dim() != 3 or item_embeddings.size(0) != 1:
            raise NotImplementedError(
                "only the (1, X, D) shared-candidates branch is implemented (per-row (B, X, D) candidates are "
                "the training-time path, out of scope: SURVEY.md §8)"
            )
        dev = query_embeddings.device
        weights = self.packed_weights(dev)
        # The item-side cache is reused while the caller passes the SAME tensor object, unmodified (in-place edits bump
        # `_version`; edits through `.data` do not - call `invalidate_index_cache()` after those).  The entry keeps a
        # strong reference to that tensor, so its address cannot be handed to another tensor while the entry lives.
        key = (item_embeddings._version, tuple(item_embeddings.shape), item_embeddings.dtype, self._packed_key)
        c = self._index_cache
        if c is None or c[0] is not item_embeddings or c[1] != key:
            c = (item_embeddings, key, engine.IndexHandle(weights, item_embeddings.squeeze(0), None))
            self._index_cache = c
        index = c[2]
        scores = engine.score_all(weights, index, self.workspace(dev), query_embeddings, kwargs.get("user_ids"))
        return scores.to(query_embeddings.dtype), {}
