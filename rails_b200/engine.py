"""Host-side glue between the torch modules and the C ABI (include/mol_b200.h).

PyTorch is used for device memory (caller-owned buffers handed to the library as raw pointers) and
for the current CUDA stream — nothing else.  Every arithmetic step of the path runs in
libmol_b200.so; there is no torch / CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from rails_b200 import _lib
from rails_b200._lib import MolIndex, MolShape, MolWeights, byref, c_int32, c_size_t

# state-dict keys of the reference MoLSimilarity (modeling/similarity_utils.py:72-214)
K_Q_GLU_W = "_query_embeddings_fn._query_emb_proj_module.1._w"
K_Q_GLU_B = "_query_embeddings_fn._query_emb_proj_module.1._b"
K_Q_OUT_W = "_query_embeddings_fn._query_emb_proj_module.2.weight"
K_Q_OUT_B = "_query_embeddings_fn._query_emb_proj_module.2.bias"
K_UID = "_query_embeddings_fn._uid_embeddings_{}.weight"
K_X_W = "_item_embeddings_fn._item_emb_proj_module.1.weight"
K_X_B = "_item_embeddings_fn._item_emb_proj_module.1.bias"
K_GQ_W1 = "_gating_fn._query_only_partial_module.0.weight"
K_GQ_B1 = "_gating_fn._query_only_partial_module.0.bias"
K_GQ_W2 = "_gating_fn._query_only_partial_module.2.weight"
K_GI_W1 = "_gating_fn._item_only_partial_module.1.weight"
K_GI_B1 = "_gating_fn._item_only_partial_module.1.bias"
K_GI_W2 = "_gating_fn._item_only_partial_module.3.weight"
K_QI_W1 = "_gating_fn._qi_partial_module.1.weight"
K_QI_B1 = "_gating_fn._qi_partial_module.1.bias"
K_QI_W2 = "_gating_fn._qi_partial_module.3.weight"
K_QI_B2 = "_gating_fn._qi_partial_module.3.bias"

_FIELD_OF_KEY = {
    K_Q_GLU_W: "q_glu_w", K_Q_GLU_B: "q_glu_b", K_Q_OUT_W: "q_out_w", K_Q_OUT_B: "q_out_b",
    K_X_W: "x_w", K_X_B: "x_b",
    K_GQ_W1: "gq_w1", K_GQ_B1: "gq_b1", K_GQ_W2: "gq_w2",
    K_GI_W1: "gi_w1", K_GI_B1: "gi_b1", K_GI_W2: "gi_w2",
    K_QI_W1: "qi_w1", K_QI_B1: "qi_b1", K_QI_W2: "qi_w2", K_QI_B2: "qi_b2",
}


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must live on a CUDA device: rails_b200 runs the MoL top-k path only through its "
            "sm_100a CUDA library (no CPU fallback)"
        )


class PackedWeights:
    """fp32, contiguous, device-resident views/copies of the module's parameters + the C structs."""

    def __init__(self, state: Dict[str, torch.Tensor], shape: MolShape, device: torch.device) -> None:
        device = torch.device(device)
        self.device = device
        self.shape = shape
        self.keep: List[torch.Tensor] = []
        self.struct = MolWeights()
        for key, field in _FIELD_OF_KEY.items():
            if key not in state:
                raise ValueError(f"MoL weights: missing parameter {key}")
            setattr(self.struct, field, self._dev(state[key]).data_ptr())
        for i in range(shape.num_uid_tables):
            key = K_UID.format(i)
            if key not in state:
                raise ValueError(f"MoL weights: missing parameter {key}")
            self.struct.uid_emb[i] = self._dev(state[key]).data_ptr()
        self.struct.prepared = None
        self.prepared: Optional[torch.Tensor] = None
        if device.type == "cuda":
            # operands that depend on the weights alone (transposed qi-MLP, fp16 UMMA images): once per weight version
            lib = _lib.load()
            nbytes = c_size_t()
            _lib.check(lib.mol_weights_prepared_bytes(byref(shape), byref(nbytes)))
            self.prepared = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device)
            base = (self.prepared.data_ptr() + 255) // 256 * 256
            with torch.cuda.device(device):
                _lib.check(lib.mol_weights_prepare(byref(shape), byref(self.struct), ctypes.c_void_p(base), nbytes.value,
                                                   _stream_ptr(device)))
            self.struct.prepared = base

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(t)
        return t


class Workspace:
    """Grow-only device scratch, owned by the caller side of the ABI."""

    def __init__(self, device: torch.device) -> None:
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        return self.buf


class IndexHandle:
    """The item-side cache of one corpus (shard): blob + C struct + borrowed caller tensors."""

    def __init__(self, weights: PackedWeights, raw_items: torch.Tensor, item_ids: Optional[torch.Tensor]) -> None:
        lib = _lib.load()
        _require_cuda(raw_items, "item_embeddings")
        if raw_items.dim() != 2:
            raise ValueError(f"raw items must be (N, D), got {tuple(raw_items.shape)}")
        if raw_items.size(1) != weights.shape.item_embedding_dim:
            raise ValueError(
                f"item_embeddings last dim {raw_items.size(1)} != item_embedding_dim {weights.shape.item_embedding_dim}"
            )
        self.weights = weights
        self.device = raw_items.device
        self.raw = raw_items.detach().to(torch.float32).contiguous()
        self.ids = None if item_ids is None else item_ids.detach().to(device=self.device, dtype=torch.int64).contiguous()
        self.N = int(self.raw.size(0))
        if self.ids is not None and self.ids.numel() != self.N:
            raise ValueError(f"item_ids has {self.ids.numel()} entries for {self.N} items")
        nbytes = c_size_t()
        _lib.check(lib.mol_index_bytes(byref(weights.shape), self.N, byref(nbytes)))
        self.blob = torch.empty(nbytes.value + 1024, dtype=torch.uint8, device=self.device)
        base = (self.blob.data_ptr() + 1023) // 1024 * 1024
        self.struct = MolIndex()
        _lib.check(
            lib.mol_index_layout(
                byref(weights.shape), self.N, _ptr(self.raw), _ptr(self.ids), ctypes.c_void_p(base),
                nbytes.value, byref(self.struct),
            )
        )
        ws_bytes = c_size_t()
        _lib.check(lib.mol_index_build_workspace_bytes(byref(weights.shape), self.N, byref(ws_bytes)))
        ws = torch.empty(max(ws_bytes.value, 1), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(
                lib.mol_index_build(
                    byref(weights.shape), byref(weights.struct), byref(self.struct), _ptr(ws), ws_bytes.value,
                    _stream_ptr(self.device),
                )
            )
        # `ws` may be freed after this point: the caching allocator is stream-ordered on the current stream.

    def xsub_f32(self) -> torch.Tensor:
        """(N, P_X, d) fp32 view of the cached item sub-embeddings (a copy, detached from the blob)."""
        s = self.weights.shape
        n = self.N * s.item_dot_product_groups * s.dot_product_dimension
        off = self.struct.xsub_f32 - self.blob.data_ptr()
        return (
            self.blob[off : off + 4 * n].view(torch.float32)
            .view(self.N, s.item_dot_product_groups, s.dot_product_dimension).clone()
        )


def search(
    weights: PackedWeights,
    index: IndexHandle,
    workspace: Workspace,
    queries: torch.Tensor,
    user_ids: Optional[torch.Tensor],
    k: int,
    sorted: bool = True,
    mode: int = _lib.MODE_AUTO,
    invalid_ids: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """mol_search with device buffers.  Returns (scores (B,k) fp32, ids (B,k) int64).  invalid_ids (B, N0) int64: ids
    excluded per query inside the search (mol_search_excluding; needs k + N0 <= N)."""
    lib = _lib.load()
    _require_cuda(queries, "query_embeddings")
    dev = index.device
    q = queries.detach().to(device=dev, dtype=torch.float32).contiguous()
    B = int(q.size(0))
    uid = None
    if weights.shape.num_uid_tables > 0:
        if user_ids is None:
            raise KeyError("user_ids")  # the reference indexes kwargs["user_ids"] (query_embeddings_fns.py:206)
        uid = user_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    nbytes = c_size_t()
    if invalid_ids is not None and invalid_ids.size(1) > 0:
        inv = invalid_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
        if inv.size(0) != B:
            raise ValueError(f"invalid_ids has {inv.size(0)} rows for {B} queries")
        n0 = int(inv.size(1))
        _lib.check(lib.mol_search_excluding_workspace_bytes(byref(weights.shape), index.N, B, k, n0, mode, byref(nbytes)))
        ws = workspace.get(nbytes.value)
        with torch.cuda.device(dev):
            _lib.check(
                lib.mol_search_excluding(
                    byref(weights.shape), byref(weights.struct), byref(index.struct), _ptr(q), _ptr(uid), B, k,
                    1 if sorted else 0, mode, _ptr(inv), n0, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(),
                    _stream_ptr(dev),
                )
            )
        return out_s, out_i
    _lib.check(lib.mol_search_workspace_bytes(byref(weights.shape), index.N, B, k, mode, byref(nbytes)))
    ws = workspace.get(nbytes.value)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_search(
                byref(weights.shape), byref(weights.struct), byref(index.struct), _ptr(q), _ptr(uid), B, k,
                1 if sorted else 0, mode, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(), _stream_ptr(dev),
            )
        )
    return out_s, out_i


class GraphedSearch:
    """One mol_search call of a fixed signature (B, k, mode, user_ids or not) captured in a CUDA graph.

    The C entry is stream-ordered, allocates nothing and never synchronises, so the ~25 launches of a search (query prologue,
    coarse passes, selections, rescoring, the idle fallback launches) are captured once and replayed as one graph launch:
    what is left of a small-batch search is kernel time.  Inputs are copied into the graph's static buffers; outputs are
    returned as fresh tensors.  The graph owns its workspace (the shared grow-only Workspace may be reallocated by other
    calls).  Build a new one when the weights or the index change."""

    def __init__(self, weights: PackedWeights, index: IndexHandle, B: int, k: int, mode: int, has_uid: bool) -> None:
        lib = _lib.load()
        dev = index.device
        self.weights, self.index, self.B, self.k, self.mode = weights, index, B, k, mode
        self.q = torch.zeros((B, weights.shape.query_embedding_dim), dtype=torch.float32, device=dev)
        self.uid = torch.zeros((B,), dtype=torch.int64, device=dev) if has_uid else None
        self.out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        self.out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
        nbytes = c_size_t()
        _lib.check(lib.mol_search_workspace_bytes(byref(weights.shape), index.N, B, k, mode, byref(nbytes)))
        self.workspace = Workspace(dev)
        self.ws = self.workspace.get(nbytes.value)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture (one-time function attributes, lazy module loading)
            self._call()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._call()

    def _call(self) -> None:
        lib = _lib.load()
        dev = self.index.device
        with torch.cuda.device(dev):
            _lib.check(
                lib.mol_search(
                    byref(self.weights.shape), byref(self.weights.struct), byref(self.index.struct), _ptr(self.q),
                    _ptr(self.uid), self.B, self.k, 1, self.mode, _ptr(self.out_s), _ptr(self.out_i), _ptr(self.ws),
                    self.ws.numel(), _stream_ptr(dev),
                )
            )

    def __call__(self, queries: torch.Tensor, user_ids: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        self.q.copy_(queries, non_blocking=True)
        if self.uid is not None:
            if user_ids is None:
                raise KeyError("user_ids")
            self.uid.copy_(user_ids, non_blocking=True)
        self.graph.replay()
        return self.out_s.clone(), self.out_i.clone()


def search_stats(workspace: Workspace) -> Dict[str, int]:
    """Counters of the last search that used `workspace` (mol_search_stats; synchronises the stream)."""
    lib = _lib.load()
    if workspace.buf is None:
        raise RuntimeError("no search has used this workspace yet")
    out = (c_int32 * _lib.NUM_STATS)()
    with torch.cuda.device(workspace.device):
        _lib.check(lib.mol_search_stats(_ptr(workspace.buf), out, _stream_ptr(workspace.device)))
    return dict(zip(_lib.STAT_NAMES, (int(v) for v in out)))


def search_host(
    weights: PackedWeights,
    index: IndexHandle,
    workspace: Workspace,
    host_queries: torch.Tensor,
    host_user_ids: Optional[torch.Tensor],
    k: int,
    host_out_scores: torch.Tensor,
    host_out_ids: torch.Tensor,
    mode: int = _lib.MODE_AUTO,
) -> None:
    """mol_search_host: pinned host buffers in/out, copies + sync inside the call (bench.py's e2e leg)."""
    lib = _lib.load()
    dev = index.device
    B = int(host_queries.size(0))
    assert host_queries.dtype == torch.float32 and host_queries.is_contiguous() and not host_queries.is_cuda
    nbytes = c_size_t()
    _lib.check(lib.mol_search_workspace_bytes(byref(weights.shape), index.N, B, k, mode, byref(nbytes)))
    ws = workspace.get(nbytes.value)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_search_host(
                byref(weights.shape), byref(weights.struct), byref(index.struct), _ptr(host_queries),
                _ptr(host_user_ids), B, k, 1, mode, _ptr(host_out_scores), _ptr(host_out_ids), _ptr(ws),
                ws.numel(), _stream_ptr(dev),
            )
        )


def score_all(
    weights: PackedWeights,
    index: IndexHandle,
    workspace: Workspace,
    queries: torch.Tensor,
    user_ids: Optional[torch.Tensor],
    coarse: bool = False,
) -> torch.Tensor:
    """(B, N) fp32 exact scores (MoLSimilarity.forward, B'==1 branch).  coarse=True returns the raw output of
    the tcgen05 coarse pass instead (diagnostic; approximate)."""
    lib = _lib.load()
    _require_cuda(queries, "query_embeddings")
    dev = index.device
    q = queries.detach().to(device=dev, dtype=torch.float32).contiguous()
    B = int(q.size(0))
    uid = None
    if weights.shape.num_uid_tables > 0:
        if user_ids is None:
            raise KeyError("user_ids")
        uid = user_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty((B, index.N), dtype=torch.float32, device=dev)
    nbytes = c_size_t()
    mode = _lib.MODE_TENSOR if coarse else _lib.MODE_EXACT
    _lib.check(lib.mol_search_workspace_bytes(byref(weights.shape), 1, B, 1, mode, byref(nbytes)))
    ws = workspace.get(nbytes.value)
    fn = lib.mol_score_all_coarse if coarse else lib.mol_score_all
    with torch.cuda.device(dev):
        _lib.check(
            fn(
                byref(weights.shape), byref(weights.struct), byref(index.struct), _ptr(q), _ptr(uid), B,
                _ptr(out), _ptr(ws), ws.numel(), _stream_ptr(dev),
            )
        )
    return out


def query_prologue(
    weights: PackedWeights, workspace: Workspace, queries: torch.Tensor, user_ids: Optional[torch.Tensor]
) -> Tuple[torch.Tensor, torch.Tensor]:
    """(Q_sub (B, P_Q, d), GQ (B, L)) fp32 from the CUDA query prologue."""
    lib = _lib.load()
    _require_cuda(queries, "query_embeddings")
    dev = queries.device
    s = weights.shape
    q = queries.detach().to(dtype=torch.float32).contiguous()
    B = int(q.size(0))
    uid = None
    if s.num_uid_tables > 0:
        if user_ids is None:
            raise KeyError("user_ids")
        uid = user_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
    L = s.query_dot_product_groups * s.item_dot_product_groups
    qsub = torch.empty((B, s.query_dot_product_groups, s.dot_product_dimension), dtype=torch.float32, device=dev)
    gq = torch.empty((B, L), dtype=torch.float32, device=dev)
    nbytes = c_size_t()
    _lib.check(lib.mol_search_workspace_bytes(byref(s), 1, B, 1, _lib.MODE_EXACT, byref(nbytes)))
    ws = workspace.get(nbytes.value)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_query_prologue(
                byref(s), byref(weights.struct), _ptr(q), _ptr(uid), B, _ptr(qsub), _ptr(gq), _ptr(ws),
                ws.numel(), _stream_ptr(dev),
            )
        )
    return qsub, gq


def merge_topk(part_scores: torch.Tensor, part_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(R, B, k) per-shard partial top-k -> global (B, k), on the GPU (mol_merge_topk)."""
    lib = _lib.load()
    _require_cuda(part_scores, "part_scores")
    dev = part_scores.device
    R, B, kk = part_scores.shape
    assert kk == k and part_ids.shape == part_scores.shape
    ps = part_scores.detach().to(torch.float32).contiguous()
    pi = part_ids.detach().to(torch.int64).contiguous()
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    nbytes = c_size_t()
    _lib.check(lib.mol_merge_topk_workspace_bytes(R, B, k, byref(nbytes)))
    ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_merge_topk(_ptr(ps), _ptr(pi), R, B, k, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(), _stream_ptr(dev))
        )
    return out_s, out_i


def pack_topk(scores: torch.Tensor, ids: torch.Tensor, k: int) -> torch.Tensor:
    """(B, k_valid <= k) partial list -> (B, k) packed entries (uint8 (B, k, 16)) for the single all-gather."""
    lib = _lib.load()
    _require_cuda(scores, "scores")
    B, kv = scores.shape
    s = scores.detach().to(torch.float32).contiguous()
    i = ids.detach().to(torch.int64).contiguous()
    out = torch.empty((B, k, _lib.MOL_PACKED_ENTRY_BYTES), dtype=torch.uint8, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(lib.mol_pack_topk(_ptr(s), _ptr(i), B, kv, k, _ptr(out), _stream_ptr(scores.device)))
    return out


def merge_topk_packed(gathered: torch.Tensor, R: int, B: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """gathered (R, B, k, 16) uint8 packed entries -> global (B, k) scores / ids (mol_merge_topk_packed)."""
    lib = _lib.load()
    _require_cuda(gathered, "gathered")
    dev = gathered.device
    assert gathered.is_contiguous() and gathered.numel() == R * B * k * _lib.MOL_PACKED_ENTRY_BYTES
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    nbytes = c_size_t()
    _lib.check(lib.mol_merge_topk_packed_workspace_bytes(R, B, k, byref(nbytes)))
    ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mol_merge_topk_packed(_ptr(gathered), R, B, k, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(),
                                             _stream_ptr(dev)))
    return out_s, out_i


def tensor_path_supported(shape: MolShape) -> bool:
    ok = c_int32(0)
    _lib.check(_lib.load().mol_shape_check(byref(shape), byref(ok)))
    return bool(ok.value)


def topk(scores: torch.Tensor, k: int, id_map: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-wise top-k of a (B, n) fp32 CUDA matrix through mol_topk (sorted, largest first)."""
    lib = _lib.load()
    _require_cuda(scores, "scores")
    assert scores.dim() == 2 and scores.dtype == torch.float32 and scores.stride(1) == 1
    dev = scores.device
    B, n = scores.shape
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    nbytes = c_size_t()
    _lib.check(lib.mol_topk_workspace_bytes(n, B, k, byref(nbytes)))
    ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_topk(
                _ptr(scores), n, scores.stride(0), B, k, _ptr(id_map), _ptr(out_s), _ptr(out_i), _ptr(ws),
                ws.numel(), _stream_ptr(dev),
            )
        )
    return out_s, out_i


def search_avg(
    weights: PackedWeights,
    index: IndexHandle,
    avg_items: torch.Tensor,
    workspace: Workspace,
    queries: torch.Tensor,
    user_ids: Optional[torch.Tensor],
    k: int,
    avg_top_k: int,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """mol_search_avg: MoLAvgTopK.forward on the GPU.  Returns (scores (B,k) fp32, ids (B,k) int64)."""
    lib = _lib.load()
    _require_cuda(queries, "query_embeddings")
    dev = index.device
    q = queries.detach().to(device=dev, dtype=torch.float32).contiguous()
    B = int(q.size(0))
    uid = None
    if weights.shape.num_uid_tables > 0:
        if user_ids is None:
            raise KeyError("user_ids")
        uid = user_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    nbytes = c_size_t()
    _lib.check(lib.mol_search_avg_workspace_bytes(byref(weights.shape), index.N, B, k, avg_top_k, byref(nbytes)))
    ws = workspace.get(nbytes.value)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_search_avg(
                byref(weights.shape), byref(weights.struct), byref(index.struct), _ptr(avg_items), _ptr(q), _ptr(uid),
                B, k, avg_top_k, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(), _stream_ptr(dev),
            )
        )
    return out_s, out_i


def search_groups(
    weights: PackedWeights,
    index: IndexHandle,
    avg_items: Optional[torch.Tensor],
    workspace: Workspace,
    queries: torch.Tensor,
    user_ids: Optional[torch.Tensor],
    k_per_group: int,
    avg_top_k: int,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """mol_search_groups: MoLNaiveTopK.forward (avg_top_k == 0) / MoLCombTopK.forward on the GPU.
    Returns (scores (B, C) fp32, ids (B, C) int64) with C = P_Q * P_X * k_per_group + avg_top_k."""
    lib = _lib.load()
    _require_cuda(queries, "query_embeddings")
    dev = index.device
    q = queries.detach().to(device=dev, dtype=torch.float32).contiguous()
    B = int(q.size(0))
    uid = None
    if weights.shape.num_uid_tables > 0:
        if user_ids is None:
            raise KeyError("user_ids")
        uid = user_ids.detach().to(device=dev, dtype=torch.int64).contiguous()
    C = weights.shape.query_dot_product_groups * weights.shape.item_dot_product_groups * k_per_group + avg_top_k
    out_s = torch.empty((B, C), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, C), dtype=torch.int64, device=dev)
    nbytes = c_size_t()
    _lib.check(lib.mol_search_groups_workspace_bytes(byref(weights.shape), index.N, B, k_per_group, avg_top_k, byref(nbytes)))
    ws = workspace.get(nbytes.value)
    with torch.cuda.device(dev):
        _lib.check(
            lib.mol_search_groups(
                byref(weights.shape), byref(weights.struct), byref(index.struct), _ptr(avg_items), _ptr(q), _ptr(uid),
                B, k_per_group, avg_top_k, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(), _stream_ptr(dev),
            )
        )
    return out_s, out_i


def avg_item_embeddings(weights: PackedWeights, index: IndexHandle) -> torch.Tensor:
    """(N, d) fp32 mean over the P_X item sub-embeddings (MoLAvgTopK's prefilter operand)."""
    lib = _lib.load()
    out = torch.empty((index.N, weights.shape.dot_product_dimension), dtype=torch.float32, device=index.device)
    with torch.cuda.device(index.device):
        _lib.check(lib.mol_index_avg_embeddings(byref(weights.shape), byref(index.struct), _ptr(out), _stream_ptr(index.device)))
    return out
