"""world_size-2 `gloo` tests (CPU) of the multi-GPU top-k host logic (corpus-sharded and query-split modes): ranges,
the single packed all-gather, short shards / ragged query slices and the merge call.  The rank-local search, the pack
and the merge are injected (oracle / numpy based here, same 16-byte entry layout as include/mol_b200.h); on a GPU box
they are the CUDA engine (tests/test_gpu_parity.py covers those)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import numpy as np

from oracle import mol_oracle as O

from rails_b200.indexing.sharded_top_k import (
    PACKED_ENTRY_BYTES, ReplicatedMoLBruteForceTopK, ShardedMoLBruteForceTopK, shard_range,
)
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


ENTRY = np.dtype([("id", "<i8"), ("score", "<f4"), ("valid", "<i4")])  # MOL_PACKED_ENTRY_BYTES layout
assert ENTRY.itemsize == PACKED_ENTRY_BYTES


def _cpu_pack(scores, ids, k):
    B, kv = scores.shape
    e = np.zeros((B, k), dtype=ENTRY)
    e["id"][:, :kv] = ids.numpy()
    e["score"][:, :kv] = scores.numpy()
    e["valid"][:, :kv] = 1
    e["id"][:, kv:] = -1
    e["score"][:, kv:] = -np.inf
    return torch.from_numpy(e.view(np.uint8).reshape(B, k, PACKED_ENTRY_BYTES).copy())


def _cpu_merge_packed(gathered, R, B, k):
    e = gathered.contiguous().numpy().view(ENTRY).reshape(R, B, k)
    sc = torch.from_numpy(np.where(e["valid"] != 0, e["score"], -np.inf).astype(np.float32))
    idv = torch.from_numpy(e["id"].copy())
    flat_s = sc.permute(1, 0, 2).reshape(B, R * k)
    flat_i = idv.permute(1, 0, 2).reshape(B, R * k)
    s, j = torch.topk(flat_s, k, dim=1)
    return s, torch.gather(flat_i, 1, j)


def _worker(rank, world, port, N, B, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = CFG_8x8x32
        mol, _ = build_module(cfg, None, "cpu", seed=5)
        sd = {k_: v.detach() for k_, v in mol.state_dict().items()}
        items, ids, q, _ = synthetic_inputs(cfg, N, B, 5)
        lo, hi = shard_range(N, rank, world)

        def local(qe, kk, sorted=True, **kw):  # the oracle stands in for the CUDA shard search
            s, i, _ = O.brute_force_top_k(cfg, sd, qe, items[lo:hi], ids[lo:hi], kk)
            return s, i

        top = ShardedMoLBruteForceTopK(local, hi - lo, pack=_cpu_pack, merge_packed=_cpu_merge_packed)
        s, i = top(q, k)
        ref_s, ref_i, _ = O.brute_force_top_k(cfg, sd, q, items, ids, k)
        ok = bool(torch.equal(i, ref_i)) and float((s - ref_s).abs().max()) < 1e-5
        # k larger than the whole corpus: the reference's RuntimeError on EVERY rank, before any collective
        try:
            top(q, N + 1)
            ok = False
        except RuntimeError as e:
            ok = ok and "out of range" in str(e)

        # query-split mode: every rank holds the whole corpus and searches its slice of the batch
        def whole(qe, kk, sorted=True, **kw):
            s_, i_, _ = O.brute_force_top_k(cfg, sd, qe, items, ids, kk)
            return s_, i_

        rep = ReplicatedMoLBruteForceTopK(whole, pack=_cpu_pack, merge_packed=_cpu_merge_packed)
        k2 = min(k, N)
        s2, i2 = rep(q, k2)
        ref_s2, ref_i2, _ = O.brute_force_top_k(cfg, sd, q, items, ids, k2)
        ok = ok and s2.shape == (B, k2) and bool(torch.equal(i2, ref_i2)) and float((s2 - ref_s2).abs().max()) < 1e-5
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,B,k", [(600, 3, 50), (9, 2, 7), (300, 5, 20)])  # 2nd: shards (4 / 5 items) shorter than k; 3rd: ragged query slices
def test_sharded_topk_equals_unsharded_world2(N, B, k):
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, N, B, k, out), nprocs=2, join=True)
    assert out[0] is True and out[1] is True


def test_shard_ranges_cover_corpus():
    for N in (0, 1, 7, 1000, 1_000_003):
        for R in (1, 2, 4, 8):
            spans = [shard_range(N, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[r][1] == spans[r + 1][0] for r in range(R - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_cpu_pack_merge_roundtrip_and_negative_ids():
    """The packed entry carries an explicit validity flag: negative ids are legal payloads."""
    g = torch.Generator().manual_seed(1)
    s = torch.sort(torch.randn(5, 7, generator=g), dim=1, descending=True).values
    i = torch.randint(-(1 << 40), 1 << 40, (5, 7), generator=g)
    s2, i2 = _cpu_merge_packed(_cpu_pack(s, i, 7).unsqueeze(0), 1, 5, 7)
    assert torch.equal(s, s2) and torch.equal(i, i2)
    # two short "shards" (4 and 3 valid entries, the rest padding) merge back to the full sorted list
    parts = torch.stack([_cpu_pack(s[:, 0::2], i[:, 0::2], 7), _cpu_pack(s[:, 1::2], i[:, 1::2], 7)])
    s3, i3 = _cpu_merge_packed(parts, 2, 5, 7)
    assert torch.equal(s, s3) and torch.equal(i, i3)


def test_single_process_passthrough_and_range_error():
    calls = []

    def local(qe, kk, sorted=True, **kw):
        calls.append(kk)
        return torch.zeros(qe.size(0), kk), torch.arange(kk).repeat(qe.size(0), 1)

    top = ShardedMoLBruteForceTopK(local, shard_items=5, pack=_cpu_pack, merge_packed=_cpu_merge_packed)
    s, i = top(torch.zeros(2, 4), 3)
    assert s.shape == (2, 3) and calls == [3]
    with pytest.raises(RuntimeError, match="out of range"):
        top(torch.zeros(2, 4), 6)
