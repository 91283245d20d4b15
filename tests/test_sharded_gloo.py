"""world_size-2 `gloo` tests (CPU) of the corpus-sharded top-k host logic: shard ranges, the single packed
all-gather, padding of short shards and the merge call.  The rank-local search and the merge are injected
(oracle-based here); on a GPU box they are the CUDA engine (tests/test_gpu_parity.py covers those)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mol_oracle as O
from rails_b200.indexing.sharded_top_k import ShardedMoLBruteForceTopK, pack_partials, shard_range, unpack_partials
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_merge(ps, pi, k):
    R, B, kk = ps.shape
    flat_s = ps.permute(1, 0, 2).reshape(B, R * kk)
    flat_i = pi.permute(1, 0, 2).reshape(B, R * kk)
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

        top = ShardedMoLBruteForceTopK(local, hi - lo, merge=_cpu_merge)
        s, i = top(q, k)
        ref_s, ref_i, _ = O.brute_force_top_k(cfg, sd, q, items, ids, k)
        ok = bool(torch.equal(i, ref_i)) and float((s - ref_s).abs().max()) < 1e-5
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,B,k", [(600, 3, 50), (9, 2, 7)])  # second case: shards (4 / 5 items) shorter than k
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


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(1)
    s = torch.randn(5, 11, generator=g)
    s[0, 0] = float("-inf")
    i = torch.randint(-1, 1 << 40, (5, 11), generator=g)
    s2, i2 = unpack_partials(pack_partials(s, i))
    assert torch.equal(s, s2) and torch.equal(i, i2)
    stacked = torch.stack([pack_partials(s, i), pack_partials(s + 1, i + 1)])
    s3, i3 = unpack_partials(stacked)
    assert torch.equal(s3[1], s + 1) and torch.equal(i3[1], i + 1)


def test_single_process_passthrough_and_range_error():
    calls = []

    def local(qe, kk, sorted=True, **kw):
        calls.append(kk)
        return torch.zeros(qe.size(0), kk), torch.arange(kk).repeat(qe.size(0), 1)

    top = ShardedMoLBruteForceTopK(local, shard_items=5, merge=_cpu_merge)
    s, i = top(torch.zeros(2, 4), 3)
    assert s.shape == (2, 3) and calls == [3]
    with pytest.raises(RuntimeError, match="out of range"):
        top(torch.zeros(2, 4), 6)
