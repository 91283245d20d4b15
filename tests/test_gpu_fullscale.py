"""GPU parity at the sizes the benchmark runs at ("parity holes"): the fused candidate-filter strategy
(>= 65 536 items) pinned DIRECTLY to reference outputs and to the CPU oracle over the full 1M-item north-star corpus,
the acceptance test of the coarse candidate set validated on 10^4 queries and on an adversarially dense score
distribution, an ordered corpus (strided threshold sample), and the counters / packed-exchange / prepared-weights entry
points.  Everything goes through the C ABI."""
import time

import pytest
import torch

from oracle import mol_oracle as O
from rails_b200 import _lib, engine
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs

pytestmark = pytest.mark.gpu

SCORE_TOL, TIE_TOL = 1e-3, 1e-4
DEV = "cuda:0"


def _stats(mol):
    return engine.search_stats(mol.workspace(torch.device(DEV)))


@pytest.fixture
def force_filter(monkeypatch):
    """The fused-filter strategy is used from 2^24 (query, item) pairs; these tests reach it with a handful of queries."""
    monkeypatch.setenv("MOL_B200_FILTER_MIN_PAIRS", "0")


@pytest.mark.parametrize("mode", [_lib.MODE_AUTO, _lib.MODE_EXACT])
def test_large_reference_fixture_filter_strategy(mode, force_filter):
    """150 000 items: above kFilterMinItems, so MODE_AUTO takes the strided-sample threshold + fused filter path.  The
    expected values are outputs of the UNMODIFIED reference (oracle/gen_golden_large.py)."""
    from tests.golden_util import load_large

    g = load_large()
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    items, ids = g["items"].to(DEV), g["item_ids"].to(DEV)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode)
    s, i = top(g["queries"].to(DEV), k=g["k"])
    st = _stats(mol)
    if mode == _lib.MODE_AUTO:
        assert st["filter_strategy"] == 1 and st["tensor_path"] == 1, st
        assert st["fallback_queries"] == 0, st
    assert (s.cpu() - g["ref_top_scores"]).abs().max().item() <= SCORE_TOL
    rs = g["ref_top_scores"].double()
    gap_prev = torch.cat([torch.full_like(rs[:, :1], float("inf")), (rs[:, :-1] - rs[:, 1:]).abs()], dim=1)
    gap_next = torch.cat([(rs[:, :-1] - rs[:, 1:]).abs(), torch.zeros_like(rs[:, :1])], dim=1)
    clear = (gap_prev > 2 * TIE_TOL) & (gap_next > 2 * TIE_TOL)
    assert float(clear.double().mean()) > 0.9
    assert bool((i.cpu()[clear] == g["ref_top_ids"][clear]).all())  # ids bit-exact on every unambiguous rank
    # and the similarity module on the same corpus against the reference's score matrix (strided columns)
    if mode == _lib.MODE_EXACT:
        sc, _ = mol(g["queries"].to(DEV), items.unsqueeze(0))
        assert (sc[:, :: g["col_stride"]].cpu() - g["ref_scores_strided"]).abs().max().item() <= 2e-4


def test_north_star_16_queries_full_corpus_vs_oracle(force_filter):
    """The benchmarked configuration (8x8x32, 1M items, top-100): 16 queries against the CPU oracle over the FULL corpus."""
    cfg = CFG_8x8x32
    N, B, k = 1_000_000, 16, 100
    mol, _ = build_module(cfg, None, DEV, seed=0)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 0, DEV)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
    s, got = top(q, k=k)
    st = _stats(mol)
    assert st["filter_strategy"] == 1 and st["fallback_queries"] == 0, st
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    torch.set_num_threads(max(1, torch.get_num_threads()))
    _, _, all_scores = O.brute_force_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), k, chunk=2)
    r = O.compare_top_k(s, got, all_scores, ids.cpu(), k, SCORE_TOL, TIE_TOL)
    assert r["ok"] == 1.0, r
    assert r["strict_row_match"] >= 0.8, r  # (a row differs only where fp32 itself reorders a near-tie)


def test_north_star_all_512_bench_queries_auto_equals_exact():
    """Every query of the benchmark batch: the tensor-core path returns exactly what the fp32 exact mode returns."""
    import torch.nn.functional as F

    cfg = CFG_8x8x32
    N, B, k = 1_000_000, 512, 100
    mol, _ = build_module(cfg, None, DEV, seed=0)
    g = torch.Generator(device=DEV).manual_seed(1)
    items = 0.02 * torch.randn(N, cfg.item_embedding_dim, device=DEV, generator=g)  # bench.py's corpus
    ids = torch.arange(1, N + 1, dtype=torch.int64, device=DEV)
    gq = torch.Generator().manual_seed(100)
    q = F.layer_norm(torch.randn(B, cfg.query_embedding_dim, generator=gq), (cfg.query_embedding_dim,)).to(DEV)
    s, i = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_AUTO)(q, k=k)
    st = _stats(mol)
    s_ex, i_ex = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)(q, k=k)
    assert st["fallback_queries"] == 0 and st["filter_overflows"] == 0, st
    assert (s - s_ex).abs().max().item() < 1e-4
    bad = i != i_ex
    assert float(bad.float().mean()) < 1e-3
    if bool(bad.any()):  # only exact ties / fp32-reordered near-ties may differ
        assert (s[bad] - s_ex[bad]).abs().max().item() < 1e-5


def test_acceptance_check_10k_queries_no_miss():
    """The candidate-set acceptance test (safety_flags_kernel) is an empirical margin, not a proof.  Here
    it is validated on 10 240 queries x 200k items: whenever it accepts the coarse candidate set, the result must equal the
    exact mode's - zero misses allowed."""
    cfg = CFG_8x8x32
    N, B, k = 200_000, 10_240, 100
    mol, _ = build_module(cfg, None, DEV, seed=41)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 41, DEV)
    s, i = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_AUTO)(q, k=k)
    st = _stats(mol)
    s_ex, i_ex = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)(q, k=k)
    assert st["fallback_queries"] <= B // 100, st  # the fast path served (nearly) every query
    assert (s - s_ex).abs().max().item() < 1e-4
    bad = i != i_ex
    if bool(bad.any()):
        assert (s[bad] - s_ex[bad]).abs().max().item() < 1e-5, int(bad.sum())
    # membership: the SETS agree for every query wherever scores are not tied at the cut
    miss = 0
    for b in torch.nonzero(bad.any(dim=1)).flatten().tolist():
        if set(i[b].tolist()) != set(i_ex[b].tolist()) and abs(float(s_ex[b, -1]) - float(s[b, -1])) > 1e-6:
            miss += 1
    assert miss == 0


@pytest.mark.parametrize("P,reps,expect", [(256, 400, "second_chance"), (64, 5000, "exact_fallback")])
def test_dense_scores_tiny_gap_falls_back_and_stays_exact(P, reps, expect, force_filter):
    """Adversarial for the acceptance test: the corpus is a few hundred prototypes, each repeated hundreds / thousands of times
    with perturbations far below the coarse pass's resolution, so the gap between rank k and rank K' is ~1e-5 while the
    coarse error is ~1e-2.  The first test must refuse the coarse top-K'.  With 400 copies per prototype the second chance
    (every survivor of the fused filter rescored: the copies of the best two or three prototypes) proves the answer; with
    5000 copies the filter threshold falls INSIDE the best prototype's cluster (the ~1100 survivors are the copies with the
    largest coarse scores, the other ~3900 copies tie with them exactly and lie just below the threshold), the second test
    must refuse as well and the exact fallback must return the exact answer."""
    cfg = CFG_8x8x32
    B, k = 6, 100
    mol, _ = build_module(cfg, None, DEV, seed=43)
    proto, _, q, _ = synthetic_inputs(cfg, P, B, 43, DEV)
    g = torch.Generator(device=DEV).manual_seed(7)
    items = proto.repeat_interleave(reps, dim=0)
    items = items + 2e-6 * torch.randn(items.shape, device=DEV, generator=g)
    N = items.size(0)
    ids = torch.randperm(N, generator=torch.Generator().manual_seed(3)).to(DEV) + 1
    s, i = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_AUTO)(q, k=k)
    st = _stats(mol)
    s_ex, i_ex = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)(q, k=k)
    assert st["tensor_path"] == 1 and st["fallback_queries"] + st["second_chance_queries"] == B, st
    if expect == "second_chance":
        assert st["second_chance_queries"] == B, st
    else:
        assert st["fallback_queries"] == B, st
    assert torch.equal(s, s_ex)
    # near-duplicate items tie EXACTLY in fp32, and which members of a tie at rank k a select keeps is unspecified (as with
    # torch.topk): ids are checked against the full exact score matrix instead of against the other run's choice
    full = mol(q, items.unsqueeze(0))[0]
    pos = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    pos[ids] = torch.arange(N, device=DEV)
    for res_s, res_i in ((s, i), (s_ex, i_ex)):
        assert torch.equal(torch.gather(full, 1, pos[res_i]), res_s)
        assert torch.equal(res_s, full.topk(k, dim=1).values)
        assert all(res_i[b].unique().numel() == k for b in range(B))


def test_ordered_corpus_no_fallback_and_no_slowdown():
    """A corpus ORDERED by a popularity-like score (the mean exact score over all queries, descending).  With the threshold
    taken from the first items of the corpus this made every query overflow its candidate buffer and fall back to the exact
    kernel (a 27x cliff); the strided sample must give zero fallbacks and the same speed as the
    shuffled corpus."""
    cfg = CFG_8x8x32
    N, B, k = 400_000, 64, 100
    mol, _ = build_module(cfg, None, DEV, seed=47)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 47, DEV)
    pop = mol(q, items.unsqueeze(0))[0].mean(dim=0)
    order = torch.argsort(pop, descending=True)
    w, wsp = mol.packed_weights(torch.device(DEV)), mol.workspace(torch.device(DEV))

    def run(it, idd):
        top = MoLBruteForceTopK(mol, it.unsqueeze(0), idd.unsqueeze(0))
        index = top._ensure_index()
        for _ in range(3):
            s, i = engine.search(w, index, wsp, q, None, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            s, i = engine.search(w, index, wsp, q, None, k)
        e1.record()
        torch.cuda.synchronize()
        return s, i, engine.search_stats(wsp), e0.elapsed_time(e1) / 5

    s0, i0, st0, ms0 = run(items, ids)
    s1, i1, st1, ms1 = run(items[order].contiguous(), ids[order].contiguous())
    assert st0["fallback_queries"] == 0 and st1["fallback_queries"] == 0, (st0, st1)
    assert st1["filter_overflows"] == 0, st1
    assert torch.equal(i0, i1) and torch.equal(s0, s1)  # same corpus, same answer (ties aside: none here)
    assert ms1 <= 1.1 * ms0 + 0.05, (ms0, ms1)


def test_prepared_weights_equal_per_call_preparation(force_filter):
    """mol_weights_prepare (once per weight version) vs weights->prepared == NULL (operands recomputed inside every call)."""
    cfg = CFG_8x8x32
    N, B, k = 70_000, 9, 50
    mol, _ = build_module(cfg, None, DEV, seed=3)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 3, DEV)
    dev = torch.device(DEV)
    w = mol.packed_weights(dev)
    assert w.struct.prepared, "PackedWeights did not prepare the weight-derived operands"
    index = mol.build_index(items, ids)
    a = engine.search(w, index, mol.workspace(dev), q, None, k)
    keep = w.struct.prepared
    w.struct.prepared = None
    try:
        b = engine.search(w, index, mol.workspace(dev), q, None, k)
    finally:
        w.struct.prepared = keep
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_pack_and_merge_packed_match_torch():
    g = torch.Generator(device=DEV).manual_seed(5)
    R, B, k = 8, 33, 100
    ps = torch.sort(torch.randn(R, B, k, device=DEV, generator=g), dim=2, descending=True).values
    pi = torch.randint(-(1 << 40), 1 << 40, (R, B, k), device=DEV, generator=g)  # negative ids are legal
    kv = [k, k, 37, 0, k, 5, k, k]  # short shards contribute fewer than k entries
    packed = torch.stack([engine.pack_topk(ps[r, :, : kv[r]].contiguous(), pi[r, :, : kv[r]].contiguous(), k) for r in range(R)])
    s, i = engine.merge_topk_packed(packed, R, B, k)
    flat_s = torch.cat([ps[r, :, : kv[r]] for r in range(R)], dim=1)
    flat_i = torch.cat([pi[r, :, : kv[r]] for r in range(R)], dim=1)
    rs, rj = torch.topk(flat_s, k, dim=1)
    assert torch.equal(s, rs) and torch.equal(i, torch.gather(flat_i, 1, rj))


def test_bf16_model_and_inputs():
    """eval_from_checkpoint.py:318-322 casts the model to bf16.  The CUDA path takes bf16 parameters / items / queries,
    computes in fp32 on their (exactly representable) values and returns the queries' dtype: the ranking must match the
    fp32 oracle evaluated on the same bf16-rounded values, scores to bf16 resolution."""
    cfg = CFG_8x8x32
    N, B, k = 30_000, 12, 50
    mol, _ = build_module(cfg, None, DEV, seed=12)
    mol = mol.to(torch.bfloat16)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 12, DEV)
    items, q = items.to(torch.bfloat16), q.to(torch.bfloat16)
    s, i = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))(q, k=k)
    assert s.dtype == torch.bfloat16 and i.dtype == torch.int64
    sd = {k_: v.detach().float().cpu() for k_, v in mol.state_dict().items()}
    rs, ri, all_scores = O.brute_force_top_k(cfg, sd, q.float().cpu(), items.float().cpu(), ids.cpu(), k)
    r = O.compare_top_k(s.float(), i, all_scores, ids.cpu(), k, score_tol=0.1, tie_tol=TIE_TOL)
    assert r["max_rank_gap"] <= TIE_TOL and r["duplicates"] == 0, r
    assert (s.float().cpu() - rs).abs().max().item() <= 2.0 ** -7 * float(rs.abs().max()) + 1e-3
    sc, _ = mol(q, items.unsqueeze(0))
    assert sc.dtype == torch.bfloat16 and sc.shape == (B, N)


def test_small_batches_take_the_matrix_strategy_and_agree_with_the_filter_strategy(monkeypatch):
    """Below 2^24 (query, item) pairs the coarse pass writes its score matrix (one launch + one select) instead of running
    the threshold pass; both strategies must return the same answer."""
    cfg = CFG_8x8x32
    N, B, k = 200_000, 4, 100
    mol, _ = build_module(cfg, None, DEV, seed=5)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 5, DEV)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
    a = top(q, k=k)
    st_a = _stats(mol)
    monkeypatch.setenv("MOL_B200_FILTER_MIN_PAIRS", "0")
    b = top(q, k=k)
    st_b = _stats(mol)
    assert st_a["tensor_path"] == 1 and st_a["filter_strategy"] == 0 and st_b["filter_strategy"] == 1, (st_a, st_b)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])


def test_cuda_graph_search_equals_eager():
    """MoLBruteForceTopK(cuda_graph=True): every (B, k) signature is captured once and replayed; results are the eager ones,
    replays with new queries see the new queries, and the real ML-1M configuration (uid embeddings) works through it."""
    from tests.golden_util import load_golden

    cfg = CFG_8x8x32
    N, k = 50_000, 50
    mol, _ = build_module(cfg, None, DEV, seed=6)
    items, ids, q, _ = synthetic_inputs(cfg, N, 24, 6, DEV)
    eager = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
    graphed = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), cuda_graph=True)
    for B in (1, 8, 8, 24, 1):
        for off in (0, 3):
            qq = q[off : off + B] if off + B <= 24 else q[:B]
            es, ei = eager(qq, k=k)
            gs, gi = graphed(qq, k=k)
            assert torch.equal(ei, gi) and torch.equal(es, gs), (B, off)
    g = load_golden("cfg1_ml1m_ckpt")
    mol1, _ = build_module(g["cfg"], g["sd"], DEV)
    top1 = MoLBruteForceTopK(mol1, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0), cuda_graph=True)
    u = g["user_ids"].to(DEV)
    for _ in range(2):
        s, i = top1(g["queries"].to(DEV), k=g["k"], user_ids=u)
        r = O.compare_top_k(s, i, g["ref_scores"], g["item_ids"], g["k"], SCORE_TOL, TIE_TOL)
        assert r["ok"] == 1.0, r


def test_small_searches_take_the_exact_kernel_in_auto_mode(monkeypatch):
    """Below 2^15 (query, item) pairs MODE_AUTO runs the fp32 kernel over every pair (BASELINE config 1: one query over
    3883 items) - same answer as the tensor path and as MODE_EXACT, `tensor_path` = 0 in the stats."""
    from tests.golden_util import load_golden

    g = load_golden("cfg1_ml1m_ckpt")
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    items, ids = g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0)
    q, uid = g["queries"][:1].to(DEV), g["user_ids"][:1].to(DEV)
    out = {}
    for name, pairs in (("switch", None), ("tensor", "0")):
        if pairs is None:
            monkeypatch.delenv("MOL_B200_TENSOR_MIN_PAIRS", raising=False)
        else:
            monkeypatch.setenv("MOL_B200_TENSOR_MIN_PAIRS", pairs)
        s, i = MoLBruteForceTopK(mol, items, ids, mode=_lib.MODE_AUTO)(q, k=g["k"], user_ids=uid)
        out[name] = (s, i, _stats(mol))
    assert out["switch"][2]["tensor_path"] == 0 and out["tensor"][2]["tensor_path"] == 1
    assert torch.equal(out["switch"][1], out["tensor"][1]) and torch.equal(out["switch"][0], out["tensor"][0])
    r = O.compare_top_k(out["switch"][0][:1], out["switch"][1][:1], g["ref_scores"][:1], g["item_ids"], g["k"], 1e-3, 1e-4)
    assert r["ok"] == 1.0, r


def test_second_chance_rescoring_all_survivors_equals_exact(force_filter, monkeypatch):
    """Filter strategy: a query the first acceptance test refuses gets EVERY survivor of the fused filter rescored in fp32
    before the full exact pass is considered.  With the first test forced to refuse everything
    (MOL_B200_FORCE_SECOND_CHANCE) all queries must be served by the second chance (no exact fallback) and equal the
    exact mode."""
    cfg = CFG_8x8x32
    N, B, k = 300_000, 24, 100
    mol, _ = build_module(cfg, None, DEV, seed=61)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 61, DEV)
    monkeypatch.setenv("MOL_B200_FORCE_SECOND_CHANCE", "1")
    s, i = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_AUTO)(q, k=k)
    st = _stats(mol)
    monkeypatch.delenv("MOL_B200_FORCE_SECOND_CHANCE")
    s_ex, i_ex = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)(q, k=k)
    assert st["filter_strategy"] == 1 and st["second_chance_queries"] == B and st["fallback_queries"] == 0, st
    assert torch.equal(i, i_ex) and torch.equal(s, s_ex)
    # and without the hook nothing is refused on this workload
    s2, i2 = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_AUTO)(q, k=k)
    st2 = _stats(mol)
    assert st2["second_chance_queries"] == 0 and st2["fallback_queries"] == 0, st2
    assert torch.equal(i2, i_ex) and torch.equal(s2, s_ex)
