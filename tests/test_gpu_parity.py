"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped Python modules) against
(a) outputs of the UNMODIFIED reference stored in tests/golden/*.npz and (b) the CPU oracle on seeded
inputs.  Bar (BASELINE.json north_star): top-k indices identical (tie-aware: only items whose oracle
scores differ by < 1e-4 may swap), scores within 1e-3 fp32."""
import numpy as np
import pytest
import torch

from oracle import mol_oracle as O
from rails_b200 import _lib, engine
from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
from tests.golden_util import golden_names, load_golden
from tests.helpers import CFG_8x4x64, CFG_8x4x128, CFG_8x8x32, CFG_16x16x64, build_module, synthetic_inputs

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-3  # north_star: "scores within 1e-3 fp32"
TIE_TOL = 1e-4    # SURVEY.md §7 hard part 2
DEV = "cuda:0"
MODES = [_lib.MODE_EXACT, _lib.MODE_AUTO]


def _kwargs(g):
    if g["user_ids"] is None:
        return {}
    u = g["user_ids"].to(DEV)
    return dict(user_ids=u, timestamps=torch.zeros_like(u), ratings=torch.zeros_like(u))


@pytest.mark.parametrize("name", golden_names())
def test_similarity_scores_match_reference(name):
    g = load_golden(name)
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    scores, aux = mol(g["queries"].to(DEV), g["items"].to(DEV).unsqueeze(0), **_kwargs(g))
    assert aux == {} and scores.shape == g["ref_scores"].shape and scores.dtype == torch.float32
    err = (scores.cpu() - g["ref_scores"]).abs().max().item()
    assert err <= SCORE_TOL, err
    assert err <= 2e-4, f"fp32 CUDA path should be within rounding of the reference, got {err}"


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", golden_names())
def test_top_k_matches_reference(name, mode):
    g = load_golden(name)
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    top = MoLBruteForceTopK(mol, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0), mode=mode)
    s, ids = top(g["queries"].to(DEV), k=g["k"], sorted=True, **_kwargs(g))
    assert s.shape == (g["queries"].size(0), g["k"]) and ids.dtype == torch.int64
    r = O.compare_top_k(s, ids, g["ref_scores"], g["item_ids"], g["k"], SCORE_TOL, TIE_TOL)
    assert r["ok"] == 1.0, r
    # scores are sorted descending
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    # and against the reference's own top-k output: same score profile
    assert (s.cpu() - g["ref_top_scores"]).abs().max().item() <= SCORE_TOL
    # ids bit-exact wherever the reference's own ranking is unambiguous (its neighbours' scores differ by more than
    # the fp32 near-tie band); inside a near-tie torch.topk's order is not defined
    rs = g["ref_top_scores"].double()
    gap_prev = torch.cat([torch.full_like(rs[:, :1], float("inf")), (rs[:, :-1] - rs[:, 1:]).abs()], dim=1)
    gap_next = torch.cat([(rs[:, :-1] - rs[:, 1:]).abs(), torch.full_like(rs[:, :1], 0.0)], dim=1)  # the cut at k is never "clear"
    clear = (gap_prev > 2 * TIE_TOL) & (gap_next > 2 * TIE_TOL)
    assert bool((ids.cpu()[clear] == g["ref_top_ids"][clear]).all())
    if g["k"] >= 5:
        assert float(clear.double().mean()) > 0.5, "fixture has too few unambiguous ranks to pin the ids"


@pytest.mark.parametrize("name", ["cfg1_ml1m_ckpt", "cfg2_8x4x128", "cfg3_8x8x32"])
def test_query_prologue_matches_oracle(name):
    g = load_golden(name)
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    qs, _ = mol.get_query_component_embeddings(g["queries"].to(DEV), **_kwargs(g))
    ref = O.query_sub_embeddings(g["cfg"], g["sd"], g["queries"], g["user_ids"])
    assert (qs.cpu() - ref).abs().max().item() < 1e-5
    xs, _ = mol.get_item_component_embeddings(g["items"].to(DEV))
    refx = O.item_sub_embeddings(g["cfg"], g["sd"], g["items"])
    assert (xs.cpu() - refx).abs().max().item() < 1e-5


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize(
    "cfg,N,B,k,seed",
    [
        (CFG_8x8x32, 20000, 24, 100, 3),
        (CFG_8x8x32, 1000, 5, 1000, 4),      # k == N
        (CFG_8x8x32, 129, 1, 7, 5),          # one item past a 128 tile
        (CFG_8x4x128, 3000, 9, 100, 6),
        (CFG_16x16x64, 2500, 6, 50, 7),
    ],
)
def test_top_k_matches_oracle_seeded(cfg, N, B, k, seed, mode):
    mol, _ = build_module(cfg, None, DEV, seed=seed)
    items, ids, q, uid = synthetic_inputs(cfg, N, B, seed, DEV)
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    kw = {} if uid is None else dict(user_ids=uid)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode)
    s, got = top(q, k=k, **kw)
    _, _, all_scores = O.brute_force_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), k, None if uid is None else uid.cpu())
    r = O.compare_top_k(s, got, all_scores, ids.cpu(), k, SCORE_TOL, TIE_TOL)
    assert r["ok"] == 1.0, r


@pytest.mark.parametrize("mode", MODES)
def test_duplicate_items_exact_score_collisions(mode):
    """Collisions: every embedding appears three times under different ids, so each score is an exact 3-way tie
    (the cut at k falls inside a tie group).  torch.topk leaves the order inside a tie unspecified; the tie-aware
    comparator accepts any member, but ids must not repeat and every returned score must be that item's score."""
    cfg = CFG_8x8x32
    base, B, k = 4000, 7, 100
    mol, _ = build_module(cfg, None, DEV, seed=9)
    items, _, q, _ = synthetic_inputs(cfg, base, B, 9, DEV)
    items = items.repeat(3, 1)
    ids = torch.randperm(3 * base, generator=torch.Generator().manual_seed(9)).to(DEV) + 1
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    s, got = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode)(q, k=k)
    _, _, all_scores = O.brute_force_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), k)
    r = O.compare_top_k(s, got, all_scores, ids.cpu(), k, SCORE_TOL, TIE_TOL)
    assert r["ok"] == 1.0, r
    # the three copies of one embedding get bit-identical scores from the rescoring kernel
    srt = s.cpu()
    assert int((srt[:, :-1] == srt[:, 1:]).sum()) >= B * (k // 3) * 2 - B * 2


def test_k_out_of_range_raises_runtime_error_like_torch_topk():
    g = load_golden("edge_tiny")
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    top = MoLBruteForceTopK(mol, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0))
    with pytest.raises(RuntimeError, match="out of range"):
        top(g["queries"].to(DEV), k=g["items"].size(0) + 1)


def test_c_abi_error_paths_return_status_not_crash():
    """The C entry points report bad calls through their int status + mol_last_error(): a workspace
    that is too small, NULL buffers, a candidate count beyond MOL_MAX_K - and the library keeps working afterwards."""
    from ctypes import byref

    cfg = CFG_8x8x32
    mol, _ = build_module(cfg, None, DEV, seed=1)
    items, ids, q, _ = synthetic_inputs(cfg, 5000, 4, 1, DEV)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
    index = top._ensure_index()
    w = mol.packed_weights(torch.device(DEV))
    lib = _lib.load()
    B, k = 4, 10
    out_s = torch.empty((B, k), dtype=torch.float32, device=DEV)
    out_i = torch.empty((B, k), dtype=torch.int64, device=DEV)
    small = torch.empty(1024, dtype=torch.uint8, device=DEV)
    stream = engine._stream_ptr(torch.device(DEV))

    def call(qp, sp, ip, wsp, nbytes):
        return lib.mol_search(byref(w.shape), byref(w.struct), byref(index.struct), qp, None, B, k, 1, _lib.MODE_AUTO,
                              sp, ip, wsp, nbytes, stream)

    P = engine._ptr
    assert call(P(q), P(out_s), P(out_i), P(small), small.numel()) == 3  # MOL_ERR_WORKSPACE
    assert b"workspace" in lib.mol_last_error()
    with pytest.raises(RuntimeError, match="workspace"):
        _lib.check(3)
    assert call(None, P(out_s), P(out_i), P(small), small.numel()) == 1  # MOL_ERR_INVALID (NULL queries)
    with pytest.raises(ValueError):
        _lib.check(1)
    # P_Q * P_X * k_per_group beyond MOL_MAX_K: refused up front
    from rails_b200.indexing.mol_top_k import MoLNaiveTopK

    with pytest.raises(ValueError, match="MOL_MAX_K"):
        MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), k_per_group=200)(q, k=10)
    # k_per_group larger than the corpus: torch.topk's RuntimeError
    with pytest.raises(RuntimeError, match="out of range"):
        MoLNaiveTopK(mol, items[:50].unsqueeze(0), ids[:50].unsqueeze(0), k_per_group=64)(q, k=10)
    # and a normal call still works
    s, i = top(q, k=k)
    assert s.shape == (B, k) and bool(torch.isfinite(s).all())


def test_empty_batch():
    g = load_golden("edge_tiny")
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    top = MoLBruteForceTopK(mol, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0))
    s, ids = top(torch.empty(0, 64, device=DEV), k=2)
    assert s.shape == (0, 2) and ids.shape == (0, 2)


def test_user_ids_required_when_configured():
    g = load_golden("cfg1_ml1m_ckpt")
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    top = MoLBruteForceTopK(mol, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0))
    with pytest.raises(KeyError):
        top(g["queries"].to(DEV), k=3)


def test_index_follows_weight_updates():
    """The reference recomputes the item side on every call; the cache must notice new weights."""
    g = load_golden("edge_ragged_kmax")
    mol, _ = build_module(g["cfg"], None, DEV, seed=123)
    top = MoLBruteForceTopK(mol, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0))
    s0, _ = top(g["queries"].to(DEV), k=5)
    mol.load_state_dict(g["sd"], strict=True)
    s1, ids1 = top(g["queries"].to(DEV), k=5)
    r = O.compare_top_k(s1, ids1, g["ref_scores"], g["item_ids"], 5, SCORE_TOL, TIE_TOL)
    assert r["ok"] == 1.0, r
    assert not torch.allclose(s0, s1)


# ------------------------------------------------------------------------------- selection kernels
@pytest.mark.parametrize(
    "B,n,k",
    [(3, 1000, 10), (1, 300000, 100), (2, 1 << 20, 200), (64, 5000, 4096), (5, 77, 77), (4, 100000, 1), (2, 50000, 8192)],
)
def test_topk_kernel_matches_torch(B, n, k):
    g = torch.Generator(device=DEV).manual_seed(B * 131 + k)
    x = torch.randn(B, n, device=DEV, generator=g)
    s, i = engine.topk(x, k)
    rs, ri = torch.topk(x, k, dim=1, largest=True, sorted=True)
    assert torch.equal(s, rs)
    assert torch.equal(i, ri)  # randn has no ties at these sizes w.h.p.; equality of scores is asserted first


def test_topk_kernel_ties_negatives_and_inf():
    x = torch.tensor([[1.0, -2.0, 1.0, float("-inf"), 0.0, -0.0, 3.5, 1.0]], device=DEV).repeat(2, 1)
    s, i = engine.topk(x, 5)
    assert s[0].tolist() == [3.5, 1.0, 1.0, 1.0, 0.0]
    assert i[0].tolist()[:4] == [6, 0, 2, 7]  # ties: lower position first
    # a long row of duplicates
    y = torch.zeros(1, 70000, device=DEV)
    y[0, 12345] = 1.0
    s, i = engine.topk(y, 4)
    assert s[0].tolist() == [1.0, 0.0, 0.0, 0.0] and i[0, 0].item() == 12345
    assert len(set(i[0].tolist())) == 4


def test_merge_topk_matches_torch():
    g = torch.Generator(device=DEV).manual_seed(5)
    R, B, k = 8, 33, 100
    ps = torch.randn(R, B, k, device=DEV, generator=g)
    pi = torch.randint(0, 1 << 40, (R, B, k), device=DEV, generator=g)
    s, i = engine.merge_topk(ps, pi, k)
    flat_s = ps.permute(1, 0, 2).reshape(B, R * k)
    flat_i = pi.permute(1, 0, 2).reshape(B, R * k)
    rs, rj = torch.topk(flat_s, k, dim=1)
    assert torch.equal(s, rs) and torch.equal(i, torch.gather(flat_i, 1, rj))


# ------------------------------------------------------------------------------- full-size properties
def test_full_size_properties_1m_items():
    """North-star shape (8x8x32, 1M items): properties that need no CPU oracle at this size."""
    cfg = CFG_8x8x32
    N, B, k = 1_000_000, 8, 100
    mol, _ = build_module(cfg, None, DEV, seed=0)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 0, DEV)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
    s, got = top(q, k=k)
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    assert all(len(set(r.tolist())) == k for r in got)
    # (1) prefix property: top-10 is the prefix of top-100
    s10, got10 = top(q, k=10)
    assert torch.equal(got10, got[:, :10]) and torch.allclose(s10, s[:, :10], atol=0, rtol=0)
    # (2) returned scores are the exact scores of the returned items (exact fp32 kernel on the gathered items)
    sub = MoLBruteForceTopK(mol, items[(got[0] - 1)].unsqueeze(0), got[0].unsqueeze(0), mode=_lib.MODE_EXACT)
    s_sub, ids_sub = sub(q[:1], k=k)
    assert torch.equal(ids_sub, got[:1]) and (s_sub - s[:1]).abs().max().item() < 1e-4
    # (3) exact mode agrees with the tensor-core + rescoring mode on a query subset
    ex = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)
    s_ex, got_ex = ex(q[:2], k=k)
    assert (s_ex - s[:2]).abs().max().item() < 1e-4
    same = (got_ex == got[:2]).float().mean().item()
    assert same == 1.0, same
    # (4) sharding: union of per-shard top-k merged == unsharded
    R = 4
    parts_s, parts_i = [], []
    for r in range(R):
        lo, hi = r * N // R, (r + 1) * N // R
        sh = MoLBruteForceTopK(mol, items[lo:hi].unsqueeze(0), ids[lo:hi].unsqueeze(0))
        a, b = sh(q, k=k)
        parts_s.append(a)
        parts_i.append(b)
    ms, mi = engine.merge_topk(torch.stack(parts_s), torch.stack(parts_i), k)
    assert torch.equal(mi, got) and torch.equal(ms, s)
    # (5) spot check against the CPU oracle for one query on a 50k-item slice containing its top items
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    sl = torch.cat([got[0].cpu() - 1, torch.arange(0, 50000)]).unique()
    _, _, sc = O.brute_force_top_k(cfg, sd, q[:1].cpu(), items[sl.to(DEV)].cpu(), ids[sl.to(DEV)].cpu(), k)
    r = O.compare_top_k(s[:1], got[:1], sc, ids[sl.to(DEV)].cpu(), k, SCORE_TOL, TIE_TOL)
    assert r["max_score_err"] <= SCORE_TOL, r


@pytest.fixture
def force_filter(monkeypatch):
    """The fused-filter strategy is used from 2^24 (query, item) pairs; these tests reach it with a handful of queries."""
    monkeypatch.setenv("MOL_B200_FILTER_MIN_PAIRS", "0")


@pytest.mark.parametrize("order", ["descending", "ascending"])
def test_filter_strategy_survives_adversarial_item_order(order, force_filter):
    """The fused candidate filter takes its per-query threshold from the first items of the corpus.  Sorting the
    corpus by one query's score makes that sample as unrepresentative as possible (threshold far too high /
    far too low): the safety check must notice and the exact fallback must still return the right answer."""
    cfg = CFG_8x8x32
    N, B, k = 300_000, 4, 100
    mol, _ = build_module(cfg, None, DEV, seed=21)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 21, DEV)
    ex0 = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)
    all0 = mol(q[:1], items.unsqueeze(0))[0][0]                      # (N,) exact scores of query 0
    perm = torch.argsort(all0, descending=(order == "descending"))
    items_p, ids_p = items[perm].contiguous(), ids[perm].contiguous()
    top = MoLBruteForceTopK(mol, items_p.unsqueeze(0), ids_p.unsqueeze(0), mode=_lib.MODE_AUTO)
    s, got = top(q, k=k)
    s_ex, got_ex = ex0(q, k=k)
    assert torch.equal(got, got_ex)
    assert (s - s_ex).abs().max().item() < 1e-4


@pytest.mark.parametrize("N,B,k", [(400_003, 2, 2500), (262_144, 1, 1), (999_963, 3, 200)])
def test_filter_strategy_edge_sizes_match_exact_mode(N, B, k, force_filter):
    """Large k (capacity 4 K' > 4096), ragged last tile, single query, k = 1 on the fused-filter strategy: the tensor
    path must return exactly what the fp32 exact mode returns."""
    cfg = CFG_8x8x32
    mol, _ = build_module(cfg, None, DEV, seed=31)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 31, DEV)
    a = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_AUTO)
    e = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=_lib.MODE_EXACT)
    s, i = a(q, k=k)
    s_ex, i_ex = e(q, k=k)
    same = (i == i_ex).float().mean().item()
    assert (s - s_ex).abs().max().item() < 1e-4
    if same < 1.0:  # only exact ties / fp32-reordered near-ties may differ
        bad = (i != i_ex)
        assert (s[bad] - s_ex[bad]).abs().max().item() < 1e-5, same


# ------------------------------------------------------------------------------- tensor-core coarse pass
@pytest.mark.parametrize(
    "cfg,N,B,seed",
    [(CFG_8x8x32, 128, 1, 1), (CFG_8x8x32, 5000, 7, 2), (CFG_8x8x32, 40000, 33, 3), (CFG_8x4x64, 3000, 5, 4),
     (CFG_8x4x128, 3000, 6, 5)],
)
def test_coarse_pass_matches_its_numerics_model(cfg, N, B, seed):
    """The raw tcgen05 output against tests/sim_coarse.py (same fp16 rounding points, exact transcendental
    functions): differences are only accumulation order + tanh.approx/ex2.approx error."""
    from tests.sim_coarse import coarse_scores

    mol, _ = build_module(cfg, None, DEV, seed=seed)
    items, ids, q, uid = synthetic_inputs(cfg, N, B, seed, DEV)
    w = mol.packed_weights(torch.device(DEV))
    idx = mol.build_index(items, ids)
    got = engine.score_all(w, idx, mol.workspace(torch.device(DEV)), q, uid, coarse=True).cpu()
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    ucpu = None if uid is None else uid.cpu()
    knobs = _lib.build_knobs()  # a tuning variant (MOL_B200_LIB) is compared with the model of ITS activation forms
    sim = coarse_scores(cfg, sd, q.cpu(), items.cpu(), ucpu, e2_h2_mask=knobs.get("e2h2", 0),
                        e3_h2_of4=knobs.get("e3h2", 0), lite=bool(knobs.get("h2lite", 0)))
    exact = O.similarity(cfg, sd, q.cpu(), items.cpu(), ucpu)
    err_sim = (got - sim).abs().max().item()
    err_exact = (got - exact).abs().max().item()
    assert torch.isfinite(got).all()
    assert err_sim < 0.03, (err_sim, err_exact)
    assert err_exact < 0.08, err_exact


def test_coarse_pass_uneven_query_split_and_many_ctas():
    """bc not a multiple of 2, fewer units than SMs, and a tile range that splits queries across CTAs."""
    cfg = CFG_8x8x32
    mol, _ = build_module(cfg, None, DEV, seed=9)
    dev = torch.device(DEV)
    for N, B in [(300, 3), (128 * 150, 1), (128 * 37 + 5, 9)]:
        items, ids, q, _ = synthetic_inputs(cfg, N, B, 11, DEV)
        idx = mol.build_index(items, ids)
        w = mol.packed_weights(dev)
        a = engine.score_all(w, idx, mol.workspace(dev), q, None, coarse=True)
        e = engine.score_all(w, idx, mol.workspace(dev), q, None)
        assert (a - e).abs().max().item() < 0.08


# ------------------------------------------------------------------------------- callers / siblings
def _next_golden(name):
    import os

    import numpy as np

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    return {k_: (int(z[k_]) if k_ == "k" else torch.from_numpy(z[k_])) for k_ in z.files}


@pytest.mark.parametrize("name", ["next_mask_basic", "next_mask_short_rows", "next_mask_wide"])
def test_mips_and_candidate_index_match_reference(name):
    from rails_b200.indexing.candidate_index import CandidateIndex
    from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK

    g = _next_golden(name)
    items, ids, q = g["items"].to(DEV), g["ids"].to(DEV), g["queries"].to(DEV)
    top = MIPSBruteForceTopK(items.unsqueeze(0), ids.unsqueeze(0))
    kp = g["ref_prime_ids"].size(1)
    s, i = top(q, k=kp)
    assert torch.equal(i.cpu(), g["ref_prime_ids"])                      # MIPS top-k' ids bit-exact
    assert (s.cpu() - g["ref_prime_scores"]).abs().max().item() < 1e-5
    index = CandidateIndex(ids=ids.unsqueeze(0), embeddings=items.unsqueeze(0))
    out_ids, out_scores, emb = index.get_top_k_outputs(
        query_embeddings=q, k=g["k"], aux_payloads={}, top_k_module=top, invalid_ids=g["invalid_ids"].to(DEV)
    )
    assert emb is None
    assert torch.equal(out_ids.cpu(), g["ref_ids"])                      # masking + back-fill bit-exact
    assert (out_scores.cpu() - g["ref_scores"]).abs().max().item() < 1e-5


def test_select_valid_matches_oracle_random():
    from oracle import next_oracle as NO

    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    for B, kp, n0, k in [(3, 50, 0, 50), (8, 400, 211, 200), (2, 2711, 211, 2500), (5, 64, 70, 10)]:
        ids = torch.stack([torch.randperm(5000, generator=g)[:kp] + 1 for _ in range(B)])
        scores = torch.sort(torch.randn(B, kp, generator=g), dim=1, descending=True).values
        inv = torch.randint(0, 5001, (B, max(n0, 1)), generator=g)
        if n0 > 0:
            inv[:, : min(n0, kp) // 2] = ids[:, : min(n0, kp) // 2]
        else:
            inv = inv[:, :0]
        rs, ri = NO.select_valid(scores, ids, inv, k) if n0 > 0 else (scores[:, :k], ids[:, :k])
        ds, di, dinv = scores.to(DEV), ids.to(DEV), inv.to(DEV).contiguous()
        out_s = torch.empty(B, k, device=DEV)
        out_i = torch.empty(B, k, dtype=torch.int64, device=DEV)
        import ctypes

        _lib.check(lib.mol_select_valid(
            ctypes.c_void_p(ds.data_ptr()), ctypes.c_void_p(di.data_ptr()), ctypes.c_void_p(dinv.data_ptr() if n0 else 0),
            B, kp, n0, k, ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert torch.equal(out_i.cpu(), ri) and torch.equal(out_s.cpu(), rs)


def test_mol_candidate_index_with_seen_items_matches_oracle():
    """The full eval-side call: CandidateIndex.get_top_k_outputs over MoLBruteForceTopK with per-row seen ids."""
    from oracle import next_oracle as NO
    from rails_b200.indexing.candidate_index import CandidateIndex

    cfg = CFG_8x8x32
    N, B, k, n0 = 5000, 6, 20, 16
    mol, _ = build_module(cfg, None, DEV, seed=8)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 8, DEV)
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    ps, pi, all_scores = O.brute_force_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), k + n0)
    inv = pi[:, ::2][:, :n0].contiguous()  # every other of the row's best items has been "seen"
    rs, ri = NO.select_valid(ps, pi, inv, k)
    index = CandidateIndex(ids=ids.unsqueeze(0), embeddings=items.unsqueeze(0))
    out_ids, out_scores, _ = index.get_top_k_outputs(q, k, {}, top, inv.to(DEV))
    assert torch.equal(out_ids.cpu(), ri)
    assert (out_scores.cpu() - rs).abs().max().item() < SCORE_TOL


@pytest.mark.parametrize("name", ["next_avg_8x8x32", "next_avg_8x4x64_uid"])
def test_mol_avg_top_k_matches_reference(name):
    from rails_b200.indexing.mol_top_k import MoLAvgTopK
    from tests.test_next_oracle_golden import load_avg

    g = load_avg(name)
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    top = MoLAvgTopK(mol, g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0), g["avg_top_k"])
    kw = {} if g["user_ids"] is None else {"user_ids": g["user_ids"].to(DEV)}
    s, i = top(g["queries"].to(DEV), k=g["k"], **kw)
    assert torch.equal(i.cpu(), g["ref_ids_f32"])
    assert (s.cpu() - g["ref_scores_f32"]).abs().max().item() < SCORE_TOL
    with pytest.raises(ValueError, match="avg_top_k"):
        top(g["queries"].to(DEV), k=g["avg_top_k"] + 1, **kw)


def test_mol_avg_top_k_recall_vs_brute_force_large():
    """At scale the approximate module must agree with the oracle restatement and mostly with brute force."""
    from oracle import next_oracle as NO
    from rails_b200.indexing.mol_top_k import MoLAvgTopK

    cfg = CFG_8x8x32
    N, B, k, a = 30000, 8, 20, 512
    mol, _ = build_module(cfg, None, DEV, seed=4)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 4, DEV)
    s, i = MoLAvgTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), a)(q, k=k)
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    rs, ri, _ = NO.mol_avg_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), k, a)
    same = np.mean([len(set(x.tolist()) & set(y.tolist())) / k for x, y in zip(i.cpu(), ri)])
    assert same >= 0.99, same  # fp32 accumulation order may flip a prefilter near-tie at the avg_top_k boundary
    assert (s.cpu() - rs).abs().max().item() < SCORE_TOL or same < 1.0


@pytest.mark.parametrize("name", ["next_naive_8x8x32", "next_naive_8x4x64_uid", "next_comb_8x8x32", "next_comb_8x4x64_uid"])
def test_mol_naive_comb_top_k_matches_reference(name):
    from rails_b200.indexing.mol_top_k import MoLCombTopK, MoLNaiveTopK
    from tests.test_next_oracle_golden import assert_union_equal_tie_aware, load_avg

    g = load_avg(name)
    mol, _ = build_module(g["cfg"], g["sd"], DEV)
    items, ids = g["items"].to(DEV).unsqueeze(0), g["item_ids"].to(DEV).unsqueeze(0)
    if g["avg_top_k"] > 0:
        top = MoLCombTopK(mol, items, ids, g["avg_top_k"], g["k_per_group"])
    else:
        top = MoLNaiveTopK(mol, items, ids, g["k_per_group"])
    kw = {} if g["user_ids"] is None else {"user_ids": g["user_ids"].to(DEV)}
    s, i = top(g["queries"].to(DEV), k=10, **kw)
    # (candidates 1.3e-6 apart in the reference swap places between the tensor-core and the CUDA-core index build)
    assert_union_equal_tie_aware(s.cpu(), i.cpu(), g["ref_scores_f32"], g["ref_ids_f32"], SCORE_TOL)


def test_mol_naive_comb_top_k_large_vs_oracle():
    """Chunked per-group selection at a size where the (rows, N) dot-product matrix is split over several launches."""
    from oracle import next_oracle as NO
    from rails_b200.indexing.mol_top_k import MoLCombTopK, MoLNaiveTopK

    cfg = CFG_8x8x32
    N, B, kpg, a = 60000, 6, 8, 300
    mol, _ = build_module(cfg, None, DEV, seed=6)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 6, DEV)
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    for top, ref in (
        (MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), kpg), NO.mol_naive_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), kpg)),
        (MoLCombTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), a, kpg), NO.mol_comb_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), a, kpg)),
    ):
        s, i = top(q, k=10)
        rs, ri = ref
        # candidate sets agree up to fp32 near-ties at a selection boundary; the best items are identical
        for b in range(B):
            nv, rnv = int((s[b] > -32767.0).sum()), int((rs[b] > -32767.0).sum())
            assert abs(nv - rnv) <= 2
            assert torch.equal(i[b, :50].cpu(), ri[b, :50])
            assert (s[b, :50].cpu() - rs[b, :50]).abs().max().item() < SCORE_TOL
    with pytest.raises(NotImplementedError):
        MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), kpg, use_faiss=True)


def test_get_top_k_module_runs_every_family():
    """indexing/utils_rails.get_top_k_module (the reference's dispatch by method name) end to end on the GPU."""
    import types

    from rails_b200.indexing.utils_rails import get_top_k_module

    cfg = CFG_8x8x32
    N, B, k = 3000, 5, 10
    mol, _ = build_module(cfg, None, DEV, seed=10)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 10, DEV)
    model = types.SimpleNamespace(_ndp_module=mol)
    exact_s, exact_i = get_top_k_module("MoLBruteForceTopK", model, items.unsqueeze(0), ids.unsqueeze(0))(q, k=N)
    score_of = torch.empty((B, N + 1), device=DEV)
    score_of.scatter_(1, exact_i, exact_s)  # exact MoL score of every item id, per query
    for name in ("MoLNaiveTopK5", "MoLAvgTopK500", "MoLCombTopK5_100"):
        s, i = get_top_k_module(name, model, items.unsqueeze(0), ids.unsqueeze(0))(q, k=k)
        assert s.size(0) == B and s.size(1) >= k and i.dtype == torch.int64
        # every (score, id) an approximate module returns is the exact MoL score of that item (duplicates: -32767)
        real = s > -32767.0
        assert (s - torch.gather(score_of, 1, i))[real].abs().max().item() < 1e-5, name
        assert bool((s[:, 1:] <= s[:, :-1]).all())
    s, i = get_top_k_module("MIPSBruteForceTopK", model, items.unsqueeze(0), ids.unsqueeze(0))(q, k=k)
    assert s.shape == (B, k)


def test_chunked_score_matrix_paths_equal_single_chunk(monkeypatch):
    """MIPS / Avg / Naive / Comb materialise a (rows, N) fp32 matrix per launch and loop over query chunks when it would
    exceed its byte budget.  With the budget shrunk (MOL_B200_SCORE_MATRIX_BYTES) the same calls take several chunks and
    must return exactly what the single-chunk run returned."""
    from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
    from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLCombTopK, MoLNaiveTopK

    cfg = CFG_8x8x32
    N, B = 6000, 11
    mol, _ = build_module(cfg, None, DEV, seed=8)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 8, DEV)
    it, idd = items.unsqueeze(0), ids.unsqueeze(0)

    def run_all():
        return [
            MIPSBruteForceTopK(it, idd)(q, k=50),
            MoLAvgTopK(mol, it, idd, 300)(q, k=20),
            MoLNaiveTopK(mol, it, idd, 4)(q, k=20),
            MoLCombTopK(mol, it, idd, 120, 3)(q, k=20),
        ]

    ref = run_all()
    # 6000 items * 4 B = 24 kB per row: 3 query rows (MIPS / Avg) and one query = 8 group rows (Naive / Comb) per launch
    monkeypatch.setenv("MOL_B200_SCORE_MATRIX_BYTES", str(N * 4 * 8 + 64))
    got = run_all()
    for (rs, ri), (s, i) in zip(ref, got):
        assert torch.equal(ri, i) and torch.equal(rs, s)


# ------------------------------------------------------------------------------- multi-GPU (needs >= 2 devices)
def _nccl_worker(rank, world, port, out):
    import os

    import torch.distributed as dist

    from rails_b200.indexing.sharded_top_k import ReplicatedMoLBruteForceTopK, ShardedMoLBruteForceTopK, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        cfg = CFG_8x8x32
        N, B, k = 400_000, 16, 100
        os.environ["MOL_B200_FILTER_MIN_PAIRS"] = "0"  # (16 queries: force the fused-filter strategy)
        mol, _ = build_module(cfg, None, dev, seed=2)
        items, ids, q, _ = synthetic_inputs(cfg, N, B, 2, dev)
        lo, hi = shard_range(N, rank, world)
        local = MoLBruteForceTopK(mol, items[lo:hi].unsqueeze(0), ids[lo:hi].unsqueeze(0))
        s, i = ShardedMoLBruteForceTopK(local, hi - lo)(q, k)
        full = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0))
        rs, ri = full(q, k)
        ok = bool(torch.equal(i, ri)) and bool(torch.equal(s, rs))
        detail = {"shard_ids": bool(torch.equal(i, ri)), "shard_scores": bool(torch.equal(s, rs))}
        # query-split layout over the replicated corpus (ragged: 15 queries over 2 ranks).  Slices of <= 8 queries take the
        # small-batch prologue kernels, whose sums may differ from the 16-query batch in the last bit: ids equal, scores
        # to 1e-4 (a last-bit change of Q_sub is amplified by 1 / tau = 20; the bar is 1e-3)
        s2, i2 = ReplicatedMoLBruteForceTopK(full)(q[:15], k)
        ok = ok and bool(torch.equal(i2, ri[:15])) and float((s2 - rs[:15]).abs().max()) < 1e-4
        detail.update(rep_ids=bool(torch.equal(i2, ri[:15])), rep_err=float((s2 - rs[:15]).abs().max()),
                      rep_mismatch=int((i2 != ri[:15]).sum()))
        # k beyond the corpus: RuntimeError on every rank, before any collective
        try:
            ShardedMoLBruteForceTopK(local, hi - lo)(q, N + 1)
            ok = False
        except RuntimeError as e:
            ok = ok and "out of range" in str(e)
        out[rank] = ok
        out[f"detail{rank}"] = detail
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_topk_two_gpus_equals_unsharded():
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] is True and out[1] is True, dict(out)
