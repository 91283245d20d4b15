"""GPU tests of the streaming dot-product top-k (rails_b200/csrc/mol_dotfilter_sm100.cu): the fused GEMM + top-k of
SURVEY.md §8 row f4 (MIPSBruteForceTopK, rails/indexing/mips_top_k.py:74-81) and the streaming prefilters of row f3
(MoLAvgTopK / MoLNaiveTopK / MoLCombTopK, rails/indexing/mol_top_k.py:239-249, :352-360, :501-511).

The bar is bit-exactness: the streaming path re-scores its survivors with the same k-ascending fp32 fmaf chain as the
materialised-matrix path, proves per row that no item outside the survivors can reach the top-k, and breaks ties by
position - so ids AND scores must equal the materialised path (MOL_B200_DOTFILTER=0) exactly, and torch's own fp32
result tie-aware."""
import pytest
import torch

from tests.helpers import CFG_8x8x32, CFG_8x4x128, CFG_16x16x64, build_module, synthetic_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mips(items, ids, q, k, monkeypatch, stream: bool):
    from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK

    monkeypatch.setenv("MOL_B200_DOTFILTER", "1" if stream else "0")
    top = MIPSBruteForceTopK(items.unsqueeze(0), ids.unsqueeze(0))
    s, i = top(q, k=k)
    torch.cuda.synchronize()
    return s, i, top.last_search_stats()


def _check_vs_torch(items, ids, q, k, s, i):
    ref = q.double() @ items.double().t()  # fp64 referee: classifies fp32 near-ties
    rs, ri = ref.topk(k, dim=1)
    got = torch.gather(ref, 1, (i - 1))  # ids are positions + 1 in these tests
    # every returned item scores (fp64) within fp32 noise of the fp64 top-k at the same rank
    assert (got - rs).abs().max().item() < 2e-5
    assert (s.double() - got).abs().max().item() < 2e-5
    assert bool((s[:, 1:] <= s[:, :-1]).all())


@pytest.mark.parametrize(
    "N,D,B,k",
    [
        (300_000, 64, 70, 100),     # one column chunk, rows padded to 128
        (262_144 + 77, 64, 300, 100),  # two column chunks, ragged last item tile
        (200_000, 32, 257, 10),     # K = 32 (one box per tile)
        (131_072, 128, 33, 200),    # K = 128
        (100_000, 256, 130, 50),    # K = 256: 128 query rows per launch
        (1_000_000, 64, 1, 1),      # one query, k = 1
        (70_000, 96, 5, 7),         # K = 96: three boxes per tile, 24 lanes per row in the norm pass
        (66_000, 32, 8300, 3),      # more than 8192 query rows: two row chunks of the whole pipeline
    ],
)
def test_mips_streaming_equals_materialised(N, D, B, k, monkeypatch):
    g = torch.Generator(device=DEV).manual_seed(N + D + B)
    items = 0.02 * torch.randn((N, D), device=DEV, generator=g)
    q = torch.nn.functional.layer_norm(torch.randn((B, D), device=DEV, generator=g), (D,))
    ids = torch.arange(1, N + 1, device=DEV)
    s1, i1, st1 = _mips(items, ids, q, k, monkeypatch, True)
    s0, i0, st0 = _mips(items, ids, q, k, monkeypatch, False)
    assert st1["filter_strategy"] == 1 and st0["filter_strategy"] == 0
    assert st1["fallback_queries"] == 0 and st1["filter_overflows"] == 0, st1
    assert torch.equal(i1, i0)
    assert torch.equal(s1, s0)
    _check_vs_torch(items, ids, q, k, s1, i1)


def test_mips_streaming_adversarial_rows_fall_back_and_stay_exact(monkeypatch):
    """Rows the filter cannot serve must come out of the fallback unchanged: a zero query (every value ties at 0: the
    survivor buffer overflows), a query whose best items all sit in ONE item tile that the strided sample never sees
    while everything else scores lower (fine), duplicated items (ties at the top), and a corpus sorted by norm."""
    N, D, B, k = 200_000, 64, 6, 100
    g = torch.Generator(device=DEV).manual_seed(5)
    items = 0.02 * torch.randn((N, D), device=DEV, generator=g)
    order = items.norm(dim=1).argsort()
    items = items[order].contiguous()          # sorted by norm: the big scores cluster at the end
    items[1000:1100] = items[150_000:150_100]  # exact duplicates
    q = torch.nn.functional.layer_norm(torch.randn((B, D), device=DEV, generator=g), (D,))
    q[0] = 0.0                                 # all-zero query
    q[1] = 50.0 * items[199_999] / items[199_999].norm()
    ids = torch.arange(1, N + 1, device=DEV)
    s1, i1, st1 = _mips(items, ids, q, k, monkeypatch, True)
    s0, i0, _ = _mips(items, ids, q, k, monkeypatch, False)
    assert st1["filter_strategy"] == 1
    assert st1["fallback_queries"] >= 1 and st1["filter_overflows"] >= 1, st1  # the zero query
    assert st1["fallback_queries"] <= 2, st1
    # (the zero query ties all 200k items at 0: which 100 of them a select keeps is unspecified, as with torch.topk)
    assert torch.equal(i1[1:], i0[1:])
    assert torch.equal(s1, s0)
    assert bool((s1[0] == 0).all()) and i1[0].unique().numel() == k and int(i1[0].min()) >= 1 and int(i1[0].max()) <= N


def test_mips_small_or_odd_shapes_keep_the_materialised_path(monkeypatch):
    g = torch.Generator(device=DEV).manual_seed(9)
    for N, D in ((3883, 50), (100_000, 50), (20_000, 64)):
        items = 0.02 * torch.randn((N, D), device=DEV, generator=g)
        q = torch.randn((4, D), device=DEV, generator=g)
        ids = torch.arange(1, N + 1, device=DEV)
        s, i, st = _mips(items, ids, q, 10, monkeypatch, True)
        assert st["filter_strategy"] == 0
        _check_vs_torch(items, ids, q, 10, s, i)


@pytest.mark.parametrize("cfg,N,B", [(CFG_8x8x32, 200_000, 24), (CFG_8x4x128, 70_000, 40)])
def test_approximate_modules_streaming_equals_materialised(cfg, N, B, monkeypatch):
    """MoLAvgTopK / MoLNaiveTopK / MoLCombTopK: the streaming prefilters select exactly the candidates the materialised
    (rows, N) matrices select, so the final (scores, ids) are identical."""
    from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLCombTopK, MoLNaiveTopK

    mol, _ = build_module(cfg, None, DEV, seed=21)
    items, ids, q, uid = synthetic_inputs(cfg, N, B, 21, DEV)
    kw = {"user_ids": uid} if uid is not None else {}
    makers = (
        lambda: MoLAvgTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 500),
        lambda: MoLNaiveTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 5),
        lambda: MoLCombTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), 200, 3),
    )
    for make in makers:
        out = {}
        for stream in (True, False):
            monkeypatch.setenv("MOL_B200_DOTFILTER", "1" if stream else "0")
            top = make()
            s, i = top(q, k=10, **kw)
            torch.cuda.synchronize()
            out[stream] = (s, i, top.last_search_stats())
        st = out[True][2]
        assert st["filter_strategy"] == 1 and out[False][2]["filter_strategy"] == 0, st
        assert st["filter_overflows"] == 0, st
        assert torch.equal(out[True][1], out[False][1]), type(top).__name__
        assert torch.equal(out[True][0], out[False][0]), type(top).__name__


@pytest.mark.parametrize("mode_name,N,force_filter", [("auto", 150_000, True), ("auto", 20_000, False), ("exact", 20_000, False)])
def test_in_search_exclusion_equals_overfetch_and_mask(mode_name, N, force_filter, monkeypatch):
    """seen ids excluded INSIDE the search (mol_search_excluding) must give exactly what the
    reference's recipe gives (indexing/candidate_index.py:144-178: over-fetch k' = k + N0, mask, keep the first k) - here
    the over-fetch + mol_select_valid path, itself pinned to the oracle in test_gpu_parity.py."""
    from rails_b200 import _lib
    from rails_b200.indexing.candidate_index import CandidateIndex
    from rails_b200.indexing.mol_top_k import MoLBruteForceTopK

    if force_filter:
        monkeypatch.setenv("MOL_B200_FILTER_MIN_PAIRS", "0")
    cfg = CFG_8x8x32
    B, k, n0 = 9, 50, 37
    mol, _ = build_module(cfg, None, DEV, seed=31)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 31, DEV)
    mode = _lib.MODE_EXACT if mode_name == "exact" else _lib.MODE_AUTO
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode)
    _, best = top(q, k=k + n0)
    g = torch.Generator(device=DEV).manual_seed(3)
    # the list of a query: 20 of its own best items (ranks 0, 2, 4, ...), random ids, duplicates and zero padding
    inv = torch.cat(
        [best[:, 0:40:2], torch.randint(1, N + 1, (B, n0 - 25), device=DEV, generator=g), best[:, 0:2],
         torch.zeros((B, 3), dtype=torch.int64, device=DEV)], dim=1)
    assert inv.size(1) == n0
    index = CandidateIndex(ids=ids.unsqueeze(0), embeddings=items.unsqueeze(0))
    got_i, got_s, _ = index.get_top_k_outputs(q, k, {}, top, inv)
    st = top.last_search_stats()
    ref_i, ref_s, _ = index.get_top_k_outputs(q, k, {}, top, inv, truncate_k_prime_to=k + n0)  # over-fetch + mask
    assert st["filter_strategy"] == (1 if force_filter else 0) and st["fallback_queries"] == 0, st
    assert torch.equal(got_i, ref_i)
    assert torch.equal(got_s, ref_s)
    for b in range(B):
        assert not set(got_i[b].tolist()) & set(inv[b].tolist())


def _index_arrays(ix):
    """fp32 and fp16 caches of an engine.IndexHandle as tensors (copies)."""
    s = ix.weights.shape
    N, nx, L = ix.N, s.item_dot_product_groups * s.dot_product_dimension, s.query_dot_product_groups * s.item_dot_product_groups
    Np = (N + 127) // 128 * 128
    base = ix.blob.data_ptr()

    def view(ptr, count, dtype, width):
        off = ptr - base
        return ix.blob[off : off + count * width].view(dtype).clone()

    return {
        "xsub_f32": view(ix.struct.xsub_f32, N * nx, torch.float32, 4).view(N, nx),
        "gi_f32": view(ix.struct.gi_f32, N * L, torch.float32, 4).view(N, L),
        "xsub_half": view(ix.struct.xsub_half, Np * nx, torch.float16, 2).view(Np, nx),
        "gi_half": view(ix.struct.gi_half, Np * L, torch.float16, 2).view(Np, L),
    }


@pytest.mark.parametrize("cfg,N", [(CFG_8x8x32, 100_000 + 37), (CFG_16x16x64, 20_000 + 1)])
def test_tensor_core_index_build_matches_cuda_core_build(cfg, N, monkeypatch):
    """the tf32 x 3 tensor-core index build (fused l2-norm / silu / fp16 image epilogues) against
    the fp32 CUDA-core build and against the CPU oracle's item side (item_embeddings_fns.py:165-182,
    similarity_fn.py:170-171).  fp32 caches: within fp32 rounding noise; fp16 copies: at most one fp16 ulp apart; the
    pad rows of the fp16 copies are zero."""
    from oracle import mol_oracle as O
    from rails_b200 import engine

    mol, _ = build_module(cfg, None, DEV, seed=41)
    items, ids, _, _ = synthetic_inputs(cfg, N, 1, 41, DEV)
    w = mol.packed_weights(torch.device(DEV))
    out = {}
    for x3 in ("1", "0"):
        monkeypatch.setenv("MOL_B200_INDEX_X3", x3)
        ix = engine.IndexHandle(w, items, ids)
        torch.cuda.synchronize()
        out[x3] = _index_arrays(ix)
    a, b = out["1"], out["0"]
    assert (a["xsub_f32"] - b["xsub_f32"]).abs().max().item() < 2e-6       # unit vectors
    gscale = b["gi_f32"].abs().max().item()
    assert (a["gi_f32"] - b["gi_f32"]).abs().max().item() < 2e-6 * max(gscale, 1.0)
    assert (a["xsub_half"].float() - b["xsub_half"].float()).abs().max().item() <= 2 ** -11 + 1e-7   # one fp16 ulp below 1
    assert (a["gi_half"].float() - b["gi_half"].float()).abs().max().item() <= 2 ** -10 * max(gscale, 1.0)
    assert bool((a["xsub_half"][N:] == 0).all()) and bool((a["gi_half"][N:] == 0).all())
    # oracle item side on a slice
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    sl = slice(N - 3000, N)
    xs_ref = O.item_sub_embeddings(cfg, sd, items[sl].cpu())
    gi_ref = O._mlp_silu(items[sl].cpu(), sd[O.K_GI_W1], sd[O.K_GI_B1], sd[O.K_GI_W2])
    assert (a["xsub_f32"][sl].cpu() - xs_ref.reshape(3000, -1)).abs().max().item() < 5e-6
    assert (a["gi_f32"][sl].cpu() - gi_ref).abs().max().item() < 5e-6 * max(gscale, 1.0)


def test_streaming_modules_vs_cpu_oracle_directly():
    """The streaming prefilters against the CPU restatement of the reference (oracle/next_oracle.py: mol_top_k.py:133-551,
    mips_top_k.py:74-81) at a corpus size that takes the streaming path (>= 64k items) - not only against the materialised
    GPU path.  Candidate sets may differ by fp32 near-ties at a selection boundary; the best items are identical."""
    import numpy as np

    from oracle import next_oracle as NO
    from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
    from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLCombTopK, MoLNaiveTopK

    cfg = CFG_8x8x32
    N, B, kpg, a, k = 80_000, 6, 8, 600, 20
    mol, _ = build_module(cfg, None, DEV, seed=52)
    items, ids, q, _ = synthetic_inputs(cfg, N, B, 52, DEV)
    sd = {k_: v.detach().cpu() for k_, v in mol.state_dict().items()}
    it3, id2 = items.unsqueeze(0), ids.unsqueeze(0)
    # MoLAvgTopK
    top = MoLAvgTopK(mol, it3, id2, a)
    s, i = top(q, k=k)
    assert top.last_search_stats()["filter_strategy"] == 1
    rs, ri, _ = NO.mol_avg_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), k, a)
    same = np.mean([len(set(x.tolist()) & set(y.tolist())) / k for x, y in zip(i.cpu(), ri)])
    assert same >= 0.99, same
    assert (s.cpu() - rs).abs().max().item() < 1e-3 or same < 1.0
    # MoLNaiveTopK / MoLCombTopK
    for top, ref in (
        (MoLNaiveTopK(mol, it3, id2, kpg), NO.mol_naive_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), kpg)),
        (MoLCombTopK(mol, it3, id2, a, kpg), NO.mol_comb_top_k(cfg, sd, q.cpu(), items.cpu(), ids.cpu(), a, kpg)),
    ):
        s, i = top(q, k=10)
        assert top.last_search_stats()["filter_strategy"] == 1
        rs, ri = ref
        for b in range(B):
            nv, rnv = int((s[b] > -32767.0).sum()), int((rs[b] > -32767.0).sum())
            assert abs(nv - rnv) <= 2
            assert torch.equal(i[b, :50].cpu(), ri[b, :50])
            assert (s[b, :50].cpu() - rs[b, :50]).abs().max().item() < 1e-3
    # MIPS: fp64 referee on the CPU
    mips = MIPSBruteForceTopK(it3, id2)
    s, i = mips(q, k=k)
    assert mips.last_search_stats()["filter_strategy"] == 1
    ref = q.double().cpu() @ items.double().cpu().t()
    rs = ref.topk(k, dim=1).values
    got = torch.gather(ref, 1, i.cpu() - 1)
    assert (got - rs).abs().max().item() < 2e-5 and (s.cpu().double() - got).abs().max().item() < 2e-5
