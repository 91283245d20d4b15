#!/usr/bin/env python
"""bench.py — MoL brute-force top-k queries/s (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config north|cfg1|cfg2|cfg3|cfg4|cfg5] [--items N --batch B --k K]
                    [--parallelism auto|replicate|shard]

Default workload (north star): synthetic MoL 8x8x32 (d=32, D=64, H=128, tau=0.05), top-100 over a 1M-item
corpus, batch of 512 queries per step.  A "step" is one MoLBruteForceTopK.forward over the batch.
With N>1 there are two layouts (rails_b200/indexing/sharded_top_k.py), both with ONE all-gather on the data path:
  replicate: every rank holds the corpus and searches its slice of the query batch (corpora that fit one GPU);
  shard:     every rank holds a contiguous item range, per-shard top-k lists are merged (BASELINE configs 4, 5).

One JSON line on rank 0:  value = whole-job queries/s with inputs resident in HBM; e2e = the same through the
host entry (pinned host queries in, scores/ids out, copies inside the timed region); roofline = the dominant scoring
kernel against MEASURED_PEAKS.json; cpu_baseline = the CPU oracle (port of the reference's PyTorch path) timed on this
box's host cores on a bounded sample, and the same sample is used to CHECK the GPU result (`parity`).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "MoL brute-force top-k queries/sec over N-item corpus"
UNIT = "queries/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="north", choices=["north", "sweep", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                   help="north: BASELINE.json's north-star workload (its `secondary` block carries configs 1-3, the B sweep of "
                        "SURVEY 8(d) and the next rows); sweep: the same line with the baseline legs skipped; cfgN: that config")
    p.add_argument("--items", type=int, default=None)
    p.add_argument("--batch", type=int, default=None)
    p.add_argument("--k", type=int, default=None)
    p.add_argument("--mode", default="auto", choices=["auto", "exact", "tensor"])
    p.add_argument("--parallelism", default="auto", choices=["auto", "replicate", "shard"])
    p.add_argument("--cpu-queries", type=int, default=32, help="queries in the cpu_baseline / parity sample (~11 s of CPU work)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-secondary", action="store_true", help="skip the secondary configs / batch sweep block")
    p.add_argument("--no-cuda-graph", action="store_true", help="N > 1: launch each rank's search eagerly instead of as a CUDA graph")
    p.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch-on-this-GPU baseline leg")
    args = p.parse_args()
    if args.config == "sweep":  # the north-star line without the CPU / eager-GPU baseline legs: its secondary block IS the sweep
        args.config = "north"
        args.no_cpu_baseline = True
        args.no_gpu_eager = True
    return args


# ---------------------------------------------------------------------------------------------- workloads
def workload(name, args):
    """(cfg, N, B, k, fixture or None, label) of a BASELINE.json config."""
    from rails_b200.workloads import CFG_8x8x32, CFG_16x16x64, MoLConfig

    if name == "cfg1":  # real ML-1M checkpoint + item table (tests/golden/cfg1_ml1m_ckpt.npz), batch 1
        from tests.golden_util import load_golden

        g = load_golden("cfg1_ml1m_ckpt")
        return g["cfg"], g["items"].size(0), 1, 10, g, "ML-1M HSTU+MoL 8x4x64 ckpt"
    table = {
        "north": (CFG_8x8x32, 1_000_000, 512, 100, "synthetic MoL 8x8x32 d=32 D=64 H=128"),
        "cfg2": (MoLConfig(256, 256, 128, 8, 4, 0.05, "swiglu", (16384,)), 27_278, 128, 100, "ML-20M MoL 8x4x128 (synthetic weights)"),
        "cfg3": (CFG_8x8x32, 695_762, 256, 200, "Amazon-Books MoL 8x8x32 (synthetic weights)"),
        "cfg4": (CFG_8x8x32, 10_000_000, 512, 100, "synthetic MoL 8x8x32 d=32"),
        "cfg5": (CFG_16x16x64, 100_000_000, 1024, 100, "synthetic MoL 16x16x64 d=64"),
    }
    cfg, N, B, k, label = table[name]
    return cfg, N, B, k, None, label


def flops_per_pair(cfg):
    L = cfg.num_logits
    return 2 * L * cfg.dot_product_dimension + 4 * L * 128  # F = 2*L*d + 4*L*H


def bytes_per_item(cfg):
    return 2 * (cfg.item_dot_product_groups * cfg.dot_product_dimension + cfg.num_logits)  # 16-bit X_sub + GI


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (nvidia-smi's clocks line via NVML)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


# ---------------------------------------------------------------------------------------------- baselines (the checker)
def cpu_oracle_run(cfg, sd, items_cpu, ids_cpu, queries_cpu, k, user_ids=None, chunk=2):
    """Runs the CPU oracle (oracle/mol_oracle.py, the port of the reference's PyTorch path) with all host threads.
    Returns (seconds, top scores, top ids, all scores)."""
    from oracle import mol_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    with torch.inference_mode():
        t0 = time.perf_counter()
        s, i, all_scores = O.brute_force_top_k(cfg, sd, queries_cpu, items_cpu, ids_cpu, k, user_ids, chunk=chunk)
        dt = time.perf_counter() - t0
    return dt, s, i, all_scores


def gpu_eager_baseline(cfg, sd_dev, items, ids, queries, k, nq, chunk=2, reps=3):
    """The reference's arithmetic as eager PyTorch ON THIS GPU (the oracle's ATen ops issued on CUDA tensors, queries
    chunked with the reference's user_max_batch_size semantics, data/eval.py:131-138): fp32 with TF32 off, and bf16 as
    eval_from_checkpoint.py:318-322 casts the model.  A reported baseline, timed with CUDA events."""
    from oracle import mol_oracle as O

    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            q = queries[:nq]

            def run():
                return O.brute_force_top_k(cfg, sd_dev, q, items, ids, k, None, dtype, chunk=chunk)

            with torch.inference_mode():
                run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"value": nq / (ms * 1e-3), "unit": UNIT, "ms_per_sample": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out["sample"] = f"{nq} queries x the full corpus per run, chunks of {chunk} queries, eager torch CUDA ops, TF32 off"
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port: the reference
    is pure PyTorch and /root/reference does not exist on the GPU box), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rails_b200.workloads import build_module, synthetic_inputs

    cfg, N, B, k, fixture, label = workload(args.config, args)
    N, B, k = args.items or N, args.batch or B, args.k or k
    mol, _ = build_module(cfg, None if fixture is None else fixture["sd"], "cpu", seed=0)
    sd = {kk: v.detach() for kk, v in mol.state_dict().items()}
    # exactly --steps K timed steps after --warmup W; the sample per step (queries against the full corpus) shrinks with
    # K + W so that the run stays within a few minutes (~0.34 s of CPU per query on 16 cores at 1M items)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = max(1, int(240e6 / max(N, 1)))  # queries affordable in total
    nq = max(1, min(8, B, budget // (steps + warm)))
    Ncpu = min(N, 2_000_000)  # (cfg4 / cfg5: a capped corpus, throughput in pairs/s is what extrapolates)
    if fixture is None:
        items, ids, q, uid = synthetic_inputs(cfg, Ncpu, nq, 0, "cpu")
    else:
        items, ids, q, uid = fixture["items"], fixture["item_ids"], fixture["queries"][:nq], fixture["user_ids"][:nq]
    for _ in range(warm):
        cpu_oracle_run(cfg, sd, items, ids, q, k, uid)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_oracle_run(cfg, sd, items, ids, q, k, uid)
    dt = (time.perf_counter() - t0) / steps
    qps = nq / dt * (Ncpu / N)
    cores = torch.get_num_threads()
    sample = f"{nq} queries x {Ncpu} items per step (chunks of 2 queries), fp32, eager"
    if Ncpu != N:
        sample += f"; queries/s scaled by {Ncpu}/{N} to the full corpus"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(label, N, B, k), "sample": sample},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(label, N, B, k):
    return f"{label}, top-{k} over {N} items, batch={B}"


# ---------------------------------------------------------------------------------------------- secondary block
def time_search(engine, weights, index, wsp, q, uid, k, mode, steps, warmup):
    for _ in range(warmup):
        engine.search(weights, index, wsp, q, uid, k, True, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        engine.search(weights, index, wsp, q, uid, k, True, mode)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def time_module(top, q, k, kw, steps, warmup):
    """ms per forward() of a top-k module (the public call, output tensors included)."""
    for _ in range(warmup):
        top(q, k=k, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        top(q, k=k, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def secondary_block(dev, mode, north=None):
    """BASELINE.json configs 1-3 and the B sweep of SURVEY.md §8(d) (1M items, top-100), a few timed steps each."""
    from rails_b200 import engine
    from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
    from rails_b200.workloads import build_module, synthetic_inputs

    out = {}
    for name in ("cfg1", "cfg2", "cfg3"):
        cfg, N, B, k, fixture, label = workload(name, None)
        mol, _ = build_module(cfg, None if fixture is None else fixture["sd"], dev, seed=0)
        if fixture is None:
            items, ids, q, uid = synthetic_inputs(cfg, N, B, 0, dev)
        else:
            items, ids = fixture["items"].to(dev), fixture["item_ids"].to(dev)
            q, uid = fixture["queries"][:B].to(dev), fixture["user_ids"][:B].to(dev)
        top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode)
        ms = time_search(engine, mol.packed_weights(dev), top._ensure_index(), mol.workspace(dev), q, uid, k, mode, 20, 5)
        out[name] = {"workload": workload_name(label, N, B, k), "ms_per_step": ms, "queries_per_s": B / (ms * 1e-3),
                     "tflops_algorithmic": B * N * flops_per_pair(cfg) / (ms * 1e-3) / 1e12,
                     "hbm_gbs_algorithmic": N * bytes_per_item(cfg) / (ms * 1e-3) / 1e9}
        if B <= 128:  # launch-bound configs: the same call replayed as one CUDA graph (MoLBruteForceTopK(cuda_graph=True))
            gtop = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode, cuda_graph=True)
            kw = {} if uid is None else {"user_ids": uid}
            gms = time_module(gtop, q, k, kw, 50, 5)
            out[name]["cuda_graph"] = {"ms_per_step": gms, "queries_per_s": B / (gms * 1e-3)}
            del gtop
        del top, items
    if north is not None:
        weights, index, wsp, q_dev, k = north[:5]
        sweep = {}
        for b in (1, 8, 32, 128, 512):
            if b > q_dev.size(0):
                continue
            ms = time_search(engine, weights, index, wsp, q_dev[:b].contiguous(), None, k, mode, 10 if b >= 128 else 30, 3)
            sweep[str(b)] = {"ms_per_step": ms, "queries_per_s": b / (ms * 1e-3),
                             "hbm_gbs_algorithmic": index.N * 640 / (ms * 1e-3) / 1e9}
            if b <= 32:
                g = engine.GraphedSearch(weights, index, b, k, mode, False)
                qb = q_dev[:b].contiguous()
                for _ in range(5):
                    g(qb, None)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(30):
                    g(qb, None)
                e1.record()
                torch.cuda.synchronize()
                gms = e0.elapsed_time(e1) / 30
                sweep[str(b)]["cuda_graph"] = {"ms_per_step": gms, "queries_per_s": b / (gms * 1e-3),
                                               "hbm_gbs_algorithmic": index.N * 640 / (gms * 1e-3) / 1e9}
                del g
        out["batch_sweep_1m_top100"] = sweep
        if len(north) > 5:
            out["next_rows_1m"] = next_rows_block(dev, north[5], index, q_dev, k)
    return out


def next_rows_block(dev, mol, index, q_dev, k):
    """SURVEY.md §8 rows f1-f4 on the north-star corpus (1M items, 8x8x32): index build, seen-item exclusion inside the
    search vs the over-fetch + mask recipe, the approximate MoL modules and MIPS - the streaming tensor-core paths and
    (for reference) the materialised-matrix paths they replace (MOL_B200_DOTFILTER=0 / MOL_B200_INDEX_X3=0)."""
    from rails_b200 import engine
    from rails_b200.indexing.candidate_index import CandidateIndex
    from rails_b200.indexing.mips_top_k import MIPSBruteForceTopK
    from rails_b200.indexing.mol_top_k import MoLAvgTopK, MoLBruteForceTopK, MoLCombTopK, MoLNaiveTopK

    items, ids = index.raw, index.ids
    it3, id2 = items.unsqueeze(0), ids.unsqueeze(0)
    out = {}

    def timed(fn, steps=5, warmup=2):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    w = mol.packed_weights(dev)
    saved = {v: os.environ.get(v) for v in ("MOL_B200_DOTFILTER", "MOL_B200_INDEX_X3")}
    try:
        for tag, env in (("tensor", "1"), ("cuda_core", "0")):
            os.environ["MOL_B200_INDEX_X3"] = env
            out[f"f1_index_build_ms_{tag}"] = timed(lambda: engine.IndexHandle(w, items, ids), 3, 1)
        os.environ["MOL_B200_INDEX_X3"] = "1"
        top = MoLBruteForceTopK(mol, it3, id2)
        ci = CandidateIndex(ids=id2, embeddings=it3)
        B = q_dev.size(0)
        inv = torch.randint(1, index.N + 1, (B, 211), device=dev)
        out["f2_seen_items_B512_n0_211_ms_in_search"] = timed(lambda: ci.get_top_k_outputs(q_dev, k, {}, top, inv), 3, 1)
        out["f2_seen_items_B512_n0_211_ms_overfetch_mask"] = timed(
            lambda: ci.get_top_k_outputs(q_dev, k, {}, top, inv, truncate_k_prime_to=k + 211), 3, 1)
        out["f2_no_seen_items_B512_ms"] = timed(lambda: top(q_dev, k=k), 3, 1)
        for tag, env in (("streaming", "1"), ("matrix", "0")):
            os.environ["MOL_B200_DOTFILTER"] = env
            mips = MIPSBruteForceTopK(it3, id2)
            for b in ((1, 64, 512) if env == "1" else (512,)):
                out[f"f4_mips_top{k}_B{b}_ms_{tag}"] = timed(lambda: mips(q_dev[:b], k=k), 10, 3)
            if env == "1":
                out["f4_mips_stats"] = mips.last_search_stats()
            for name, mod in (("avg2000", MoLAvgTopK(mol, it3, id2, 2000)), ("naive_kpg5", MoLNaiveTopK(mol, it3, id2, 5)),
                              ("comb_kpg5_avg200", MoLCombTopK(mol, it3, id2, 200, 5))):
                out[f"f3_mol_{name}_B64_ms_{tag}"] = timed(lambda: mod(q_dev[:64], k=k), 3, 1)
                if env == "1":
                    out[f"f3_mol_{name}_stats"] = mod.last_search_stats()
            del mips
    finally:
        for v, val in saved.items():
            if val is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = val
    return out


# ---------------------------------------------------------------------------------------------- main arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    from rails_b200 import _lib, engine
    from rails_b200.indexing.mol_top_k import MoLBruteForceTopK
    from rails_b200.workloads import build_module

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        print("bench.py: --gpus N>1 must be launched with torch.distributed.run", file=sys.stderr)
        sys.exit(2)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    mode = {"auto": _lib.MODE_AUTO, "exact": _lib.MODE_EXACT, "tensor": _lib.MODE_TENSOR}[args.mode]

    cfg, N, B, k, fixture, label = workload(args.config, args)
    N, B, k = args.items or N, args.batch or B, args.k or k
    mol, _ = build_module(cfg, None if fixture is None else fixture["sd"], dev, seed=0)
    # layout over the ranks: replicate the corpus (query-split) when the whole index fits one GPU, else shard it
    index_bytes = N * (6 * (cfg.item_dot_product_groups * cfg.dot_product_dimension + cfg.num_logits) + 4 * cfg.item_embedding_dim)
    layout = args.parallelism
    if layout == "auto":
        layout = "shard" if (args.config in ("cfg4", "cfg5") or index_bytes > 40e9) else "replicate"
    if world == 1:
        layout = "single"

    def gen_items(lo, hi, seed_rank):
        g = torch.Generator(device=dev).manual_seed(1 + seed_rank)
        return 0.02 * torch.randn(hi - lo, cfg.item_embedding_dim, device=dev, generator=g)

    uid_dev = None
    if fixture is not None:
        items, ids = fixture["items"].to(dev), fixture["item_ids"].to(dev)
        q_host = fixture["queries"][:B].contiguous().pin_memory()
        uid_dev = fixture["user_ids"][:B].to(dev)
        lo, hi = 0, N
    else:
        # corpus (shard) of this rank, generated on device: shard r of an R-way split uses seed 1 + r,
        # the replicated / single layouts seed 1 for the whole corpus
        if layout == "shard":
            lo, hi = rank * N // world, (rank + 1) * N // world
            items = gen_items(lo, hi, rank)
        else:
            lo, hi = 0, N
            items = gen_items(0, N, 0)
        ids = torch.arange(lo + 1, hi + 1, dtype=torch.int64, device=dev)
        gq = torch.Generator().manual_seed(100)
        q_host = F.layer_norm(torch.randn(B, cfg.query_embedding_dim, generator=gq), (cfg.query_embedding_dim,)).pin_memory()
        if cfg.uid_embedding_hash_sizes:
            uid_dev = torch.randint(1, 100000, (B,), generator=gq, dtype=torch.int64).to(dev)
    q_dev = q_host.to(dev)
    # N > 1: each rank's search is a few ms, so its ~30 launches are replayed as one CUDA graph (the public
    # MoLBruteForceTopK(cuda_graph=True) option); the single-GPU legs below call the engine directly
    top = MoLBruteForceTopK(mol, items.unsqueeze(0), ids.unsqueeze(0), mode=mode, cuda_graph=(world > 1 and not args.no_cuda_graph))
    index = top._ensure_index()
    weights = mol.packed_weights(dev)
    wsp = mol.workspace(dev)
    out_s_host = torch.empty((B, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((B, k), dtype=torch.int64).pin_memory()
    uid_host = None if uid_dev is None else uid_dev.cpu().pin_memory()
    tensor_path = engine.tensor_path_supported(weights.shape) and mode != _lib.MODE_EXACT
    kw = {} if uid_dev is None else {"user_ids": uid_dev}

    multi = None
    if world > 1:
        from rails_b200.indexing.sharded_top_k import ReplicatedMoLBruteForceTopK, ShardedMoLBruteForceTopK

        multi = ShardedMoLBruteForceTopK(top, hi - lo) if layout == "shard" else ReplicatedMoLBruteForceTopK(top)

    def step_device():
        if world == 1:
            return engine.search(weights, index, wsp, q_dev, uid_dev, k, True, mode)
        return multi(q_dev, k, **kw)

    def step_e2e():
        if world == 1:
            engine.search_host(weights, index, wsp, q_host, uid_host, k, out_s_host, out_i_host, mode)
        else:
            qd = q_host.to(dev, non_blocking=True)
            s, i = multi(qd, k, **kw)
            out_s_host.copy_(s, non_blocking=True)
            out_i_host.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (no profiling events inside the timed region)
    # N > 1: a step is a few ms and the first NCCL all-gathers / graph replays / clock ramp of a fresh process are not; with
    # only W = 3 short warm-up steps the 8-GPU device number came out BELOW the later-measured end-to-end one (5.14 vs 4.36
    # ms per step).  At least 20 untimed steps (>= 100 ms) precede the K timed ones there.
    n_warm = args.warmup if world == 1 else max(args.warmup, 20)
    for _ in range(n_warm):
        res_s, res_i = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.mol_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res_s, res_i = step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.mol_launch_count())
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = B / (ms_step * 1e-3)
    stats = engine.search_stats(wsp) if world == 1 else top.last_search_stats()  # counters of the last timed search on this rank
    tensor_path = bool(stats.get("tensor_path", tensor_path))  # (MODE_AUTO serves tiny searches with the fp32 kernel)

    # ---- the scoring kernel alone: a separate pass with the library's CUDA-event profiling switched on
    prof_steps = min(args.steps, 3)
    graphed, top._cuda_graph = top._cuda_graph, False  # (a replayed graph does not pass through the library's event records)
    lib.mol_profile_enable(1)
    for _ in range(prof_steps):
        step_device()
    torch.cuda.synchronize()
    top._cuda_graph = graphed
    k_ms, k_n = ctypes.c_double(), ctypes.c_int32()
    _lib.check(lib.mol_profile_collect(ctypes.byref(k_ms), ctypes.byref(k_n)))
    lib.mol_profile_enable(0)

    # ---- end to end through the host entry (pinned host buffers, copies inside the timed region)
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = {
        "value": B / (ms_e2e * 1e-3), "unit": UNIT,
        "h2d_bytes_per_step": q_host.numel() * 4 + (0 if uid_host is None else uid_host.numel() * 8),
        "d2h_bytes_per_step": out_s_host.numel() * 4 + out_i_host.numel() * 8,
        "ms_per_step": ms_e2e,
    }
    e2e_ids_equal = bool(torch.equal(out_i_host, res_i.cpu()))

    # ---- roofline of the dominant kernel (scoring pass) on this rank
    tf_peak, hbm_peak, peak_src = peaks()
    n_local = hi - lo
    b_local = B if layout != "replicate" else (rank + 1) * B // world - rank * B // world
    # the timed launches are the scoring kernel(s) of every step (tensor path: threshold pass + main pass, which
    # together score each (query, item) pair exactly once); achieved = algorithmic FLOPs / summed launch time
    kern_ms_step = k_ms.value / prof_steps
    flops_step = b_local * n_local * flops_per_pair(cfg)
    achieved_tf = flops_step / (kern_ms_step * 1e-3) / 1e12 if kern_ms_step > 0 else 0.0
    peak_note = peak_src if tensor_path else peak_src + "; NOTE fp32 CUDA-core kernel measured against the tensor roofline"
    traffic, traffic_note = None, None
    try:  # measured once under ncu (never inside a timed run): profiles/ncu_traffic.json
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            for name, t in json.load(f).items():
                if isinstance(t, dict) and tensor_path and world == 1 and (t.get("batch"), t.get("items"), t.get("k")) == (B, N, k):
                    traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
                    traffic_note = f"{name}; algorithmic {t['algorithmic_bytes']} B; {t['source']}"
    except (OSError, ValueError, KeyError):
        pass
    roofline = {
        "bound": "tensor", "kernel": "mol_coarse_kernel (tcgen05)" if tensor_path else "exact_scores_kernel(fp32)",
        "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved_tf / tf_peak,
        "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_note,
        "kernel_ms_per_step": kern_ms_step, "kernel_launches_per_step": k_n.value / prof_steps,
        "kernel_ms_per_launch": k_ms.value / max(k_n.value, 1),
        "kernel_share_of_step": kern_ms_step / ms_step,
        "hbm_gbs_algorithmic": (n_local * bytes_per_item(cfg)) / (kern_ms_step * 1e-3) / 1e9 if kern_ms_step > 0 else 0.0,
        "hbm_peak_gbs": hbm_peak,
        "co_limit": "MUFU (16 transcendentals/clk/SM) and tensor-pipe occupancy of small-N MMAs bind before the FLOP roofline: DESIGN.md section 4.1",
    }

    # ---- N > 1: the multi-GPU result must equal the single-GPU result (rank 0, outside every timed region)
    multi_check = None
    if world > 1 and rank == 0 and N <= 4_000_000 and fixture is None:
        if layout == "shard":
            full_items = torch.cat([gen_items(r * N // world, (r + 1) * N // world, r) for r in range(world)])
            full_ids = torch.arange(1, N + 1, dtype=torch.int64, device=dev)
            full = MoLBruteForceTopK(mol, full_items.unsqueeze(0), full_ids.unsqueeze(0), mode=mode)
            fs, fi = full(q_dev, k, **kw)
            del full, full_items
        else:
            fs, fi = top(q_dev, k, **kw)
        multi_check = {"ids_equal_single_gpu": bool(torch.equal(fi, res_i)), "scores_equal_single_gpu": bool(torch.equal(fs, res_s)),
                       "n_queries": B}

    # ---- CPU baseline + parity (rank 0, N=1 only): oracle port on a bounded sample of the same workload; the GPU
    #      result for those queries is CHECKED against it
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import mol_oracle as O

        sd = {kk: v.detach().cpu() for kk, v in mol.state_dict().items()}
        nq = max(1, min(args.cpu_queries, B, int(64e6 // max(N, 1)) or 1))  # ~0.34 s of CPU per query per 1M items
        ids_cpu = ids.cpu()
        t, _, _, all_scores = cpu_oracle_run(cfg, sd, items.cpu(), ids_cpu, q_host[:nq].clone(), k,
                                             None if uid_host is None else uid_host[:nq].clone(), chunk=2 if N <= 2_000_000 else 1)
        cpu_baseline = {
            "value": nq / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nq} of the {B} queries x the full {N}-item corpus, fp32 eager torch CPU (oracle/mol_oracle.py), {t:.1f} s",
        }
        r = O.compare_top_k(res_s[:nq], res_i[:nq], all_scores, ids_cpu, k, 1e-3, 1e-4)
        parity = {
            "n_queries": nq, "items": N, "strict_id_match": r["strict_row_match"], "tie_aware_ok": r["ok"],
            "max_score_err": r["max_score_err"], "max_rank_gap": r["max_rank_gap"],
            "checked_against": "CPU oracle (oracle/mol_oracle.py) over the full corpus, same queries as the timed step",
        }

    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager and fixture is None and N <= 4_000_000:
        sd_dev = {kk: v.detach() for kk, v in mol.state_dict().items()}
        try:
            gpu_eager = gpu_eager_baseline(cfg, sd_dev, items, ids, q_dev, k, min(16, B))
        except torch.cuda.OutOfMemoryError as e:  # (reported, not fatal: it is a baseline leg)
            gpu_eager = {"error": str(e)[:200]}

    secondary = None
    if rank == 0 and world == 1 and args.config == "north" and not args.no_secondary and args.items is None and args.batch is None:
        secondary = secondary_block(dev, mode, (weights, index, wsp, q_dev, k, mol))

    if rank == 0:
        par = {"single": "single GPU", "replicate": f"corpus replicated, queries split x{world}, one all-gather",
               "shard": f"corpus sharded x{world}, one all-gather + merge"}[layout]
        if world > 1 and not args.no_cuda_graph:
            par += "; per-rank search replayed as a CUDA graph"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16" if tensor_path else "f32", "data": "synthetic" if fixture is None else "ML-1M checkpoint + synthetic queries",
            "config": {
                "workload": workload_name(label, N, B, k), "parallelism": par,
                "mode": args.mode, "tensor_path": bool(tensor_path),
                "l2_policy": f"inputs larger than L2 (16-bit index: {n_local * bytes_per_item(cfg) / 1e6:.0f} MB per corpus pass)"
                if n_local * bytes_per_item(cfg) > 126e6 else "corpus smaller than L2 (latency-bound config; stated, not flushed)",
                "library": lib.mol_version().decode(),  # build knobs of the loaded libmol_b200 (tuning variants differ)
            },
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "search_stats": stats, "fallback_queries": stats["fallback_queries"], "e2e_ids_equal_device_run": e2e_ids_equal,
            "parity": parity, "multi_gpu_check": multi_check,
            "cpu_baseline": cpu_baseline, "gpu_eager_baseline": gpu_eager, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
