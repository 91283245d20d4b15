"""Runs one tcgen05/TMA/MUFU probe of libmol_probe.so on cuda:0 and prints a JSON verdict.

    python tools/run_probe.py mma 0|1|2
    python tools/run_probe.py tma
    python tools/run_probe.py mufu

Each invocation is a separate process so that a wrong descriptor (illegal instruction / hang) cannot
take the other probes down; wrap in `timeout`.
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "rails_b200", "lib", "libmol_probe.so"))
P = ctypes.c_void_p


def ptr(t):
    return P(t.data_ptr())


def mma(variant):
    K1, N1, N2 = [(32, 16, 16), (64, 128, 64), (128, 64, 128)][variant]
    g = torch.Generator(device="cuda").manual_seed(variant)
    A = torch.randn(128, K1, device="cuda", generator=g).bfloat16()
    B1 = torch.randn(N1, K1, device="cuda", generator=g).bfloat16()
    B2 = (torch.randn(N2, N1, device="cuda", generator=g) * 0.1).bfloat16()
    out1 = torch.full((128, N1), float("nan"), device="cuda")
    out2 = torch.full((128, N2), float("nan"), device="cuda")
    rc = lib.probe_mma(variant, ptr(A), ptr(B1), ptr(B2), ptr(out1), ptr(out2))
    ref1 = A.float() @ B1.float().t()
    ref2 = ref1.bfloat16().float() @ B2.float().t()
    # if the layout guess is wrong, try to characterise: compare against a few permutations
    e1 = (out1 - ref1).abs().max().item()
    e2 = (out2 - out1.bfloat16().float() @ B2.float().t()).abs().max().item()
    print(json.dumps({"probe": "mma", "variant": variant, "rc": rc, "ss_max_err": e1, "ts_max_err": e2,
                      "ref1_absmax": ref1.abs().max().item(), "ref2_absmax": ref2.abs().max().item()}))
    if not (e2 < 1e-2):
        # diagnose packing order: maybe high half = even k
        sw = out1.bfloat16().float().view(128, N1 // 2, 2).flip(-1).reshape(128, N1)
        e2b = (out2 - sw @ B2.float().t()).abs().max().item()
        print(json.dumps({"probe": "mma", "diag": "swapped bf16 halves", "ts_max_err": e2b}))


def tma():
    rows = 1024
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn(rows, 256, device="cuda", generator=g).bfloat16()
    Q = torch.randn(16, 32, device="cuda", generator=g).bfloat16()
    for tile in (0, 5):
        out = torch.full((8, 128, 16), float("nan"), device="cuda")
        rc = lib.probe_tma(ptr(X), ctypes.c_longlong(rows), ptr(Q), ptr(out), tile)
        xt = X[tile * 128:(tile + 1) * 128].float().view(128, 8, 32)
        ref = torch.einsum("rmk,nk->mrn", xt, Q.float())
        err = (out - ref).abs().amax(dim=(1, 2)).tolist()
        print(json.dumps({"probe": "tma", "tile": tile, "rc": rc, "max_err_per_m": err}))


def mufu():
    names = ["tanh.f32", "tanh.bf16x2", "ex2.f32", "ex2.bf16x2", "ffma", "cvt.bf16x2"]
    sink = torch.empty(148 * 8 * 1024, device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    iters = 4096
    for which, name in enumerate(names):
        for threads in (128, 512, 1024):
            lib.probe_mufu(which, 64, threads, 148, ptr(sink), ptr(cyc))  # warm
            rc = lib.probe_mufu(which, iters, threads, 148, ptr(sink), ptr(cyc))
            c = cyc.item()
            ops = iters * 4 * threads  # per SM (1 block per SM)
            print(json.dumps({"probe": "mufu", "op": name, "threads_per_sm": threads, "rc": rc, "cycles": c,
                              "thread_ops_per_clk_per_sm": ops / c}))


def pipe():
    names = [("ffma", 8), ("ffma2", 8), ("hfma2.f16x2", 8), ("cvt.f16x2", 8), ("min.xorsign.abs.f16x2", 8),
             ("tanh.f16x2", 8), ("mix 2 tanh.f32 + 12 ffma", 14), ("mix 2 ex2.f32 + 6 ffma2", 8),
             ("mix 2 tanh.f16x2 + 8 hfma2", 10), ("fadd2", 8)]
    sink = torch.empty(148 * 8 * 1024, device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    iters = 4096
    for which, (name, per_iter) in enumerate(names):
        for threads in (256, 512):
            lib.probe_pipe(which, 64, threads, 148, ptr(sink), ptr(cyc))  # warm
            rc = lib.probe_pipe(which, iters, threads, 148, ptr(sink), ptr(cyc))
            c = cyc.item()
            print(json.dumps({"probe": "pipe", "op": name, "threads_per_sm": threads, "rc": rc, "cycles": c,
                              "thread_instr_per_clk_per_sm": iters * per_iter * threads / c}))


def tmem():
    names = [("ld.x32 + wait", 4096), ("ld.x16 + wait", 2048), ("st.x32 + wait", 4096), ("4 x ld.x16 then wait", 8192)]
    sink = torch.empty(148 * 1024, device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    iters = 2048
    for which, (name, bytes_per_warp_iter) in enumerate(names):
        for threads in (128, 256, 512):
            lib.probe_tmem(which, 16, threads, 148, ptr(sink), ptr(cyc))
            rc = lib.probe_tmem(which, iters, threads, 148, ptr(sink), ptr(cyc))
            c = cyc.item()
            print(json.dumps({"probe": "tmem", "op": name, "threads_per_sm": threads, "rc": rc, "cycles": c,
                              "cycles_per_iter": c / iters, "bytes_per_clk_per_sm": iters * bytes_per_warp_iter * (threads // 32) / c}))


def mma_rate():
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    iters = 256
    for mode, mname in ((0, "SS"), (1, "TS"), (2, "SS collector fill/lastuse pairs"), (3, "SS pairs, no hint"), (4, "SS collector fill/use/use/lastuse")):
        for n in (16, 32, 64, 128, 256) if mode < 2 else (16, 32):
            lib.probe_mma_rate(n, mode, 8, 148, ptr(cyc))
            rc = lib.probe_mma_rate(n, mode, iters, 148, ptr(cyc))
            c = cyc.tolist()
            print(json.dumps({"probe": "mma_rate", "mode": mname, "N": n, "rc": rc, "issue_clk_per_mma": c[0] / iters,
                              "total_clk_per_mma": c[1] / iters, "floor": 128 * n / 256}))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "mma":
        mma(int(sys.argv[2]))
    elif what == "tma":
        tma()
    elif what == "mma_rate":
        mma_rate()
    elif what == "tmem":
        tmem()
    elif what == "pipe":
        pipe()
    else:
        mufu()
