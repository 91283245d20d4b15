#!/usr/bin/env python
"""Per-wait-site stall samples of an .ncu-rep source page: every SYNCS.PHASECHK (mbarrier try_wait) with the samples of
its polling loop, and per-region sample totals between role markers (UTCHMMA = issuer, MUFU.EX2 = E3, MUFU.TANH.F16 = E2).

    python tools/ncu_waits.py gpurun_out/prof_coarse_b512.ncu-rep
"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    src = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[start]; ix = {n: i for i, n in enumerate(h)}
    data = [r for r in src[start + 1:] if len(r) == len(h)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    S = [int(r[ix["# Samples"]] or 0) for r in data]
    T = [r[ix["Source"]].strip() for r in data]
    E = [int(r[ix["Instructions Executed"]] or 0) for r in data]
    print("total samples", tot)
    # wait sites: PHASECHK followed by loop; sum samples from the PHASECHK until the next instruction that is not part of a loop (heuristic: +-6 instrs)
    for i, t in enumerate(T):
        if "SYNCS.PHASECHK" in t:
            lo, hi = max(0, i - 3), min(len(T), i + 8)
            s = sum(S[lo:hi])
            if s > 0.002 * tot:
                print(f"wait@{i:5d} exec={E[i]:>10d} samples={s:>7d} ({100.0*s/tot:4.1f}%)  {t[:90]}")
    # region totals by 100-instruction windows with opcode hints
    print("-- windows of 200 SASS instructions: samples, hints")
    for w in range(0, len(T), 200):
        seg = T[w:w + 200]
        hints = [k for k in ("UTCHMMA", "UTMALDG", "MUFU.EX2", "MUFU.TANH.F16", "MUFU.TANH ", "LDTM", "STTM", "HFMA2", "FFMA2", "ATOMG", "RED", "STG") if any(k in x for x in seg)]
        print(f"{w:5d}: {sum(S[w:w+200]):>8d} ({100.0*sum(S[w:w+200])/tot:4.1f}%) exec~{max(E[w:w+200]):>10d} {' '.join(hints)}")

main()
