"""Design study (CPU only): a MUFU-free silu(2u) for the coarse kernel's E2 / E3 stages, evaluable in packed half2.

    silu(2u) = u + u*tanh(u) = (u + |u|) - s(|u|),      s(a) = a*(1 - tanh a)  in [0, 0.2785], -> 0 for a > ~5

The large part u + |u| = 2*relu(u) is exact in fp16; only the small bump s needs approximating, so its error budget is
absolute, not relative (a polynomial for tanh itself is ill-conditioned in fp16: DESIGN.md 4.5).  A plain polynomial
in a is still poor (a bump with an exponential tail: degree 7 on [0, 4.5] leaves 2e-3 in float64 and 3e-2 once
evaluated in fp16 - `--naive` prints that table).  What works is the compressed variable

    y = relu(1 - a/A),  w = y^p (p = 2 or 4: one or two squarings),  s(a) ~ a * w * Q_m(w)

w lies in [0, 1], is 0 with zero slope where the tail ends (no clamp needed: a > A gives s = 0 exactly), and stretches the
region of the bump, so a degree-3 Q with O(1) coefficients reaches the accuracy of MUFU.TANH.F16 + fp16 rounding.
Kernel sequence per pair of values (packed half2), 11 instructions for p = 4, m = 3, no MUFU:

    a = |u|; y = fma.relu(a, -1/A, 1); y2 = y*y; w = y2*y2; aw = a*w;
    q = fma(fma(fma(-c3, w, -c2), w, -c1), w, -c0); r = u + a; h = fma(aw, q, r)

The script fits Q (Lawson-weighted least squares -> near-minimax), rounds the coefficients to fp16, emulates every HFMA2
rounding, reports the error of h against float64 over every fp16 value of u in [-12, 12], refines the coefficients by a
+-ulp search on the EMULATED error, and prints the fp16 bit patterns that rails_b200/csrc/mol_coarse_sm100.cu uses.

    python tools/fit_silu_h2.py             # candidate table + the chosen constants
    python tools/fit_silu_h2.py --naive     # plain polynomial in a (what does not work)
"""
import itertools
import sys

import numpy as np


def s_true(a):
    return a * (1.0 - np.tanh(a))


def h_true(u):
    return u + u * np.tanh(u)


def f16(x):
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def fma16(a, b, c):
    return f16(a * b + c)  # exact product + sum in float64, one rounding (HFMA2)


def lawson(V, y, iters=80):
    w = np.ones(len(y))
    c = None
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
        e = np.abs(V @ c - y)
        w = w * (0.5 + e / (e.max() + 1e-30))
        w /= w.max()
    return c


def all_f16_in(lo, hi):
    bits = np.arange(0, 1 << 16, dtype=np.uint16)
    v = bits.view(np.float16).astype(np.float64)
    v = v[np.isfinite(v)]
    return v[(v >= lo) & (v <= hi)]


def mufu_path_f16(u):
    """Today's MUFU path with an exact tanh: t = f16(tanh u), h = fma(u, t, u) in fp16 (MUFU.TANH.F16 adds ~5e-4 to t)."""
    u = f16(u)
    return fma16(u, f16(np.tanh(u)), u)


# ---- compressed-variable form ---------------------------------------------------------------------------------------
def fit_q(A, p, m):
    a = np.linspace(0.0, A, 6001)
    w = np.maximum(1.0 - a / A, 0.0) ** p
    V = np.stack([a * w * w**k for k in range(m + 1)], axis=1)
    return lawson(V, s_true(a))


def eval_h_f16(u, c, A, p):
    """The kernel's instruction sequence, one fp16 rounding per instruction."""
    u = f16(u)
    a = np.abs(u)
    y = np.maximum(fma16(a, f16(-1.0 / A), 1.0), 0.0)
    w = y
    for _ in range(int(round(np.log2(p)))):
        w = f16(w * w)
    aw = f16(a * w)
    nc = f16(-np.asarray(c))
    q = np.full_like(a, nc[-1])
    for k in range(len(nc) - 2, -1, -1):
        q = fma16(q, w, nc[k])
    r = f16(u + a)
    return fma16(aw, q, r)


def eval_s_f16(u, c, A, p):
    """E3 variant: only s comes from half2 (the caller combines it with u + |u| in fp32).  Returns s as fp16 values."""
    u = f16(u)
    a = np.abs(u)
    y = np.maximum(fma16(a, f16(-1.0 / A), 1.0), 0.0)
    w = y
    for _ in range(int(round(np.log2(p)))):
        w = f16(w * w)
    aw = f16(a * w)
    cc = f16(np.asarray(c))
    q = np.full_like(a, cc[-1])
    for k in range(len(cc) - 2, -1, -1):
        q = fma16(q, w, cc[k])
    return f16(aw * q)


def refine(c, A, p, u, ht, rounds=3):
    c16 = np.asarray(f16(c), dtype=np.float16)
    best = np.abs(eval_h_f16(u, c16.astype(np.float64), A, p) - ht).max()
    for _ in range(rounds):
        improved = False
        for k in range(len(c16)):
            for step in (1, -1, 2, -2, 3, -3):
                t = c16.copy()
                t[k] = (t[k : k + 1].view(np.int16) + step).view(np.float16)[0]
                e = np.abs(eval_h_f16(u, t.astype(np.float64), A, p) - ht).max()
                if e < best:
                    best, c16, improved = e, t, True
        if not improved:
            break
    return c16.astype(np.float64), best


def bits(x):
    return int(np.asarray([x], dtype=np.float16).view(np.uint16)[0])


def report(A, p, c, u, ht, label):
    e = np.abs(eval_h_f16(u, c, A, p) - ht)
    print(f"{label}: p={p} A={A} m={len(c) - 1}: max |dh| {e.max():.2e} rms {np.sqrt((e**2).mean()):.2e}")
    print(f"  -1/A = {-1.0 / A:.8g} -> 0x{bits(-1.0 / A):04X}")
    for k, v in enumerate(c):
        print(f"  c{k} = {v:.8g} -> +0x{bits(v):04X}  (negated 0x{bits(-v):04X})")


def main():
    u = all_f16_in(-12.0, 12.0)
    ht = h_true(u)
    e0 = np.abs(mufu_path_f16(u) - ht)
    print(f"{u.size} fp16 values of u in [-12, 12]")
    print(f"reference point - exact tanh rounded to fp16, h = fma(u, t, u) in fp16: max {e0.max():.2e} rms {np.sqrt((e0**2).mean()):.2e}")
    if "--naive" in sys.argv:
        for n, A in itertools.product((4, 5, 6, 7), (3.5, 4.0, 4.5)):
            x = np.linspace(0.0, A, 4001)
            c = lawson(np.stack([x**k for k in range(n + 1)], axis=1), s_true(x))
            a = np.minimum(np.abs(f16(u)), f16(A))
            c16 = f16(c)
            pv = np.full_like(a, c16[-1])
            for k in range(n - 1, -1, -1):
                pv = fma16(pv, a, c16[k])
            h = fma16(np.maximum(f16(u), 0.0), 2.0, -pv)
            xs = np.linspace(0, 12, 24001)
            f64 = np.abs(np.polyval(c[::-1], np.minimum(xs, A)) - s_true(xs)).max()
            print(f"naive deg {n} clamp {A}: float64 fit {f64:.2e}, fp16 Horner max |dh| {np.abs(h - ht).max():.2e}")
        return
    xs = np.linspace(0, 12, 24001)
    print(f"{'p':>2} {'A':>4} {'m':>2} {'instr':>5} {'fit64':>9} {'f16 max':>9} {'f16 rms':>9}")
    for p, A, m in itertools.product((2, 4), (4.5, 5.0, 5.5, 6.0, 6.5), (1, 2, 3, 4)):
        c = fit_q(A, p, m)
        w = np.maximum(1 - xs / A, 0) ** p
        fit64 = np.abs(xs * w * np.polyval(c[::-1], w) - s_true(xs)).max()
        e = np.abs(eval_h_f16(u, c, A, p) - ht)
        instr = 6 + int(round(np.log2(p))) + m  # cvt, abs, fma.relu, squarings, a*w, m fma, u+|u|, final fma
        print(f"{p:>2} {A:>4} {m:>2} {instr:>5} {fit64:>9.2e} {e.max():>9.2e} {np.sqrt((e**2).mean()):>9.2e}")
    print()
    for label, (A, p, m) in (("MAIN", (6.0, 4, 3)), ("LITE", (5.5, 4, 1))):
        c, _ = refine(fit_q(A, p, m), A, p, u, ht)
        report(A, p, c, u, ht, label)
        s = eval_s_f16(u, c, A, p)
        es = np.abs(s - s_true(np.abs(u)))
        print(f"  s alone (E3 form, fp16 result): max |ds| {es.max():.2e} rms {np.sqrt((es**2).mean()):.2e}")


if __name__ == "__main__":
    main()
