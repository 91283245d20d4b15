"""CPU checks of the MUFU-free half2 silu(2u) that the coarse kernel can be built with (MOL_E2_H2_MASK):
the fp16 constants in the CUDA source are the ones the numerics model uses, and the emulated instruction sequence is as
accurate as the shipped MUFU.TANH.F16 path (tools/fit_silu_h2.py is the design script)."""
import os
import re

import numpy as np
import torch

from tests import sim_coarse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "rails_b200", "csrc", "mol_coarse_sm100.cu")


def _all_f16(lo, hi):
    v = np.arange(0, 1 << 16, dtype=np.uint16).view(np.float16).astype(np.float64)
    v = v[np.isfinite(v)]
    return torch.tensor(v[(v >= lo) & (v <= hi)], dtype=torch.float32)


def _f16_of_bits(bits: int) -> float:
    return float(np.array([bits], dtype=np.uint16).view(np.float16)[0])


def _source_constants():
    """{lite: (inv_a, (nc0, nc1, ...))} parsed from the kH2* definitions of the kernel source (the degree-1 "lite" form
    was measured in round 1 and removed from the kernel; it only lives on in the numerics model)."""
    block = open(SRC).read()
    out = {}
    for lite in (False,):
        inv_a = _f16_of_bits(int(re.search(r"kH2NegInvA = h2x2\((0x[0-9A-Fa-f]+)\)", block).group(1), 16))
        nc = [
            _f16_of_bits(int(m, 16))
            for m in re.findall(r"kH2NC\d = h2x2\((0x[0-9A-Fa-f]+)\)", block)
        ]
        out[lite] = (inv_a, tuple(nc))
    return out


def test_kernel_constants_equal_the_model_constants():
    src = _source_constants()
    for lite in (False,):
        inv_a, nc = src[lite]
        k = sim_coarse._H2[lite]
        assert np.float16(k["inv_a"]) == np.float16(inv_a)
        assert len(nc) == len(k["nc"])
        for a, b in zip(nc, k["nc"]):
            assert np.float16(a) == np.float16(b), (lite, a, b)


def test_half2_silu_is_as_accurate_as_the_mufu_path():
    u = _all_f16(-12.0, 12.0)
    ref = (u.double() + u.double() * torch.tanh(u.double()))
    # shipped path with an exact tanh: t = fp16(tanh u), h = fp16(fma(u, t, u))
    t = torch.tanh(u).to(torch.float16).float()
    shipped = (u + u * t).to(torch.float16).double()
    e_shipped = (shipped - ref).abs()
    e_h2 = (sim_coarse.silu2_h2(u).double() - ref).abs()
    assert e_h2.max().item() <= 3.0e-3
    assert e_h2.max().item() <= 1.05 * e_shipped.max().item()
    assert e_h2.pow(2).mean().sqrt().item() <= 4.0e-4
    # the degree-1 "lite" form is 8x coarser but bounded
    e_lite = (sim_coarse.silu2_h2(u, lite=True).double() - ref).abs()
    assert e_lite.max().item() <= 8.0e-3


def test_half2_silu_limits():
    # exact outside the bump: u + |u| for |u| >= A (A = 6), including values far outside the fitted range
    u = torch.tensor([-1000.0, -60.0, -6.0, 6.0, 7.5, 60.0, 1000.0])
    h = sim_coarse.silu2_h2(u)
    assert torch.equal(h, torch.tensor([0.0, 0.0, 0.0, 12.0, 15.0, 120.0, 2000.0]))
    assert sim_coarse.silu2_h2(torch.zeros(1)).item() == 0.0
    # E3 form: bump in half2, large part in fp32
    uu = torch.linspace(-12, 12, 20001)
    ref = uu.double() + uu.double() * torch.tanh(uu.double())
    assert (sim_coarse.silu2_h2_f32(uu).double() - ref).abs().max().item() <= 1.5e-3


def test_model_knobs_change_only_the_selected_units():
    from tests.helpers import CFG_8x8x32, build_module, synthetic_inputs

    cfg = CFG_8x8x32
    mol, _ = build_module(cfg, None, "cpu", seed=3)
    sd = {k: v.detach() for k, v in mol.state_dict().items()}
    items, _, q, _ = synthetic_inputs(cfg, 512, 3, 3, "cpu")
    base = sim_coarse.coarse_scores(cfg, sd, q, items)
    assert torch.equal(base, sim_coarse.coarse_scores(cfg, sd, q, items, e2_h2_mask=0, e3_h2_of4=0))
    for kw in (dict(e2_h2_mask=0x0E), dict(e2_h2_mask=0xFF), dict(e2_h2_mask=0xFF, e3_h2_of4=4)):
        v = sim_coarse.coarse_scores(cfg, sd, q, items, **kw)
        d = (v - base).abs().max().item()
        assert 0.0 < d < 0.03, (kw, d)
