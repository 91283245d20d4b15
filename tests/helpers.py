"""Test helpers: the product-side workload builders (rails_b200/workloads.py), re-exported for the tests."""
from rails_b200.workloads import (  # noqa: F401
    CFG_8x4x64, CFG_8x4x128, CFG_8x8x32, CFG_16x16x64, MoLConfig, build_module, factory_kwargs, synthetic_inputs,
)
