import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # MOL_MODE_AUTO serves searches below 2^15 (query, item) pairs with the fp32 kernel (faster there); the reference
    # fixtures are that small, and the suite wants MODE_AUTO to keep exercising the tensor-core path on them
    # (test_small_searches_take_the_exact_kernel_in_auto_mode covers the switch itself)
    os.environ.setdefault("MOL_B200_TENSOR_MIN_PAIRS", "0")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
