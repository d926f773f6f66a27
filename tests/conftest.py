import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


# tolerances (north_star): fp32 tier 1e-5 rel-L2, tensor-core (TF32) tier 2e-3 rel-L2
TOL_FP32 = 1e-5
TOL_TF32 = 2e-3
