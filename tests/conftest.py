import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))       # tests/helpers (harness), tests/helpers_r2.py

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "tdnet_reference.npz"))


@pytest.fixture(scope="session")
def schemas():
    with open(os.path.join(GOLDEN_DIR, "state_dict_schema.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test session starts from a freshly built libnsdp_b200.so and oracle (no-op when up to date)."""
    from nsdp_b200.build import build
    build()
    from oracle.build import build as build_oracle
    build_oracle()


@pytest.fixture(scope="session")
def golden_r2():
    """Round-2 fixtures from the live reference (tests/golden/make_golden_r2.py): per-block encoder activations and the
    staged FlowArbitrary training step."""
    return np.load(os.path.join(GOLDEN_DIR, "tdnet_reference_r2.npz"))
