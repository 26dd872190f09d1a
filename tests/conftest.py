import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cab():
    return importlib.import_module("chess-ai_b200")


@pytest.fixture(scope="session")
def golden():
    g = dict(np.load(os.path.join(GOLDEN, "golden.npz")))
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        g["json"] = json.load(f)
    g["latest_weights"] = np.fromfile(os.path.join(GOLDEN, "latest_weights.f64"), dtype="<f8")
    return g


@pytest.fixture(scope="session")
def port():
    from oracle_bind import PortOracle, build
    build(ref=False)
    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own code, when oracle/_ref was built (it travels to the GPU box as a built file)."""
    from oracle_bind import RefOracle
    if not RefOracle.available():
        pytest.skip("oracle/_ref/libchessai_ref.so not built (needs /root/reference)")
    return RefOracle()


DEFAULT_DIMS = [64, 144, 225, 100, 1]


def rel_err(got, want, floor=1e-12):
    """Relative error with an absolute floor: |got-want| / max(|want|, floor)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), floor)
