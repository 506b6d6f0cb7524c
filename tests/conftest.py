import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_literals.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ishigami_design():
    """The design of test/sobol_method.jl:26-30 (n=600000, d=4) from the C oracle."""
    import gsa_ref
    n = 600000
    A, B = gsa_ref.sobol_design(n, -np.pi * np.ones(4), np.pi * np.ones(4))
    return A, B


def parity_err(x, ref):
    """|delta| <= 1e-12 * max(1, |ref|)  ->  returns max of |delta| / max(1,|ref|)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if x.size == 0:
        return 0.0
    return float(np.max(np.abs(x - ref) / np.maximum(1.0, np.abs(ref))))
