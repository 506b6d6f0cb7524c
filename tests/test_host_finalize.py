"""CPU tests of the product's host-side pieces: the C ABI loads and exports every declared symbol, the finalize epilogue
(reference :318-404) reproduces the oracle from sufficient statistics, and error reporting works without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gsa_loader
import sobol_oracle as so
import stat_records as sr
from conftest import ROOT, parity_err

g = gsa_loader.load()


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gsa_cuda.h")).read()
    declared = set(re.findall(r"GSA_API\s+int\s+(gsa_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    L = C.CDLL(g.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"libgsa_cuda.so does not export {name}"
    from importlib import import_module
    assert declared == set(import_module("gsa_b200._lib").SIGNATURES)
    assert g.lib().gsa_version() == 100


def test_error_reporting_and_models():
    m = g.DeviceModel("sobol_g", np.ones(3))
    assert m.nout(3) == 1
    with pytest.raises(g.GsaError, match="takes d parameters"):
        m.nout(4)
    with pytest.raises(g.GsaError, match="unknown model"):
        g.DeviceModel("nope")
    assert g.DeviceModel("linear", np.ones((5, 3))).nout(3) == 5
    assert g.DeviceModel("linear_gaussian", np.ones((5, 3))).nout(3) == 5
    with pytest.raises(ValueError, match="Ei_estimator"):
        g.sobol_finalize(g.Sobol(), np.zeros((1, 1, 14)), 3, 10, Ei_estimator="Nope")


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
@pytest.mark.parametrize("nboot,second", [(1, False), (1, True), (5, True)])
def test_finalize_matches_oracle(est, nboot, second):
    lb, ub = -np.pi * np.ones(4), np.pi * np.ones(4)
    A, B = so.generate_design_matrices(1003, lb, ub)
    order = (0, 1, 2) if second else (0, 1)
    ref = so.gsa_sobol(so.ishi_linear_batch, A, B, order=order, nboot=nboot, Ei_estimator=est)
    P = sr.shard_records(so.ishi_linear_batch, A, B, nboot, second, est, 0, A.shape[1])
    res = g.sobol_finalize(g.Sobol(order=order, nboot=nboot), P, 4, A.shape[1], Ei_estimator=est)
    for f in ("S1", "ST", "S2", "S1_Conf_Int", "ST_Conf_Int", "S2_Conf_Int"):
        a, b = getattr(res, f), getattr(ref, f)
        assert (a is None) == (b is None), f
        if a is not None:
            assert parity_err(a, b) < 1e-12, (f, est)      # |delta| <= 1e-12 * max(1, |ref|)


def test_finalize_scalar_shapes_and_additivity():
    a = [0.0, 1.0, 4.5, 9.0, 99.0]
    A, B = so.generate_design_matrices(2048, np.zeros(5), np.ones(5))
    f = lambda X: so.sobol_g_batch(X, a)
    ref = so.gsa_sobol(f, A, B, order=(0, 1, 2), nboot=3)
    # three shards that cut through replicates: the records are additive
    N = A.shape[1]
    total = None
    for rank in range(3):
        c0, nc = g.shard_columns(N, rank, 3)
        P = sr.shard_records(f, A, B, 3, True, "Jansen1999", c0, nc)
        total = P if total is None else total + P
    res = g.sobol_finalize(g.Sobol(order=(0, 1, 2), nboot=3), total, 5, N)
    assert res.S1.shape == (5,) and res.S2.shape == (5, 5) and res.S2_Conf_Int.shape == (5, 5)
    for fld in ("S1", "ST", "S2", "S1_Conf_Int", "ST_Conf_Int", "S2_Conf_Int"):
        assert parity_err(getattr(res, fld), getattr(ref, fld)) < 1e-12, fld
    assert g.shard_columns(10, 0, 3) == (0, 4) and g.shard_columns(10, 2, 3) == (7, 3)


def test_variance_from_dd_sums_is_robust():
    """mean^2 >> var: the one-pass double-double variance must still match the reference's two-pass var."""
    rng = np.random.default_rng(0)
    n, d = 500, 2
    A = rng.random((d, n)); B = rng.random((d, n))
    f = lambda X: 1.0e6 + 1.0e-3 * (X[0] + 2 * X[1])
    ref = so.gsa_sobol(f, A, B)
    P = sr.shard_records(f, A, B, 1, False, "Jansen1999", 0, n)
    res = g.sobol_finalize(g.Sobol(), P, d, n)
    assert parity_err(res.S1, ref.S1) < 1e-9 and parity_err(res.ST, ref.ST) < 1e-12
