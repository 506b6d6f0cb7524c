"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs.

Tolerance (SURVEY 8 / BASELINE.md): indices |delta| <= 1e-12 * max(1, |ref|); Sobol points bit-exact."""
import os

import numpy as np
import pytest

import gsa_loader
import gsa_ref
import sobol_oracle as so
from conftest import parity_err

pytestmark = pytest.mark.gpu
g = gsa_loader.load()

TOL = 1e-12
FIELDS = ("S1", "ST", "S2", "S1_Conf_Int", "ST_Conf_Int", "S2_Conf_Int")


def assert_parity(res, ref, tol=TOL, what=""):
    for f in FIELDS:
        a, b = getattr(res, f), getattr(ref, f)
        assert (a is None) == (b is None), (f, what)
        if a is not None:
            assert np.shape(a) == np.shape(b), (f, what, np.shape(a), np.shape(b))
            e = parity_err(a, b)
            assert e <= tol, (f, what, e)


def ulp_sensitivity(f, A, B, trials=8, **kw):
    """How far +-1 ulp on every model value moves each index: max over `trials` random sign patterns of the distance between the
    extended-precision evaluation (oracle/sobol_oracle.py gsa_sobol_exact) of the perturbed and of the unperturbed values."""
    nboot, second = kw.get("nboot", 1), 2 in kw.get("order", (0, 1))
    d, N = A.shape
    n = N // nboot
    y = np.asarray(f(so.all_points(np.asarray(A), np.asarray(B), nboot, second)), dtype=np.float64)
    ex = so.gsa_sobol_exact(None, None, None, all_y=y, d=d, n=n, **kw)
    rng = np.random.default_rng(0)
    sens = {}
    for _ in range(trials):
        yp = np.where(rng.integers(0, 2, y.shape) == 1, np.nextafter(y, np.inf), np.nextafter(y, -np.inf))
        pe = so.gsa_sobol_exact(None, None, None, all_y=yp, d=d, n=n, **kw)
        for fld in FIELDS:
            if getattr(ex, fld) is not None:
                sens[fld] = max(sens.get(fld, 0.0), parity_err(getattr(pe, fld), getattr(ex, fld)))
    return ex, sens


def assert_parity_adjudicated(res, ref, f, A, B, what="", **kw):
    """1e-12 against the oracle, or -- where the model values themselves differ by an ulp between CUDA's and glibc's libm --
    within 4x what +-1 ulp on the model values does to the exact evaluation (and the oracle must be inside the same bound)."""
    ex, sens = ulp_sensitivity(f, A, B, **kw)
    for fld in FIELDS:
        a, b = getattr(res, fld), getattr(ref, fld)
        assert (a is None) == (b is None), (fld, what)
        if a is None:
            continue
        bound = max(TOL, 4.0 * sens[fld])
        e_gpu, e_orc = parity_err(a, getattr(ex, fld)), parity_err(b, getattr(ex, fld))
        assert e_gpu <= bound and e_orc <= bound, (fld, what, e_gpu, e_orc, bound)


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("n,d,num_mats", [(4, 4, 1), (8, 2, 1), (5, 3, 2), (1000, 3, 2), (600000, 4, 2), (4097, 100, 2),
                                           (300, 7, 40), (1, 2, 2), (65536, 1, 2)])
def test_design_bit_exact(n, d, num_mats):
    rng = np.random.default_rng(n + d)
    lb = -rng.random(d) * 3
    ub = lb + rng.random(d) * 5 + 0.1
    got = g.generate_design_matrices(n, lb, ub, num_mats)
    want = gsa_ref.sobol_design(n, lb, ub, num_mats)
    for m in range(num_mats):
        assert got[m].shape == (d, n)
        assert np.array_equal(got[m], want[m]), (m, np.abs(got[m] - want[m]).max())


def test_design_golden_points(golden):
    k = golden["sobol_points_kat"]
    assert np.array_equal(g.generate_design_matrices(4, np.zeros(4), np.ones(4), 1)[0], np.array(k["n4_dim4"]))
    assert np.array_equal(g.generate_design_matrices(8, np.zeros(2), np.ones(2), 1)[0], np.array(k["n8_dim2"]))
    A, B = g.generate_design_matrices(600000, -np.pi * np.ones(4), np.pi * np.ones(4))
    assert np.array_equal(A[:, 0], np.array(k["n600000_A_col1"])) and np.array_equal(B[:, 0], np.array(k["n600000_B_col1"]))


def test_design_device_shard_and_far_indices():
    import torch
    n, d = (1 << 26), 100          # BASELINE config 5 geometry; only a window of columns is generated
    lb, ub = np.zeros(d), np.ones(d)
    col0, ncols = (1 << 26) - 5000, 5000
    A, B = g.generate_design_matrices(n, lb, ub, 2, device=True, col0=col0, ncols=ncols)
    assert A.shape == (d, ncols) and A.is_cuda
    wA, wB = gsa_ref.sobol_design(n, lb, ub, 2, col0=col0, ncols=ncols)
    assert np.array_equal(A.cpu().numpy(), wA) and np.array_equal(B.cpu().numpy(), wB)
    # against SciPy's independent generator at index 2^26 + col0
    from scipy.stats import qmc
    sp = qmc.Sobol(2 * d, scramble=False, bits=32)
    sp.fast_forward((1 << 26) + col0)
    pts = sp.random(16).T
    assert np.array_equal(A.cpu().numpy()[:, :16], pts[:d]) and np.array_equal(B.cpu().numpy()[:, :16], pts[d:])


def test_design_at_the_far_end_of_the_table():
    """K1 on the last dimensions of the 21201-dimensional table (a 106-matrix design of d = 200), bit-exact against SciPy's own
    generator (dimensions above 280 were never cross-checked against an independent generator)."""
    from scipy.stats import qmc
    n, d, mats = 64, 200, 106
    got = g.generate_design_matrices(n, np.zeros(d), np.ones(d), mats)
    sp = qmc.Sobol(d * mats, scramble=False, bits=32)
    sp.fast_forward(so.sobol_first_index(n))
    want = sp.random(n).T
    for m in (0, 52, 104, 105):
        assert np.array_equal(got[m], want[m * d:(m + 1) * d]), m
    with pytest.raises(g.GsaError, match="21201"):
        g.generate_design_matrices(n, np.zeros(d), np.ones(d), 107)


# ------------------------------------------------------------------------------------------------ K2..K4
def test_reference_goldens_on_gpu(golden, ishigami_design):
    """The reference's own literal vectors (test/sobol_method.jl:41-45, :55-71) through the CUDA path."""
    A, B = ishigami_design
    ishi = g.DeviceModel("ishigami", [7.0, 0.1])
    r = g.gsa(ishi, g.Sobol(order=[0, 1, 2]), A, B, batch=True)
    assert np.abs(r.S2 - np.array(golden["ishigami_S2_order012"]["value"])).max() < 1e-13
    assert np.abs(r.S1 - np.array(golden["ishigami_S1_atol1e-4"]["value"])).max() < 1e-4
    assert np.abs(r.ST - np.array(golden["ishigami_ST_atol1e-4"]["value"])).max() < 1e-4
    assert r.S1_Conf_Int is None and r.S2_Conf_Int is None
    r = g.gsa(ishi, g.Sobol(order=[0, 1, 2], nboot=20), A, B, batch=True)
    assert np.abs(r.S1_Conf_Int - np.array(golden["ishigami_nboot20_S1_Conf_Int"]["value"])).max() < 1e-13
    assert np.abs(r.ST_Conf_Int - np.array(golden["ishigami_nboot20_ST_Conf_Int"]["value"])).max() < 1e-13
    assert np.abs(r.S2_Conf_Int - np.array(golden["ishigami_nboot20_S2_Conf_Int"]["value"])).max() < 1e-13
    lin = g.DeviceModel("linear", np.array([[7.0, 0.1, 0.0, 0.0]]))
    r = g.gsa(lin, g.Sobol(), A, B, batch=True)
    want = np.array(golden["linear_S1_ST_atol1e-4"]["value"])
    assert np.abs(r.S1 - want).max() < 1e-4 and np.abs(r.ST - want).max() < 1e-4


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
@pytest.mark.parametrize("nboot", [1, 20])
def test_ishigami_all_estimators(ishigami_design, est, nboot):
    A, B = ishigami_design
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=nboot, Ei_estimator=est)
    res = g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True, Ei_estimator=est)
    assert_parity(res, ref, what=(est, nboot))


def test_ci_properties_issue_181(ishigami_design):
    """test/sobol_method.jl:74-105."""
    A, B = ishigami_design
    ishi = g.DeviceModel("ishigami", [7.0, 0.1])
    res = {c: g.gsa(ishi, g.Sobol(order=[0, 1, 2], nboot=20, conf_level=c), A, B, batch=True) for c in (0.8, 0.95, 0.99)}
    for f in ("S1", "ST", "S2"):
        assert np.array_equal(getattr(res[0.8], f), getattr(res[0.95], f)) and np.array_equal(getattr(res[0.95], f), getattr(res[0.99], f))
    z = {c: so.norm_quantile((1 + c) / 2) for c in res}
    for f in ("S1_Conf_Int", "ST_Conf_Int", "S2_Conf_Int"):
        a, b, c = (getattr(res[k], f) for k in (0.8, 0.95, 0.99))
        assert np.all(a <= b) and np.all(b <= c) and not np.array_equal(a, b)
        assert np.allclose(a / z[0.8], b / z[0.95], rtol=1e-13, atol=0) and np.allclose(a / z[0.8], c / z[0.99], rtol=1e-13, atol=0)


@pytest.mark.parametrize("d,N,nboot,second", [(8, 1 << 16, 1, True), (8, 50000, 7, True), (3, 1000, 1, False), (13, 4099, 3, False),
                                               (40, 3001, 2, True), (100, 2000, 1, False), (128, 700, 1, False), (1, 999, 1, True),
                                               (2, 5003, 1, True), (5, 30011, 3, True), (20, 20011, 2, True), (32, 7001, 1, True), (33, 3001, 1, True)])
def test_sobol_g(d, N, nboot, second):
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    order = (0, 1, 2) if second else (0, 1)
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=order, nboot=nboot)
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=order, nboot=nboot), A, B, batch=True)
    assert_parity(res, ref, what=(d, N, nboot, second))


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
def test_multi_output(ishigami_design, est):
    """res.S1[out, :] layout (test/sobol_method.jl:160-175); S2 is d x d x nout."""
    A, B = ishigami_design
    A, B = A[:, :100001], B[:, :100001]
    ref = gsa_ref.gsa_sobol("ishi_linear", [0.0], A, B, order=(0, 1, 2), nboot=3, Ei_estimator=est)
    res = g.gsa(g.DeviceModel("ishi_linear"), g.Sobol(order=[0, 1, 2], nboot=3), A, B, batch=True, Ei_estimator=est)
    assert res.S1.shape == (2, 4) and res.S2.shape == (4, 4, 2) and res.S2_Conf_Int.shape == (4, 4, 2)
    assert_parity(res, ref, what=est)
    W = np.arange(1, 16, dtype=np.float64).reshape(3, 5) / 7.0
    A, B = gsa_ref.sobol_design(20000, -np.ones(5), np.ones(5) * 2)
    ref = gsa_ref.gsa_sobol("linear", W, A, B, order=(0, 1, 2), Ei_estimator=est)
    res = g.gsa(g.DeviceModel("linear", W), g.Sobol(order=[0, 1, 2]), A, B, batch=True, Ei_estimator=est)
    assert_parity(res, ref, what=("linear", est))


def test_device_inputs_and_torch(ishigami_design):
    import torch
    A, B = ishigami_design
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=4)
    At, Bt = torch.from_numpy(np.ascontiguousarray(A)).cuda(), torch.from_numpy(np.ascontiguousarray(B)).cuda()   # plain (d, N) row-major tensors
    res = g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=4), At, Bt, batch=True)
    assert_parity(res, ref)
    # pinned host tensors go through the streamed (chunked H2D) path
    Ap, Bp = torch.from_numpy(np.ascontiguousarray(A.T)).pin_memory().t(), torch.from_numpy(np.ascontiguousarray(B.T)).pin_memory().t()
    res = g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=4), Ap, Bp, batch=True)
    assert_parity(res, ref)


def test_edge_cases():
    ishi = g.DeviceModel("ishigami", [7.0, 0.1])
    lb, ub = -np.pi * np.ones(3), np.pi * np.ones(3)
    # N not a multiple of nboot: trailing columns silently dropped (:118)
    A, B = gsa_ref.sobol_design(1048583, lb, ub)
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=1000)
    res = g.gsa(ishi, g.Sobol(order=[0, 1, 2], nboot=1000), A, B, batch=True)
    assert_parity(res, ref)
    # tiny designs: a handful of values decide Var, so a 1-ulp difference between CUDA's and glibc's sin moves the indices by
    # more than 1e-12.  Adjudicated: the bound is what +-1 ulp on the model values does to the
    # extended-precision evaluation of the reference's formulas -- the kernel must stay within it, the oracle does too.
    for N, nboot in ((2, 1), (5, 2), (33, 1)):
        A, B = gsa_ref.sobol_design(N, lb, ub)
        ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=nboot)
        res = g.gsa(ishi, g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True)
        assert_parity_adjudicated(res, ref, so.ishigami_batch, A, B, order=(0, 1, 2), nboot=nboot, what=(N, nboot))
    # nboot > N  =>  n = 0  =>  NaNs, as in the reference (SURVEY 8(b) "Errors")
    A, B = gsa_ref.sobol_design(3, lb, ub)
    res = g.gsa(ishi, g.Sobol(nboot=5), A, B, batch=True)
    assert np.all(np.isnan(res.S1)) and np.all(np.isnan(res.ST))
    with pytest.raises(ValueError):
        g.gsa(ishi, g.Sobol(), A, B[:, :2], batch=True)
    with pytest.raises(g.GsaError, match="ishigami needs d >= 3"):
        g.gsa(ishi, g.Sobol(), A[:2], B[:2], batch=True)


# ------------------------------------------------------------------------------------------------ K3 (estimator on Y)
@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
@pytest.mark.parametrize("nboot,second", [(1, False), (1, True), (6, True)])
def test_analyze_y_scalar(est, nboot, second):
    d, N = 6, 30011
    a = np.array([0.0, 0.5, 3.0, 9.0, 99.0, 99.0])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    order = (0, 1, 2) if second else (0, 1)
    P = gsa_ref.all_points(A, B, nboot, second)
    Y = gsa_ref.model_batch("sobol_g", a, P)
    ref = gsa_ref.analyze_y(Y, d, N // nboot, order=order, nboot=nboot, Ei_estimator=est)
    res = g.gsa_sobol_all_y_analysis(g.Sobol(order=order, nboot=nboot), Y, d, N // nboot, est)
    assert_parity(res, ref, what=(est, nboot, second))


def test_analyze_y_multi_output_and_device():
    import torch
    d, N, nboot = 4, 20001, 3
    A, B = gsa_ref.sobol_design(N, -np.pi * np.ones(d), np.pi * np.ones(d))
    P = gsa_ref.all_points(A, B, nboot, True)
    Y = gsa_ref.model_batch("ishi_linear", [0.0], P)           # (2, cols)
    ref = gsa_ref.analyze_y(Y, d, N // nboot, order=(0, 1, 2), nboot=nboot)
    res = g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=nboot), Y, d, N // nboot)
    assert res.S1.shape == (2, d) and res.S2.shape == (d, d, 2)
    assert_parity(res, ref)
    res = g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=nboot), torch.from_numpy(Y).cuda(), d, N // nboot)
    assert_parity(res, ref)


def test_closure_path_and_all_points():
    """An arbitrary (Python) closure with batch=true: fuse_designs on the GPU + estimator kernels."""
    d, N, nboot = 4, 5003, 2
    A, B = gsa_ref.sobol_design(N, -np.pi * np.ones(d), np.pi * np.ones(d))
    seen = {}

    def f(X):
        seen["X"] = np.array(X)
        return so.ishigami_batch(X)

    res = g.gsa(f, g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True)
    assert np.array_equal(seen["X"], so.all_points(A, B, nboot, True))          # bit-exact copy semantics
    ref = so.gsa_sobol(so.ishigami_batch, A, B, order=(0, 1, 2), nboot=nboot)
    assert_parity(res, ref)
    res = g.gsa(lambda x: so.ishigami_batch(x[:, None])[0], g.Sobol(), A[:, :200], B[:, :200])     # batch=false (:151)
    ref = so.gsa_sobol(so.ishigami_batch, A[:, :200], B[:, :200])
    assert_parity(res, ref)


def test_p_range_entry():
    """gsa(f, Sobol(nboot), p_range; samples) (:407-417): a 2*nboot*d-dimensional design."""
    p_range = [[-np.pi, np.pi]] * 3
    ref = so.gsa_sobol_p_range(so.ishigami_batch, p_range, 4000, order=(0, 1, 2), nboot=5)
    res = g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=5), p_range, samples=4000, batch=True)
    assert_parity(res, ref)
    res = g.gsa(so.ishigami_batch, g.Sobol(order=[0, 1, 2], nboot=5), p_range, samples=4000, batch=True)
    assert_parity(res, ref)


def test_sample_matches_single_matrix_design():
    """QuasiMonteCarlo.sample(n, lb, ub, SobolSample()): K1 with one matrix."""
    lb, ub = np.array([-1.0, 0.0, 2.0, -5.0, 0.5]), np.array([1.0, 3.0, 2.5, 5.0, 0.75])
    X = g.sample(777, lb, ub)
    assert X.shape == (5, 777) and np.array_equal(X, gsa_ref.sobol_design(777, lb, ub, 1)[0])


# ------------------------------------------------------------------------------------------------ K5 (sharding)
def test_sharded_partials_are_additive():
    import torch
    d, N, nboot = 8, 100003, 3
    a = np.array([0.0, 1.0, 4.5, 9.0, 99.0, 99.0, 99.0, 99.0])
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2], nboot=nboot)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2), nboot=nboot)
    total = None
    for rank in range(4):
        c0, nc = g.shard_columns(N, rank, 4)
        # each "rank" generates its own shard on the device (K1 is random-access)
        Al, Bl = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, col0=c0, ncols=nc)
        p = g.sobol_partials(model, method, Al, Bl, N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    res = g.sobol_finalize(method, total, d, N)
    assert_parity(res, ref)


def test_product_path_matches_generic_path(monkeypatch):
    """The product-form fast kernel (prefix/suffix products, reciprocal coefficients) against the generic kernel
    (full model evaluation per swapped coordinate) and the oracle, at BASELINE config 5's dimension."""
    d, N = 100, 30000
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=3)
    fast = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=3), A, B, batch=True)
    monkeypatch.setenv("GSA_FORCE_GENERIC", "1")
    slow = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=3), A, B, batch=True)
    monkeypatch.delenv("GSA_FORCE_GENERIC")
    assert_parity(fast, ref)
    assert_parity(slow, ref)
    for T in ("16", "32"):
        monkeypatch.setenv("GSA_PRODUCT_T", T)
        alt = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=3), A, B, batch=True)
        assert_parity(alt, ref, what=T)


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
def test_survey_estimator_known_answers_on_gpu(golden, ishigami_design, est):
    """SURVEY Appendix B known answers (the survey's independent probe) for all four Ei_estimators on the reference's test design."""
    A, B = ishigami_design
    ka = golden["survey_appendix_b_estimators"]
    for name, model in (("ishigami", g.DeviceModel("ishigami", [7.0, 0.1])), ("linear", g.DeviceModel("linear", np.array([[7.0, 0.1, 0.0, 0.0]])))):
        r = g.gsa(model, g.Sobol(), A, B, batch=True, Ei_estimator=est)
        assert np.abs(np.ravel(r.S1) - np.array(ka[name]["S1"])).max() < 1e-12, (name, est)
        assert np.abs(np.ravel(r.ST) - np.array(ka[name]["ST"][est])).max() < 1e-12, (name, est)


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
@pytest.mark.parametrize("d,N,nboot", [(100, 8000, 2), (5, 30011, 1), (40, 12345, 3)])
def test_product_kernel_all_estimators(est, d, N, nboot):
    """Config-5-type models with every Ei_estimator: Jansen1999, Sobol2007 and Homma1996 share the product-form kernel (one
    accumulator per coordinate each); Janon2014 takes the generic kernel.  Host and device inputs."""
    import torch
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=nboot, Ei_estimator=est)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(nboot=nboot)
    assert_parity(g.gsa(model, method, A, B, batch=True, Ei_estimator=est), ref, what=(est, d, "host"))
    At, Bt = torch.from_numpy(np.ascontiguousarray(A)).cuda(), torch.from_numpy(np.ascontiguousarray(B)).cuda()
    assert_parity(g.gsa(model, method, At, Bt, batch=True, Ei_estimator=est), ref, what=(est, d, "device"))


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
@pytest.mark.parametrize("d,second,nboot", [(3, False, 1), (3, True, 1000), (4, False, 7)])
def test_ishigami_small_kernels(est, d, second, nboot):
    """BASELINE configs 1 and 2 shapes (d = 3; README.md:46-56) through the thread-per-sample register kernel."""
    N = 600000 if nboot == 1 else (1 << 20)
    A, B = gsa_ref.sobol_design(N, -np.pi * np.ones(d), np.pi * np.ones(d))
    order = (0, 1, 2) if second else (0, 1)
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=order, nboot=nboot, Ei_estimator=est)
    res = g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=order, nboot=nboot), A, B, batch=True, Ei_estimator=est)
    assert_parity(res, ref, what=(est, d, second, nboot))


@pytest.mark.parametrize("second,d", [(False, 14), (True, 9), (True, 14), (True, 20), (True, 24), (True, 26), (False, 40)])
def test_analyze_y_large_d(second, d):
    """Beyond the register kernel's range (d > 24, or d > 8 with second order): the shared-memory estimator kernel (estimator_y.cuh)."""
    N, nboot = 9001, 2
    a = np.linspace(0.0, 20.0, d)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    order = (0, 1, 2) if second else (0, 1)
    Y = gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(A, B, nboot, second))
    for est in so.EI_ESTIMATORS:
        ref = gsa_ref.analyze_y(Y, d, N // nboot, order=order, nboot=nboot, Ei_estimator=est)
        res = g.gsa_sobol_all_y_analysis(g.Sobol(order=order, nboot=nboot), Y, d, N // nboot, est)
        assert_parity(res, ref, what=(est, second))


@pytest.mark.parametrize("d,N,nboot", [(2, 5003, 1), (3, 4099, 2), (7, 30011, 1), (8, 1 << 16, 3), (9, 9001, 2), (16, 20000, 1), (20, 33333, 4),
                                       (25, 4100, 1), (32, 6007, 3), (8, 5, 1), (12, 700, 50), (4, 10007, 2), (6, 12345, 1), (8, 33, 1)])
def test_pairs_mma_kernel(d, N, nboot, monkeypatch):
    """Second order for product-form models, d <= 32: the pair sums run as FP64 mma over the samples (fused_pairs_mma.cuh).
    Every block count, ragged tiles (N not a multiple of 4), many short replicates; fused source and Y source."""
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2), nboot=nboot)
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True)
    assert_parity(res, ref, what=(d, N, nboot, "fused"))
    monkeypatch.setenv("GSA_PAIRS_MMA", "1")                       # also for d <= 8, where the register kernel is the default
    Y = np.stack([gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(A, B, nboot, True)),
                  gsa_ref.model_batch("sobol_g", a[::-1].copy(), gsa_ref.all_points(A, B, nboot, True))])
    refy = gsa_ref.analyze_y(Y, d, N // nboot, order=(0, 1, 2), nboot=nboot)
    resy = g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=nboot), Y, d, N // nboot)
    assert_parity(resy, refy, what=(d, N, nboot, "Y, 2 outputs"))


@pytest.mark.parametrize("est", ["Sobol2007", "Homma1996", "Janon2014"])
@pytest.mark.parametrize("d,N,nboot", [(8, 20011, 2), (20, 9001, 1), (6, 4099, 3)])
def test_pairs_mma_other_estimators(est, d, N, nboot):
    """Second order with the other Ei_estimators: Sobol2007 and Homma1996 share the mma pair kernels (fused sources and Y source),
    Janon2014 falls back to the generic / shared-memory kernels; all against the oracle."""
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2), nboot=nboot, Ei_estimator=est)
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True, Ei_estimator=est)
    assert_parity(res, ref, what=(est, d, "fused"))
    Y = gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(A, B, nboot, True))
    resy = g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=nboot), Y, d, N // nboot, est)
    assert_parity(resy, ref, what=(est, d, "Y"))


def test_pairs_mma_sources_agree(monkeypatch):
    """d = 8 through all three fused second-order kernels (register kernel, lane-per-coordinate mma source, thread-per-sample
    tile source) and through a design that is NOT 16-byte aligned (the tile source needs 16-byte rows and must step aside)."""
    import torch
    d, N, nboot = 8, 50001, 2
    a = np.array([0.0, 1.0, 4.5, 9.0, 99.0, 99.0, 99.0, 99.0])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2), nboot=nboot)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2], nboot=nboot)
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("GSA_PAIRS_MMA", mode)
        assert_parity(g.gsa(model, method, A, B, batch=True), ref, what=("mode", mode))
    monkeypatch.delenv("GSA_PAIRS_MMA")
    bufA = torch.empty(d * N + 1, dtype=torch.float64, device="cuda")
    bufB = torch.empty(d * N + 1, dtype=torch.float64, device="cuda")
    At, Bt = bufA[1:].view(N, d).t(), bufB[1:].view(N, d).t()          # (d, N) views in Julia layout, 8 bytes off a 16-byte boundary
    At.copy_(torch.from_numpy(np.ascontiguousarray(A)))
    Bt.copy_(torch.from_numpy(np.ascontiguousarray(B)))
    assert At.data_ptr() % 16 == 8
    assert_parity(g.gsa(model, method, At, Bt, batch=True), ref, what="unaligned")


def test_pairs_mma_large_analytic_and_additive():
    """Second order at a size the oracle cannot materialise (d = 20, N = 2^20: all_points would be 7 GB): analytic Sobol-G indices
    (S2_ij = V_i V_j / V) and shard additivity of the mma kernel's records; the same on the model's Y is not
    possible at this size, so the Y source is checked against the fused source on a slice."""
    import torch
    d, N = 20, 1 << 20
    a = np.array([0.0, 0.5, 1.0, 2.0, 4.5, 9.0] + [50.0] * 14)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2])
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
    res = g.gsa(model, method, A, B, batch=True)
    Vi = 1.0 / (3.0 * (1.0 + a) ** 2)
    V = np.prod(1.0 + Vi) - 1.0
    assert np.abs(res.S1 - Vi / V).max() < 2e-3 and np.abs(res.ST - Vi * np.prod(1.0 + Vi) / (1.0 + Vi) / V).max() < 2e-3
    S2 = np.triu(np.outer(Vi, Vi) / V, 1)
    assert np.abs(res.S2 - S2).max() < 3e-3 and np.all(np.tril(res.S2) == 0.0)
    total = None
    for rank in range(3):
        c0, nc = g.shard_columns(N, rank, 3)
        p = g.sobol_partials(model, method, A[:, c0:c0 + nc], B[:, c0:c0 + nc], N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    assert_parity(g.sobol_finalize(method, total, d, N), res)
    n = 1 << 14                                                   # Y source against the fused source and the oracle on a slice
    An, Bn = A[:, :n].cpu().numpy(), B[:, :n].cpu().numpy()
    Y = gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(An, Bn, 1, True))
    ref = gsa_ref.analyze_y(Y, d, n, order=(0, 1, 2))
    assert_parity(g.gsa_sobol_all_y_analysis(method, Y, d, n), ref, what="Y source")
    assert_parity(g.gsa(model, method, An, Bn, batch=True), ref, what="fused source")


def test_pairs_mma_mean_dominates():
    """Outputs with mean >> spread (all a_k = 99: y = 1 +- 1e-2): the shifted products keep the pair sums exact enough."""
    d, N = 8, 1 << 17
    a = np.full(d, 99.0)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2))
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2]), A, B, batch=True)
    assert_parity(res, ref)


def test_small_kernels_match_generic(monkeypatch):
    d, N = 8, 1 << 17
    a = np.array([0.0, 1.0, 4.5, 9.0, 99.0, 99.0, 99.0, 99.0])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2), nboot=2)
    Y = gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(A, B, 2, True))
    for force in ("0", "1"):
        monkeypatch.setenv("GSA_FORCE_GENERIC", force)
        res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2], nboot=2), A, B, batch=True)
        assert_parity(res, ref, what=("fused", force))
        res = g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=2), Y, d, N // 2)
        assert_parity(res, ref, what=("y", force))


@pytest.mark.parametrize("d,nout,N,nboot", [(20, 256, 1 << 15, 1), (5, 3, 20000, 4), (31, 7, 5000, 1), (3, 1, 4097, 2), (12, 40, 30011, 3),
                                            (32, 4, 9001, 2), (1, 2, 3001, 1), (4, 1, 1000, 5), (17, 3, 1 << 16, 1), (8, 300, 777, 37), (20, 2, 3, 1)])
def test_linear_gram_kernel(d, nout, N, nboot):
    """BASELINE config 4 shape (d=20, 256 outputs; smaller N) and odd shapes through the input-moment (Gram, FP64 mma) kernel:
    every block count 1..8 of the stacked [a; b] vector, tiles that are not a multiple of 4 samples, many short replicates."""
    rng = np.random.default_rng(d * 1000 + nout)
    W = rng.normal(size=(nout, d)) * (1.0 + np.arange(d))[None, :] / d
    lb, ub = -rng.random(d) - 0.5, rng.random(d) + 2.0            # a box with non-zero mean: exercises the shift
    A, B = gsa_ref.sobol_design(N, lb, ub)
    ref = gsa_ref.gsa_sobol("linear", W, A, B, nboot=nboot)
    res = g.gsa(g.DeviceModel("linear", W), g.Sobol(nboot=nboot), A, B, batch=True)
    if nout == 1:
        ref.S1, ref.ST = np.ravel(ref.S1), np.ravel(ref.ST)
        if nboot > 1:
            ref.S1_Conf_Int, ref.ST_Conf_Int = np.ravel(ref.S1_Conf_Int), np.ravel(ref.ST_Conf_Int)
    assert_parity(res, ref, what=(d, nout, N, nboot))
    # analytic: S1 = ST = W^2 sigma^2 / sum(W^2 sigma^2)   
    if nboot == 1 and N >= 20000:
        var = (ub - lb) ** 2 / 12.0
        S = W ** 2 * var[None, :]
        S = S / S.sum(axis=1, keepdims=True)
        assert np.abs(np.reshape(res.ST, (nout, d)) - S).max() < 2e-2


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
@pytest.mark.parametrize("d,nout,N,nboot", [(20, 40, 8192, 1), (7, 3, 5001, 3), (32, 2, 3000, 1), (1, 2, 1000, 2)])
def test_linear_all_estimators_and_second_order(est, d, nout, N, nboot):
    """Every Ei_estimator and the second-order indices of a linear multi-output model are moment forms of the same Gram matrix
    (fused_linear.cuh): no limit on the number of outputs, nothing evaluated per sample and output."""
    rng = np.random.default_rng(100 * d + nout)
    W = rng.normal(size=(nout, d)) + 0.5
    lb, ub = -rng.random(d), rng.random(d) + 1.0
    A, B = gsa_ref.sobol_design(N, lb, ub)
    for order in ((0, 1), (0, 1, 2)):
        ref = gsa_ref.gsa_sobol("linear", W, A, B, order=order, nboot=nboot, Ei_estimator=est)
        res = g.gsa(g.DeviceModel("linear", W), g.Sobol(order=list(order), nboot=nboot), A, B, batch=True, Ei_estimator=est)
        assert_parity(res, ref, what=(est, d, nout, N, nboot, order))


def test_linear_large_offset_inputs():
    """mean >> spread: the tile-local shift + double-double recentring keeps the variance exact."""
    d, nout, N = 6, 5, 50000
    W = np.arange(1, 31, dtype=np.float64).reshape(nout, d) / 10.0
    A, B = gsa_ref.sobol_design(N, 1000.0 * np.ones(d), 1001.0 * np.ones(d))
    ref = gsa_ref.gsa_sobol("linear", W, A, B)
    res = g.gsa(g.DeviceModel("linear", W), g.Sobol(), A, B, batch=True)
    # Adjudicated: y ~ 1e4 carries ~1e-12 of rounding, which the reference's fB*(fAB-fA) amplifies -- the fp64
    # oracle is ~5e-11 off the extended-precision value of its own formula (tests/test_oracle_golden.py
    # test_offset_inputs_fp64_model_rounding_dominates); the kernel's shifted moment form is not: 1e-12 against the exact value.
    ex = so.gsa_sobol_exact(lambda X: W.astype(np.longdouble) @ X, A, B)
    assert parity_err(res.S1, ex.S1) < TOL and parity_err(res.ST, ex.ST) < TOL, (parity_err(res.S1, ex.S1), parity_err(res.ST, ex.ST))
    assert parity_err(ref.S1, ex.S1) < 1e-9 and parity_err(ref.ST, ex.ST) < TOL


def test_linear_streamed_shards():
    """Linear model on HOST shards that start mid-replicate and span several streaming chunks: the per-launch shift row and the
    records added to the partial sums chunk by chunk must give the same result as one resident launch."""
    import torch
    d, nout, N, nboot, world = 24, 3, 900000, 2, 3            # 172.8 MB per matrix: 3 pageable chunks per shard
    rng = np.random.default_rng(7)
    W = rng.normal(size=(nout, d))
    model, method = g.DeviceModel("linear", W), g.Sobol(nboot=nboot)
    A, B = g.generate_design_matrices(N, -np.ones(d), 3.0 * np.ones(d), 2, device=True)
    whole = g.gsa(model, method, A, B, batch=True)
    An, Bn = A.cpu().numpy(), B.cpu().numpy()
    total = None
    for rank in range(world):
        c0, nc = g.shard_columns(N, rank, world)
        Al, Bl = np.asfortranarray(An[:, c0:c0 + nc]), np.asfortranarray(Bn[:, c0:c0 + nc])
        p = g.sobol_partials(model, method, Al, Bl, N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    assert_parity(g.sobol_finalize(method, total, d, N), whole)
    ref = gsa_ref.gsa_sobol("linear", W, An[:, :200000], Bn[:, :200000], nboot=nboot)
    res = g.gsa(model, method, np.asfortranarray(An[:, :200000]), np.asfortranarray(Bn[:, :200000]), batch=True)
    assert_parity(res, ref, what="host prefix vs oracle")


USER_MODEL_CU = r'''
extern "C" __device__ void my_model(const double* x, int d, const double* p, int np, double* y, int nout) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += p[i] * x[i] * x[i];
    y[0] = s + x[0] * x[1];
    if (nout > 1) y[1] = sin(x[0]) + x[d - 1];
}
'''

USER_MODEL_DEFAULT_NAME_CU = r'''
extern "C" __device__ void gsa_user_model(const double* x, int d, const double* p, int np, double* y, int nout) {
    y[0] = x[0] + p[0] * x[1] * x[2];
}
'''


def test_user_device_model_from_fatbin(tmp_path):
    """A user `__device__` model compiled to a relocatable fatbin, linked into the fused kernel with nvJitLink
    (BASELINE north star: 'user-supplied __device__ models linked from a fatbin')."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    def build(name, text):
        src = tmp_path / (name + ".cu")
        src.write_text(text)
        out = tmp_path / (name + ".fatbin")
        subprocess.check_call([nvcc, "-ccbin", "/usr/bin/g++", "-rdc=true", "-fatbin", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", str(out), str(src)])
        return out

    fat, fat1 = build("um", USER_MODEL_CU), build("um1", USER_MODEL_DEFAULT_NAME_CU)
    d, N = 5, 20011
    p = np.array([0.5, 1.0, 1.5, 2.0, 2.5])
    A, B = gsa_ref.sobol_design(N, -np.ones(d), 2 * np.ones(d))

    def f2(X):
        return np.stack([(p[:, None] * X * X).sum(axis=0) + X[0] * X[1], np.sin(X[0]) + X[d - 1]])

    # symbol with its own name -> goes through the PTX trampoline; 2 outputs, second order, replicates
    m = g.DeviceModel.from_fatbin(str(fat), "my_model", nout=2, params=p)
    res = g.gsa(m, g.Sobol(order=[0, 1, 2], nboot=3), A, B, batch=True)
    ref = so.gsa_sobol(f2, A, B, order=(0, 1, 2), nboot=3)
    assert res.S1.shape == (2, d) and res.S2.shape == (d, d, 2)
    assert_parity_adjudicated(res, ref, f2, A, B, order=(0, 1, 2), nboot=3)     # sin: CUDA's and glibc's libm differ by an ulp
    # the default symbol name, one output, another estimator
    m1 = g.DeviceModel.from_fatbin(fat1.read_bytes(), "gsa_user_model", nout=1, params=[3.0])
    res = g.gsa(m1, g.Sobol(), A, B, batch=True, Ei_estimator="Sobol2007")
    ref = so.gsa_sobol(lambda X: X[0] + 3.0 * X[1] * X[2], A, B, Ei_estimator="Sobol2007")
    assert_parity(res, ref)
    with pytest.raises(g.GsaError, match="nvJitLink"):
        g.DeviceModel.from_fatbin(fat.read_bytes(), "no_such_symbol", nout=1)          # undefined reference
    with pytest.raises(g.GsaError, match="nvJitLink"):
        g.DeviceModel.from_fatbin(fat1.read_bytes(), "other_name", nout=1)             # trampoline would redefine gsa_user_model


USER_SEPARABLE_CU = r'''
// product form: f(x) = prod_k (1 + p_k (x_k - 0.5)) ;  additive: f(x) = sum_k p_k sin(2 x_k) ... with its own symbol names
extern "C" __device__ double my_factor(int k, double x, const double* p, int np) { return 1.0 + p[k] * (x - 0.5); }
extern "C" __device__ double my_term(int k, double x, const double* p, int np) { return p[k] * x * x + 0.25 * x; }
'''
USER_SEPARABLE_DEFAULT_CU = r'''
extern "C" __device__ double gsa_user_factor(int k, double x, const double* p, int np) { return (fabs(4.0 * x - 2.0) + p[k]) / (1.0 + p[k]); }
'''


@pytest.mark.parametrize("d,N,nboot", [(5, 20011, 3), (100, 8000, 2), (40, 33334, 1), (13, 4099, 1)])
def test_user_separable_models(tmp_path, d, N, nboot):
    """gsa_model_register_separable: a user model given as its per-coordinate function (product form or additive) runs first /
    total order on the O(d) product kernel, second order and Janon2014 on the generic kernel through the same function -- all
    against the oracle evaluating the full model."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

    def build(name, text):
        src = tmp_path / (name + ".cu")
        src.write_text(text)
        out = tmp_path / (name + ".fatbin")
        subprocess.check_call([nvcc, "-ccbin", "/usr/bin/g++", "-rdc=true", "-fatbin", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", str(out), str(src)])
        return out

    fat, fatd = build("sep", USER_SEPARABLE_CU), build("sepd", USER_SEPARABLE_DEFAULT_CU)
    p = np.linspace(0.2, 1.8, d)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    fprod = lambda X: np.prod(1.0 + p[:, None] * (X - 0.5), axis=0)
    fadd = lambda X: (p[:, None] * X * X + 0.25 * X).sum(axis=0)
    mp = g.DeviceModel.from_fatbin(str(fat), "my_factor", params=p, kind="product")
    ma = g.DeviceModel.from_fatbin(str(fat), "my_term", params=p, kind="additive")
    h = g.default_handle()
    for est in ("Jansen1999", "Sobol2007", "Homma1996", "Janon2014"):
        for m, f, name in ((mp, fprod, "product"), (ma, fadd, "additive")):
            res = g.gsa(m, g.Sobol(nboot=nboot), A, B, batch=True, Ei_estimator=est)
            ref = so.gsa_sobol(f, A, B, nboot=nboot, Ei_estimator=est)
            assert_parity(res, ref, what=(name, est, d))
    if d <= 13:     # second order: the generic kernel evaluates the whole model through the user's per-coordinate function
        res = g.gsa(mp, g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True)
        assert_parity(res, so.gsa_sobol(fprod, A, B, order=(0, 1, 2), nboot=nboot), what="product second order")
    # the default symbol name (no trampoline): Sobol G written by the user == the built-in model, on the same kernel
    a = np.linspace(0.0, 9.0, d)
    mg = g.DeviceModel.from_fatbin(fatd.read_bytes(), "gsa_user_factor", params=a, kind="product")
    res = g.gsa(mg, g.Sobol(nboot=nboot), A, B, batch=True)
    assert_parity(res, gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=nboot), what="user sobol_g")
    with pytest.raises(g.GsaError, match="nvJitLink"):
        g.DeviceModel.from_fatbin(fat.read_bytes(), "no_such_symbol", kind="product")
    # an image that carries LTO-IR next to the SASS, default symbol name: the per-coordinate function is inlined into the kernel at
    # the first launch (nvJitLink -lto -kernels-used=...); same numbers
    src = tmp_path / "lto.cu"
    src.write_text(USER_SEPARABLE_DEFAULT_CU)
    fl = tmp_path / "lto.fatbin"
    subprocess.check_call([nvcc, "-ccbin", "/usr/bin/g++", "-rdc=true", "-fatbin", "-gencode", "arch=compute_100a,code=lto_100a",
                           "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(fl), str(src)])
    ml = g.DeviceModel.from_fatbin(fl.read_bytes(), "gsa_user_factor", params=a, kind="product")
    for est in ("Jansen1999", "Sobol2007"):
        res = g.gsa(ml, g.Sobol(nboot=nboot), A, B, batch=True, Ei_estimator=est)
        assert_parity(res, gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=nboot, Ei_estimator=est), what=("user sobol_g, LTO", est))


def test_multi_device_user_model(tmp_path):
    """A user `__device__` model on the single-process multi-device entry: linked once per device, same id everywhere."""
    import shutil
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2; evidence: profiles/r2_multi_gpu_tests.txt)")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = tmp_path / "um.cu"
    src.write_text(USER_MODEL_CU)
    fat = tmp_path / "um.fatbin"
    subprocess.check_call([nvcc, "-ccbin", "/usr/bin/g++", "-rdc=true", "-fatbin", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(fat), str(src)])
    d, N = 5, 30011
    p = np.array([0.5, 1.0, 1.5, 2.0, 2.5])
    A, B = gsa_ref.sobol_design(N, -np.ones(d), 2 * np.ones(d))
    m = g.GsaMulti(ndev=torch.cuda.device_count())
    model = m.model_from_fatbin(str(fat), "my_model", nout=1, params=p)
    res = m.gsa(model, g.Sobol(order=[0, 1, 2], nboot=2), A, B)
    fm = lambda X: (p[:, None] * X * X).sum(axis=0) + X[0] * X[1]
    ref = so.gsa_sobol(fm, A, B, order=(0, 1, 2), nboot=2)
    assert_parity_adjudicated(res, ref, fm, A, B, order=(0, 1, 2), nboot=2)     # the sum over k is contracted to FMAs on the device
    m.close()


def test_multi_device_entry_on_one_device(ishigami_design):
    """The gsa_multi entry (what a plain Julia ccall uses) with a single device: the same code path minus the peer gather."""
    m = g.GsaMulti(ndev=1)
    A, B = ishigami_design
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=20)
    assert_parity(m.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=20), A, B), ref)
    m.close()


def test_multi_device_single_process(ishigami_design):
    """gsa_multi: one process, all visible GPUs, one gather-and-add kernel over NVLink peer memory for the partial sums."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2; evidence: profiles/r2_multi_gpu_tests.txt)")
    m = g.GsaMulti(ndev=ndev)
    A, B = ishigami_design
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=20)
    res = m.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=20), A, B)
    assert_parity(res, ref, what=("multi", ndev))
    d, N = 100, 100003
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=3)
    res = m.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=3), A, B)
    assert_parity(res, ref, what=("multi sobol_g", ndev))
    W = np.arange(1, 61, dtype=np.float64).reshape(6, 10) / 7.0
    A, B = gsa_ref.sobol_design(30001, -np.ones(10), np.ones(10))
    ref = gsa_ref.gsa_sobol("linear", W, A, B)
    res = m.gsa(g.DeviceModel("linear", W), g.Sobol(), A, B)
    assert_parity(res, ref, what=("multi linear", ndev))
    # generated design over all devices (nothing but lb/ub goes in)
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 16)
    ref = so.gsa_sobol_p_range(lambda X: so.sobol_g_batch(X, a), [[0.0, 1.0]] * 20, 30011, nboot=3)
    res = m.gsa_p_range(g.DeviceModel("sobol_g", a), g.Sobol(nboot=3), [[0.0, 1.0]] * 20, 30011)
    assert_parity(res, ref, what=("multi generated", ndev))
    m.close()


def test_plain_c_client_of_the_abi(tmp_path):
    """include/gsa_cuda.h from plain C (gcc, no Python/torch/CUDA headers on the client side): same numbers as the Python mirror."""
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    pkg = os.path.join(root, "globalsensitivity.jl_b200")
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c_abi_smoke.c"),
                           "-o", exe, "-L", pkg, "-lgsa_cuda", "-Wl,-rpath," + pkg, "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    d, n, nboot = 4, 40000, 4
    A, B = g.generate_design_matrices(n, -np.pi * np.ones(d), np.pi * np.ones(d))
    res = g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=nboot), A, B, batch=True)
    ref = gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=(0, 1, 2), nboot=nboot)
    got = {"S1": np.zeros(d), "S1c": np.zeros(d), "ST": np.zeros(d), "STc": np.zeros(d), "S2": np.zeros((d, d)), "S2c": np.zeros((d, d))}
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] in ("S1", "ST"):
            got[t[0]][int(t[1])], got[t[0] + "c"][int(t[1])] = float(t[2]), float(t[3])
        elif t[0] == "S2":
            got["S2"][int(t[1]), int(t[2])], got["S2c"][int(t[1]), int(t[2])] = float(t[3]), float(t[4])
    assert "ERR" in r.stdout
    for name, mine, theirs in (("S1", got["S1"], res.S1), ("ST", got["ST"], res.ST), ("S2", got["S2"], res.S2),
                               ("S1c", got["S1c"], res.S1_Conf_Int), ("STc", got["STc"], res.ST_Conf_Int), ("S2c", got["S2c"], res.S2_Conf_Int)):
        assert np.array_equal(mine, theirs), name                     # the same library, the same kernels: bit-identical
    assert_parity(res, ref)


def test_peer_allreduce_two_ranks():
    """K5 over NVLink peer memory (csrc/peer.cu), one process per GPU under torchrun: sums equal NCCL's, are bit-identical on
    every rank, and a sharded analysis equals the single-GPU one."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2; evidence: profiles/r2_multi_gpu_tests.txt)")
    nproc = 2
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "peer_probe.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"e2e_max_abs_diff_vs_single_gpu"' in r.stdout


def test_one_call_step_with_fused_tail(monkeypatch):
    """gsa_sobol_run_sharded without a peer group (the shard is the whole design or a slice of it): for product-form models the
    kernel's last CTA does stage 2 and the host copy (tail.cuh) -- same numbers as the classic launch sequence, one launch."""
    import torch
    d, N = 100, 100003
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    model = g.DeviceModel("sobol_g", a)
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
    An, Bn = A.cpu().numpy(), B.cpu().numpy()
    h = g.default_handle()
    for nboot, est in ((1, "Jansen1999"), (3, "Jansen1999"), (7, "Sobol2007"), (1000, "Homma1996")):
        method = g.Sobol(nboot=nboot)
        ref = gsa_ref.gsa_sobol("sobol_g", a, An, Bn, nboot=nboot, Ei_estimator=est)
        step = g.ShardedStep(model, method, A, B, N, 0, Ei_estimator=est)
        fused = step.run()
        fused2 = step.run()                                 # the counters and the flag are monotonic over launches
        launches = h.last_timing()[2]
        # one launch for a few long replicates; many short replicates go through the classic launches (kernel + stage 2), which
        # are cheaper there (profiles/r2_tail_probe.txt)
        assert launches == 1 if nboot <= 4 else launches in (1, 2), (nboot, launches)     # (one tile per replicate needs no stage 2 either)
        h.set_stream(None)
        monkeypatch.setenv("GSA_NO_TAIL", "1")
        classic = g.gsa(model, method, A, B, batch=True, Ei_estimator=est)
        assert h.last_timing()[2] >= 1
        monkeypatch.delenv("GSA_NO_TAIL")
        assert_parity(fused, ref, what=("fused", nboot, est))
        assert_parity(classic, ref, what=("classic", nboot, est))
        for f in ("S1", "ST"):
            assert np.array_equal(getattr(fused, f), getattr(fused2, f))
    # the generated design, and a shard that is a slice (untouched replicates are written as zeros by the tail)
    method = g.Sobol(nboot=3)
    gen = g.ShardedStep(model, method, None, None, 3 * 20011, 0, p_range=(np.zeros(d), np.ones(d)), ncols=3 * 20011).run()
    g.default_handle().set_stream(None)
    assert_parity(gen, g.gsa(model, method, [[0.0, 1.0]] * d, samples=20011, batch=True))
    n = N // 3
    sl = g.ShardedStep(model, method, A[:, n + 5:2 * n - 7], B[:, n + 5:2 * n - 7], N, n + 5).run()      # inside replicate 1 only
    g.default_handle().set_stream(None)
    p = g.sobol_partials(model, method, A[:, n + 5:2 * n - 7], B[:, n + 5:2 * n - 7], N, n + 5)
    torch.cuda.synchronize()
    want = g.sobol_finalize(method, p, d, N)
    for f in ("S1", "ST"):                                   # replicates 0 and 2 are empty: NaN in both
        assert np.array_equal(np.isnan(getattr(sl, f)), np.isnan(getattr(want, f)))


def test_one_call_step_on_a_slice_of_a_long_replicate():
    """A shard that is one eighth of a long replicate (what each rank of an 8-GPU run holds): the replicate has ~7000 tiles, the
    shard ~900 of them, starting in the middle of a group -- still ONE launch (the tail counts the groups with launched tiles)."""
    import torch
    d, N = 100, 1 << 20
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol()
    h = g.default_handle()
    for col0, ncols in ((5 * N // 8 + 24, N // 8), (0, N // 8 - 8), (7 * N // 8, N // 8)):
        A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, col0=col0, ncols=ncols)
        fused = g.ShardedStep(model, method, A, B, N, col0).run()
        assert h.last_timing()[2] == 1, h.last_timing()
        h.set_stream(None)
        p = g.sobol_partials(model, method, A, B, N, col0)                 # classic launches: kernel, stage 2
        torch.cuda.synchronize()
        want = g.sobol_finalize(method, p, d, N)
        assert_parity(fused, want, what=("slice", col0, ncols))


def test_streamed_host_inputs_large():
    """Host designs larger than one streaming chunk (pinned: 256 MB per matrix; pageable: 64 MB): many chunk kernels."""
    import torch
    d, N = 100, 700001                       # 560 MB per matrix
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(nboot=2)
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
    dev = g.gsa(model, method, A, B, batch=True)
    An, Bn = A.cpu().numpy(), B.cpu().numpy()
    pageable = g.gsa(model, method, An, Bn, batch=True)
    Ap, Bp = A.t().contiguous().cpu().pin_memory().t(), B.t().contiguous().cpu().pin_memory().t()
    pinned = g.gsa(model, method, Ap, Bp, batch=True)
    ref = gsa_ref.gsa_sobol("sobol_g", a, An, Bn, nboot=2)
    for name, r in (("device", dev), ("pageable", pageable), ("pinned", pinned)):
        assert_parity(r, ref, what=name)


@pytest.mark.parametrize("d,N,nboot", [(100, 100003, 3), (100, 1000 + 7, 1), (20, 4099, 2), (40, 33334, 1), (128, 700, 1), (200, 531, 1)])
def test_product_kernel_ragged_tiles(d, N, nboot):
    """Tile lengths that are not a multiple of the sample groups per CTA: the lanes of a warp then leave the sample loop at
    different iterations (regression test for a full-mask shuffle inside that loop, which hung)."""
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=nboot)
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=nboot), A, B, batch=True)            # host input: streamed tiles
    assert_parity(res, ref, what=(d, N, nboot, "host"))
    import torch
    At, Bt = torch.from_numpy(np.ascontiguousarray(A)).cuda(), torch.from_numpy(np.ascontiguousarray(B)).cuda()
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=nboot), At, Bt, batch=True)          # device input: shard-sized tiles
    assert_parity(res, ref, what=(d, N, nboot, "device"))


def test_sharded_host_inputs_with_empty_chunks():
    """Host shards that start in the middle of a replicate: whole streaming chunks of the shard's tile range are empty
    (regression test: their tile records must be zeroed, and the copy hull must ignore empty tiles)."""
    import torch
    d, N, nboot, world = 100, 400000, 2, 3
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(nboot=nboot)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, nboot=nboot)
    total = None
    for rank in range(world):
        c0, nc = g.shard_columns(N, rank, world)
        Al, Bl = np.asfortranarray(A[:, c0:c0 + nc]), np.asfortranarray(B[:, c0:c0 + nc])      # exactly the shard, nothing beyond it
        p = g.sobol_partials(model, method, Al, Bl, N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    assert_parity(g.sobol_finalize(method, total, d, N), ref)


@pytest.mark.parametrize("d,samples,nboot,unit", [(100, 20000, 1, True), (100, 4099, 3, False), (13, 30011, 5, True), (40, 1 << 14, 2, False),
                                                    (3, 1000, 7, True), (200, 777, 1, True), (20, 1, 1, True)])
def test_fused_design_generation(d, samples, nboot, unit):
    """gsa(model, Sobol(nboot), p_range; samples) with the design generated inside the fused kernel (no A/B in memory): against the
    oracle's p_range entry (:407-417) and against the materialised-design path of this library."""
    a = np.array(([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])
    p_range = [[0.0, 1.0]] * d if unit else [[-0.25 - 0.01 * k, 1.5 + 0.02 * k] for k in range(d)]
    # the G-function is defined on the unit cube; on another box it is just a product-form test function
    ref = so.gsa_sobol_p_range(lambda X: so.sobol_g_batch(X, a), p_range, samples, nboot=nboot)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol(nboot=nboot)
    fused = g.gsa(model, method, p_range, samples=samples, batch=True)
    mat = g.gsa(model, method, p_range, samples=samples, batch=True, fused_design=False)
    if samples > 1:
        assert_parity(fused, ref, what=("fused", d, samples, nboot))
        assert_parity(mat, ref, what=("materialised", d, samples, nboot))
    else:
        assert np.all(np.isnan(fused.S1)) and np.all(np.isnan(mat.S1))     # a single sample: Var = 0/0, as in the reference
    # sharded: two column ranges that cut a replicate
    import torch
    N = nboot * samples
    total = None
    for rank in range(2):
        c0, nc = g.shard_columns(N, rank, 2)
        p = g.sobol_partials_generated(model, method, [q[0] for q in p_range], [q[1] for q in p_range], samples, c0, nc)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    if samples > 1:
        assert_parity(g.sobol_finalize(method, total, d, N), ref, what=("fused sharded", d, samples, nboot))
    # second order has no generating kernel: falls back to the materialised design transparently
    if d <= 13 and samples > 1:
        ref2 = so.gsa_sobol_p_range(lambda X: so.sobol_g_batch(X, a), p_range, samples, order=(0, 1, 2), nboot=nboot)
        assert_parity(g.gsa(model, g.Sobol(order=[0, 1, 2], nboot=nboot), p_range, samples=samples, batch=True), ref2)


@pytest.mark.parametrize("est", so.EI_ESTIMATORS)
def test_generated_design_every_model_and_order(est):
    """gsa(model, Sobol(order, nboot), p_range; samples) never materialises A/B, whatever the model, order or estimator: where no
    kernel generates its own points, K1 writes 32 MB column chunks just ahead of the regular kernel (gsa_sobol_run_generated never
    answers GSA_ERR_UNSUPPORTED any more: VERDICT r1 missing #3).  Against the oracle's p_range entry (:407-417)."""
    import ctypes as C
    cases = [
        ("ishigami", [7.0, 0.1], so.ishigami_batch, [[-np.pi, np.pi]] * 3, 20011, (0, 1, 2), 3),
        ("ishigami", [7.0, 0.1], so.ishigami_batch, [[-np.pi, np.pi]] * 4, 1048, (0, 1), 50),          # many short replicates
        ("ishi_linear", [0.0], so.ishi_linear_batch, [[-np.pi, np.pi]] * 3, 9001, (0, 1, 2), 2),       # two outputs, generic kernel
        ("sobol_g", np.linspace(0, 9, 8), lambda X: so.sobol_g_batch(X, np.linspace(0, 9, 8)), [[0.0, 1.0]] * 8, 1 << 15, (0, 1, 2), 2),
        ("sobol_g", np.linspace(0, 9, 20), lambda X: so.sobol_g_batch(X, np.linspace(0, 9, 20)), [[0.1, 0.9]] * 20, 12345, (0, 1, 2), 1),
        ("sobol_g", np.linspace(0, 9, 40), lambda X: so.sobol_g_batch(X, np.linspace(0, 9, 40)), [[0.0, 1.0]] * 40, 30011, (0, 1), 2),
    ]
    W = np.arange(1, 61, dtype=np.float64).reshape(6, 10) / 7.0
    cases.append(("linear", W, lambda X: so.linear_multi_batch(X, W), [[-1.0, 2.0]] * 10, 40001, (0, 1, 2), 2))
    for name, params, f, p_range, samples, order, nboot in cases:
        model, method = g.DeviceModel(name, params), g.Sobol(order=list(order), nboot=nboot)
        ref = so.gsa_sobol_p_range(f, p_range, samples, order=order, nboot=nboot, Ei_estimator=est)
        # straight through the C entry: it must succeed (no fallback in the Python mirror involved)
        d = len(p_range)
        nout = model.nout(d)
        o = g.sobol._opts(method, est, 0)
        rc, finish, _ = g.sobol._alloc_result(nout, d, nboot, 2 in order, nout > 1)
        lb = np.ascontiguousarray([q[0] for q in p_range], dtype=np.float64)
        ub = np.ascontiguousarray([q[1] for q in p_range], dtype=np.float64)
        code = g.lib().gsa_sobol_run_generated(g.default_handle().ptr, C.byref(o), model.model_id, model.params.ctypes.data, model.params.size,
                                               lb.ctypes.data, ub.ctypes.data, d, samples, C.byref(rc))
        assert code == 0, (name, order, est, code)
        assert_parity(finish(), ref, what=(name, order, nboot, est))


def test_host_all_y_and_all_points_are_streamed(monkeypatch):
    """The estimator-only entry and the all_points writer with HOST buffers go through the device in chunks (VERDICT r1 missing #4:
    config 4's all_y is 189 GB, config 5's all_points 5.5 TB): whole-replicate chunks, sample-range chunks, pinned and pageable."""
    import torch
    d, N = 6, 200003
    a = np.linspace(0.0, 9.0, d)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    for nboot, order in ((1, (0, 1, 2)), (7, (0, 1)), (40, (0, 1, 2))):
        n = N // nboot
        P = gsa_ref.all_points(A, B, nboot, 2 in order)
        Y = gsa_ref.model_batch("sobol_g", a, P)
        ref = gsa_ref.analyze_y(Y, d, n, order=order, nboot=nboot)
        method = g.Sobol(order=list(order), nboot=nboot)
        whole = g.gsa_sobol_all_y_analysis(method, Y, d, n)
        monkeypatch.setenv("GSA_Y_CHUNK_MB", "1")                       # 1 MB chunks: hundreds of them
        pageable = g.gsa_sobol_all_y_analysis(method, Y, d, n)
        Yp = torch.from_numpy(Y).pin_memory()
        pinned = g.gsa_sobol_all_y_analysis(method, Yp, d, n)
        monkeypatch.delenv("GSA_Y_CHUNK_MB")
        for name, r in (("one chunk", whole), ("pageable chunks", pageable), ("pinned chunks", pinned)):
            assert_parity(r, ref, what=(name, nboot, order))
        # all_points: host A/B -> host P in chunks, bit-identical to the oracle's fuse_designs
        monkeypatch.setenv("GSA_P_CHUNK_MB", "1")
        Pg = g.sobol._all_points(g.default_handle(), method, g.sobol._Mat(A), g.sobol._Mat(B), False)
        monkeypatch.delenv("GSA_P_CHUNK_MB")
        assert np.array_equal(Pg, P), (nboot, order)
    # two-output all_y (nout x cols), sample-range chunks
    Ai, Bi = gsa_ref.sobol_design(150001, -np.pi * np.ones(3), np.pi * np.ones(3))
    Y2 = gsa_ref.model_batch("ishi_linear", [0.0], gsa_ref.all_points(Ai, Bi, 2, True))
    ref = gsa_ref.analyze_y(Y2, 3, 150001 // 2, order=(0, 1, 2), nboot=2)
    monkeypatch.setenv("GSA_Y_CHUNK_MB", "2")
    assert_parity(g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=2), Y2, 3, 150001 // 2), ref, what="two outputs")
    monkeypatch.delenv("GSA_Y_CHUNK_MB")


# ------------------------------------------------------------------------------------------------ BASELINE shapes at (near) full size
def test_config3_full_size_vs_oracle():
    """BASELINE config 3 at full size: Sobol G d=8, order=[0,1,2], N=2^22, against the oracle and the analytic indices."""
    d, N = 8, 1 << 22
    a = np.array([0.0, 1.0, 4.5, 9.0, 99.0, 99.0, 99.0, 99.0])
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d))
    ref = gsa_ref.gsa_sobol("sobol_g", a, A, B, order=(0, 1, 2))
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2]), A, B, batch=True)
    assert_parity(res, ref)
    Y = gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(A, B, 1, True))        # 604 MB: the estimator-only kernel on it
    assert_parity(g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2]), Y, d, N), ref)
    Vi = 1.0 / (3.0 * (1.0 + a) ** 2)
    V = np.prod(1.0 + Vi) - 1.0
    assert np.abs(res.S1 - Vi / V).max() < 2e-4 and np.abs(res.ST - Vi * np.prod(1.0 + Vi) / (1.0 + Vi) / V).max() < 2e-4
    assert abs(res.S2[0, 1] - Vi[0] * Vi[1] / V) < 2e-4


def test_config5_shape_properties():
    """BASELINE config 5's shape (Sobol G, d=100, S1/ST) at N=2^22 (6.7 GB): size-independent properties instead of the oracle --
    shard additivity, generated == materialised design, analytic indices."""
    import torch
    d, N = 100, 1 << 22
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol()
    lb, ub = np.zeros(d), np.ones(d)
    A, B = g.generate_design_matrices(N, lb, ub, 2, device=True)
    whole = g.gsa(model, method, A, B, batch=True)
    total = None
    for rank in range(8):
        c0, nc = g.shard_columns(N, rank, 8)
        p = g.sobol_partials(model, method, A[:, c0:c0 + nc], B[:, c0:c0 + nc], N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    assert_parity(g.sobol_finalize(method, total, d, N), whole)
    assert_parity(g.gsa(model, method, [[0.0, 1.0]] * d, samples=N, batch=True), whole)      # design generated in the kernel
    Vi = 1.0 / (3.0 * (1.0 + a) ** 2)
    V = np.prod(1.0 + Vi) - 1.0
    assert np.abs(whole.S1 - Vi / V).max() < 1e-4 and np.abs(whole.ST - Vi * np.prod(1.0 + Vi) / (1.0 + Vi) / V).max() < 1e-4


def test_config4_full_size_analytic():
    """BASELINE config 4 at full size (linear, d=20, 256 outputs, N=2^22; the reference's all_y would be 189 GB): analytic
    S1 = ST = W^2 sigma^2 / sum W^2 sigma^2 and shard additivity."""
    import torch
    d, nout, N = 20, 256, 1 << 22
    o, i = np.meshgrid(np.arange(1, nout + 1), np.arange(1, d + 1), indexing="ij")
    W = 1.0 + ((o * 31 + i * 17) % 13) / 13.0
    model, method = g.DeviceModel("linear", W), g.Sobol()
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
    res = g.gsa(model, method, A, B, batch=True)
    S = W ** 2 / (W ** 2).sum(axis=1, keepdims=True)
    assert res.S1.shape == (nout, d)
    assert np.abs(res.S1 - S).max() < 2e-4 and np.abs(res.ST - S).max() < 2e-4
    total = None
    for rank in range(3):
        c0, nc = g.shard_columns(N, rank, 3)
        p = g.sobol_partials(model, method, A[:, c0:c0 + nc], B[:, c0:c0 + nc], N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    assert_parity(g.sobol_finalize(method, total, d, N), res)


# ------------------------------------------------------------------------------------------------ BASELINE configs at FULL size
# against the streamed exact-sum oracle (oracle/gsa_ref_streamed.c: full model per swapped coordinate, reference-formed terms,
# exact sums; gated against the materialising oracle and the reference's literals in tests/test_oracle_golden.py).  The oracle
# is fed the SAME A/B the kernels read, copied back from the device chunk by chunk.  Tolerance: the full 1e-12.
def _stream_oracle_from_device(model, params, A, B, chunk, **kw):
    d, N = A.shape
    s = gsa_ref.Stream(model, params, d, N, **kw)
    for c0 in range(0, N, chunk):
        c1 = min(N, c0 + chunk)
        s.push(A[:, c0:c1].t().cpu().numpy().T, B[:, c0:c1].t().cpu().numpy().T, c0)
    res = s.finish()
    s.close()
    return res


def test_config1_and_2_full_size_vs_both_oracles():
    """BASELINE configs 1 (README example: Ishigami, Sobol(), N = 600000) and 2 (order=[0,1,2], nboot=1000, N = 2^20)."""
    lb, ub = -np.pi * np.ones(3), np.pi * np.ones(3)
    ishi = g.DeviceModel("ishigami", [7.0, 0.1])
    for N, order, nboot in ((600000, (0, 1), 1), (1 << 20, (0, 1, 2), 1000)):
        A, B = g.generate_design_matrices(N, lb, ub)
        res = g.gsa(ishi, g.Sobol(order=list(order), nboot=nboot), A, B, batch=True)
        assert_parity(res, gsa_ref.gsa_sobol("ishigami", [7.0, 0.1], A, B, order=order, nboot=nboot), what=("materialising", N))
        assert_parity(res, gsa_ref.gsa_sobol_streamed("ishigami", [7.0, 0.1], A, B, order=order, nboot=nboot), what=("streamed", N))
        gen = g.gsa(ishi, g.Sobol(order=list(order), nboot=1), [[-np.pi, np.pi]] * 3, samples=N, batch=True)     # p_range entry
        exg = gsa_ref.gsa_sobol_streamed_generated("ishigami", [7.0, 0.1], lb, ub, N, order=order, nboot=1)
        assert_parity(gen, exg, what=("generated", N))


def test_config3_full_size_vs_streamed_oracle():
    d, N = 8, 1 << 22
    a = np.array([0.0, 1.0, 4.5, 9.0, 99.0, 99.0, 99.0, 99.0])
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
    res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2]), A, B, batch=True)
    ex = _stream_oracle_from_device("sobol_g", a, A, B, 1 << 20, order=(0, 1, 2))
    assert_parity(res, ex)
    for est in ("Sobol2007", "Homma1996", "Janon2014"):
        res = g.gsa(g.DeviceModel("sobol_g", a), g.Sobol(order=[0, 1, 2], nboot=4), A, B, batch=True, Ei_estimator=est)
        assert_parity(res, _stream_oracle_from_device("sobol_g", a, A, B, 1 << 20, order=(0, 1, 2), nboot=4, Ei_estimator=est), what=est)


def test_config4_full_size_vs_streamed_oracle():
    """BASELINE config 4 at FULL size (linear, d = 20, 256 outputs, N = 2^22; the reference's all_y would be 189 GB)."""
    d, nout, N = 20, 256, 1 << int(os.environ.get("GSA_TEST_C4_LOG2N", "22"))
    o, i = np.meshgrid(np.arange(1, nout + 1), np.arange(1, d + 1), indexing="ij")
    W = 1.0 + ((o * 31 + i * 17) % 13) / 13.0
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
    res = g.gsa(g.DeviceModel("linear", W), g.Sobol(), A, B, batch=True)
    ex = _stream_oracle_from_device("linear", W, A, B, 1 << 20)
    assert res.S1.shape == (nout, d)
    assert_parity(res, ex)


def test_config5_full_size_vs_streamed_oracle():
    """BASELINE config 5 at FULL size (Sobol G, d = 100, S1/ST, N = 2^26: 107 GB of design in HBM): the resident fused kernel,
    the 8-shard sum of partials (what 8 GPUs add up) and the design-generating kernel, each within 1e-12 of the oracle."""
    import torch
    d, N = 100, 1 << int(os.environ.get("GSA_TEST_C5_LOG2N", "26"))
    a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
    model, method = g.DeviceModel("sobol_g", a), g.Sobol()
    lb, ub = np.zeros(d), np.ones(d)
    A, B = g.generate_design_matrices(N, lb, ub, 2, device=True)
    whole = g.gsa(model, method, A, B, batch=True)
    total = None
    for rank in range(8):
        c0, nc = g.shard_columns(N, rank, 8)
        p = g.sobol_partials(model, method, A[:, c0:c0 + nc], B[:, c0:c0 + nc], N, c0)
        total = p.clone() if total is None else total + p
    torch.cuda.synchronize()
    sharded = g.sobol_finalize(method, total, d, N)
    gen = g.gsa(model, method, [[0.0, 1.0]] * d, samples=N, batch=True)
    # spot-check that the oracle's own generator and K1 agree at this scale (a chunk at the far end of the sequence)
    c0 = N - (1 << 12)
    Ar, Br = gsa_ref.sobol_design(N, lb, ub, 2, col0=c0, ncols=1 << 12)
    assert np.array_equal(A[:, c0:].t().cpu().numpy().T, Ar) and np.array_equal(B[:, c0:].t().cpu().numpy().T, Br)
    ex = _stream_oracle_from_device("sobol_g", a, A, B, 1 << 19)
    for name, r in (("resident", whole), ("8 shards", sharded), ("generated", gen)):
        assert_parity(r, ex, what=name)
    Vi = 1.0 / (3.0 * (1.0 + a) ** 2)
    V = np.prod(1.0 + Vi) - 1.0
    assert np.abs(whole.S1 - Vi / V).max() < 1e-4


@pytest.mark.parametrize("d", [13, 16, 17, 24, 32, 33])
def test_analyze_y_mid_d_first_order(d):
    """First-order estimator-on-Y for 12 < d <= 32 runs on the register kernel (Jansen) -- d = 33 and the three-term estimators
    beyond d = 16 take the shared-memory kernel; all must agree with the oracle."""
    N, nboot = 6007, 2
    a = np.linspace(0.0, 30.0, d)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    Y = gsa_ref.model_batch("sobol_g", a, gsa_ref.all_points(A, B, nboot, False))
    for est in so.EI_ESTIMATORS:
        ref = gsa_ref.analyze_y(Y, d, N // nboot, nboot=nboot, Ei_estimator=est)
        res = g.gsa_sobol_all_y_analysis(g.Sobol(nboot=nboot), Y, d, N // nboot, est)
        assert_parity(res, ref, what=(d, est))


@pytest.mark.parametrize("d,second", [(4, True), (6, False), (30, False)])
def test_variance_robust_when_mean_dominates(d, second):
    """mean^2 / var ~ 1e16: a one-pass variance in plain fp64 would return noise; the kernels' shifted, compensated sums
    (un-shifted exactly in double-double) must reproduce the reference's two-pass var (:183).  ST (Jansen) isolates Var."""
    N, nboot = 40000, 2
    a = np.linspace(0.0, 9.0, d)
    A, B = gsa_ref.sobol_design(N, np.zeros(d), np.ones(d))
    order = (0, 1, 2) if second else (0, 1)
    P = gsa_ref.all_points(A, B, nboot, second)
    Y = 1.0e5 + 1.0e-3 * gsa_ref.model_batch("sobol_g", a, P)
    ref = gsa_ref.analyze_y(Y, d, N // nboot, order=order, nboot=nboot)
    ex = gsa_ref.analyze_y(Y, d, N // nboot, order=order, nboot=nboot, exact=True)
    res = g.gsa_sobol_all_y_analysis(g.Sobol(order=order, nboot=nboot), Y, d, N // nboot)
    # Adjudicated: the reference's two-pass var is exact here (fp64 oracle == exact-sum oracle to 1e-13,
    # tests/test_oracle_golden.py test_mean_dominated_data_all_three_evaluations_agree_on_ST), so the full 1e-12 applies.
    assert parity_err(ref.ST, ex.ST) < 1e-13
    assert parity_err(res.ST, ex.ST) < TOL, parity_err(res.ST, ex.ST)
    assert parity_err(res.ST_Conf_Int, ex.ST_Conf_Int) < TOL, parity_err(res.ST_Conf_Int, ex.ST_Conf_Int)


def test_empty_and_degenerate_inputs():
    """Empty designs and one-column designs do not crash: the reference returns NaN (0/0) there, and so do we."""
    ishi = g.DeviceModel("ishigami", [7.0, 0.1])
    A = np.zeros((3, 0), order="F")
    res = g.gsa(ishi, g.Sobol(order=[0, 1, 2]), A, A.copy(order="F"), batch=True)
    assert res.S1.shape == (3,) and np.all(np.isnan(res.S1)) and res.S2.shape == (3, 3)
    A, B = gsa_ref.sobol_design(1, -np.pi * np.ones(3), np.pi * np.ones(3))
    res = g.gsa(ishi, g.Sobol(), A, B, batch=True)
    assert np.all(np.isnan(res.S1)) or np.all(np.isinf(res.S1))          # Var of two values / zero spread: not finite
    res = g.gsa_sobol_all_y_analysis(g.Sobol(), np.zeros(0), 3, 0)
    assert np.all(np.isnan(res.S1))
    W = np.ones((2, 3))
    res = g.gsa(g.DeviceModel("linear", W), g.Sobol(), np.zeros((3, 0), order="F"), np.zeros((3, 0), order="F"), batch=True)
    assert res.S1.shape == (2, 3) and np.all(np.isnan(res.S1))
    assert g.generate_design_matrices(1, np.zeros(2), np.ones(2))[0].shape == (2, 1)
