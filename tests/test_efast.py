"""eFAST (reference src/eFAST_sensitivity.jl:75-213): the oracle against the reference's literals (CPU) and the
CUDA path against the oracle (GPU)."""
import numpy as np
import pytest

import efast_oracle as ef
import sobol_oracle as so

PI = np.pi
P4 = [[-PI, PI]] * 4


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_reproduces_reference_literals(seed):
    """test/eFAST_method.jl:34-38, :66-70, :98-113 -- the reference draws the phases at random, so its goldens hold for any phase
    to its own atol = 1e-4."""
    phi = 2.0 * np.random.default_rng(seed).random(4)
    r = ef.gsa_efast(so.ishigami_batch, P4, 15000, phi=phi)
    assert np.abs(r.S1 - np.array([[0.307599, 0.442412, 3.0941e-25, 3.42372e-28]])).max() < 1e-4
    assert np.abs(r.ST - np.array([[0.556244, 0.446861, 0.239259, 0.027099]])).max() < 1e-4
    r = ef.gsa_efast(so.linear_batch, P4, 15000, phi=phi)
    assert np.abs(r.S1 - np.array([[0.997504, 0.000203575, 2.1599e-10, 2.18296e-10]])).max() < 1e-4
    assert np.abs(r.ST - np.array([[0.999796, 0.000204698, 7.26874e-7, 7.59996e-7]])).max() < 1e-4
    r = ef.gsa_efast(so.ishi_linear_batch, P4, 15000, phi=phi)                       # multi-output layout: nout x d
    assert r.S1.shape == (2, 4)
    assert np.abs(r.S1 - np.array([[0.307598, 0.442411, 0, 0], [0.997498, 0.000203571, 0, 0]])).max() < 1e-4
    assert np.abs(r.ST - np.array([[0.556243, 0.446861, 0.239258, 0.0271024], [0.999796, 0.00020404, 6.35579e-8, 6.36016e-8]])).max() < 1e-4


def test_oracle_frequencies_and_design():
    assert ef.efast_frequencies(4, 15000).tolist() == [1874, 1, 117, 234]          # floor.(range(1, 234, length = 3))
    assert ef.efast_frequencies(40, 300, 4).tolist()[:5] == [37, 1, 2, 3, 4]       # m = 4 < d - 1: (k mod m) + 1
    P = ef.efast_design([[0.0, 1.0], [2.0, 2.0], [-1.0, 3.0]], 301, phi=[0.1, 0.7, 1.3])
    assert P.shape == (3, 903) and np.all(P[1] == 2.0)                               # degenerate range: constant (:113)
    assert P[0].min() >= 0.0 and P[0].max() <= 1.0 and P[2].min() >= -1.0 and P[2].max() <= 3.0


# ------------------------------------------------------------------------------------------------------------------ GPU
def _parity(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


@pytest.mark.gpu
@pytest.mark.parametrize("samples,H", [(15000, 4), (15003, 4), (301, 4), (300, 4), (4099, 2), (1 << 16, 6), (3000, 4), (1 << 14, 4), (16875, 4), (405, 4),
                                       (2 * 3 ** 9, 4)])
def test_efast_analysis_matches_oracle(samples, H):
    """gsa_efast_all_y_analysis on the GPU against numpy's rfft restatement, on the same all_y.  Sizes whose half (even) or whole
    (odd) factors into 2, 3, 5 take the Stockham FFT (every radix and the real-input first pass are covered here); 15003, 301 and
    4099 take the pruned DFT."""
    import gsa_loader
    g = gsa_loader.load()
    phi = [0.3, 1.1, 0.7, 1.9]
    P = ef.efast_design(P4, samples, H, phi)
    for f in (so.ishigami_batch, so.ishi_linear_batch):
        Y = f(P)
        ref = ef.efast_analysis(Y, 4, samples, H)
        res = g.gsa_efast_all_y_analysis(g.eFAST(H), Y, 4, samples)
        assert _parity(res.S1, ref.S1) <= 1e-12 and _parity(res.ST, ref.ST) <= 1e-12, (_parity(res.S1, ref.S1), _parity(res.ST, ref.ST))
    # a large offset does not matter (the spectrum is taken of the centred series)
    Y = 1.0e6 + so.ishigami_batch(P)
    ref = ef.efast_analysis(Y - Y.mean(), 4, samples, H)
    res = g.gsa_efast_all_y_analysis(g.eFAST(H), Y, 4, samples)
    assert _parity(res.S1, ref.S1) <= 1e-9 and _parity(res.ST, ref.ST) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("samples", [15000, 16875, 1 << 13, 600])
def test_efast_fft_and_pruned_dft_agree(samples, monkeypatch):
    """the two spectrum routes (efast_fft.cuh / efast.cuh) on the same series: 1e-13"""
    import gsa_loader
    g = gsa_loader.load()
    rng = np.random.default_rng(samples)
    Y = rng.normal(size=(2, samples * 5)) + 3.0
    fft = g.gsa_efast_all_y_analysis(g.eFAST(4), Y, 5, samples)
    assert g.default_handle().last_timing()[2] > 3                     # one launch per radix + centre, power, indices
    monkeypatch.setenv("GSA_EFAST_DFT", "1")
    dft = g.gsa_efast_all_y_analysis(g.eFAST(4), Y, 5, samples)
    assert g.default_handle().last_timing()[2] == 3
    assert _parity(fft.S1, dft.S1) <= 1e-13 and _parity(fft.ST, dft.ST) <= 1e-13, (_parity(fft.S1, dft.S1), _parity(fft.ST, dft.ST))


@pytest.mark.gpu
def test_efast_design_and_device_models():
    """the design against the oracle (transcendentals: a few ulp), the fused device-model entry and the closure entry"""
    import gsa_loader
    g = gsa_loader.load()
    phi = np.array([0.3, 1.1, 0.7, 1.9])
    P = g.efast_design(g.eFAST(), P4, 15000, phi=phi)
    Pr = ef.efast_design(P4, 15000, 4, phi)
    # asin(sinpi(.)) is ill-conditioned at the turning points of a search curve (d asin / dx -> infinity at |x| = 1): an ulp of
    # difference between CUDA's and NumPy's sinpi shows up as ~1e-13 there -- any two libms differ like this, Julia's included
    assert P.shape == Pr.shape and np.abs(P - Pr).max() < 1e-12
    ref = ef.gsa_efast(so.ishigami_batch, P4, 15000, phi=phi)
    res = g.gsa_efast(g.DeviceModel("ishigami", [7.0, 0.1]), g.eFAST(), P4, samples=15000, phi=phi)
    assert res.S1.shape == (1, 4)
    assert _parity(res.S1, ref.S1) <= 1e-11 and _parity(res.ST, ref.ST) <= 1e-11           # the model sees design points that differ by ulps
    assert np.abs(res.S1 - np.array([[0.307599, 0.442412, 0, 0]])).max() < 1e-4          # the reference's literal (test/eFAST_method.jl:34)
    res2 = g.gsa_efast(so.ishigami_batch, g.eFAST(), P4, samples=15000, batch=True, phi=phi)     # closure on the GPU-built design
    assert _parity(res2.S1, ref.S1) <= 1e-11 and _parity(res2.ST, ref.ST) <= 1e-11
    res3 = g.gsa_efast(g.DeviceModel("ishi_linear"), g.eFAST(), P4[:3], samples=9003, phi=phi[:3])
    ref3 = ef.gsa_efast(so.ishi_linear_batch, P4[:3], 9003, phi=phi[:3])
    assert res3.S1.shape == (2, 3) and _parity(res3.S1, ref3.S1) <= 1e-11 and _parity(res3.ST, ref3.ST) <= 1e-11
    a = np.linspace(0.0, 9.0, 20)
    res4 = g.gsa_efast(g.DeviceModel("sobol_g", a), g.eFAST(), [[0.0, 1.0]] * 20, samples=20000, phi=np.linspace(0.1, 1.9, 20))
    ref4 = ef.gsa_efast(lambda X: so.sobol_g_batch(X, a), [[0.0, 1.0]] * 20, 20000, phi=np.linspace(0.1, 1.9, 20))
    assert _parity(res4.S1, ref4.S1) <= 1e-11 and _parity(res4.ST, ref4.ST) <= 1e-11
    with pytest.raises(g.GsaError, match="too small"):
        g.gsa_efast(g.DeviceModel("ishigami", [7.0, 0.1]), g.eFAST(), P4, samples=40, phi=phi)
    # samples = 15001: harmonic 4 of omega = 1875 is bin 7500 = samples/2, one past the reference's buffer (a BoundsError there,
    # an IndexError in the oracle, an error code here)
    with pytest.raises(IndexError):
        ef.gsa_efast(so.ishigami_batch, P4, 15001, phi=phi)
    with pytest.raises(g.GsaError, match="BoundsError"):
        g.gsa_efast(g.DeviceModel("ishigami", [7.0, 0.1]), g.eFAST(), P4, samples=15001, phi=phi)
    # random phases by default: results move within the reference's own tolerance
    r5 = g.gsa_efast(g.DeviceModel("ishigami", [7.0, 0.1]), g.eFAST(), P4, samples=15000)
    assert np.abs(r5.ST - np.array([[0.556244, 0.446861, 0.239259, 0.027099]])).max() < 1e-4
