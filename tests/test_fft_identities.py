"""CPU checks of the algebra behind csrc/efast_fft.cuh (pure NumPy, no GPU, no oracle): the mixed-radix Stockham (autosort)
addressing -- one pass per radix, butterfly j reads x[j + t*M/R], twiddles exp(-2 pi i t k / (Ls R)) with k = j mod Ls, writes
y[(j - k) R + k + t Ls] -- and the untangling of a real series packed as M = S/2 complex points, against numpy.fft
(the reference's rfft, src/eFAST_sensitivity.jl:185)."""
import numpy as np
import pytest


def plan(M):
    """efast_fft_plan: radices 4 (pairs of twos), 2, 5, 3 -- or None when M has another prime factor"""
    r, twos = [], 0
    if M < 2:
        return None
    while M % 2 == 0:
        M //= 2
        twos += 1
    r += [4] * (twos // 2) + [2] * (twos % 2)
    while M % 5 == 0:
        M //= 5
        r.append(5)
    while M % 3 == 0:
        M //= 3
        r.append(3)
    return r if M == 1 else None


def stockham(x, radices):
    """the passes of efast_fft_pass_kernel, vectorised over the butterflies j"""
    M = x.size
    Ls = 1
    for R in radices:
        q = M // R
        j = np.arange(q)
        k = j % Ls
        v = np.stack([x[j + t * q] * np.exp(-2j * np.pi * ((t * k) / (Ls * R))) for t in range(R)])        # [R][q]
        W = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)                                # the R-point transform
        v = W @ v
        y = np.empty(M, dtype=complex)
        for t in range(R):
            y[(j - k) * R + k + t * Ls] = v[t]
        x, Ls = y, Ls * R
    return x


@pytest.mark.parametrize("M", [2, 4, 8, 150, 7500, 1 << 12, 16875, 405, 3 ** 9, 6, 10, 1000])
def test_stockham_mixed_radix_is_the_dft(M):
    rng = np.random.default_rng(M)
    x = rng.normal(size=M) + 1j * rng.normal(size=M)
    r = plan(M)
    assert r is not None and int(np.prod(r)) == M
    got, want = stockham(x, r), np.fft.fft(x)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_plan_rejects_other_primes():
    for M in (7, 14, 4099, 15003, 2 * 1667, 1):
        assert plan(M) is None
    assert plan(7500) == [4, 5, 5, 5, 5, 3] and plan(65536) == [4] * 8 and plan(32768) == [4] * 7 + [2]


@pytest.mark.parametrize("S", [300, 15000, 1 << 13, 3000])
def test_packed_real_transform_untangles_to_rfft(S):
    """S even: z_n = y_2n + i y_2n+1, Z = FFT_M(z), M = S/2; Y_k = E_k + W^k O_k with E_k = (Z_k + conj Z_{M-k}) / 2,
    O_k = (Z_k - conj Z_{M-k}) / 2i, W = exp(-2 pi i / S)  (efast_fft_power_kernel), for the bins 1 .. M-1 the analysis reads"""
    rng = np.random.default_rng(S)
    y = rng.normal(size=S)
    M = S // 2
    Z = stockham(y[0::2] + 1j * y[1::2], plan(M))
    k = np.arange(1, M)
    zk, zm = Z[k], Z[M - k]
    E = 0.5 * (zk + np.conj(zm))
    O = (zk - np.conj(zm)) / 2j
    Y = E + np.exp(-2j * np.pi * k / S) * O
    want = np.fft.rfft(y)[1:M]
    assert np.abs(Y - want).max() <= 1e-12 * np.abs(want).max()
    # the component form the kernel uses: e = ((zk.x + zm.x)/2, (zk.y - zm.y)/2), o = ((zk.y + zm.y)/2, -(zk.x - zm.x)/2)
    e = 0.5 * (zk.real + zm.real) + 0.5j * (zk.imag - zm.imag)
    o = 0.5 * (zk.imag + zm.imag) - 0.5j * (zk.real - zm.real)
    assert np.abs(e - E).max() <= 1e-15 * np.abs(E).max() and np.abs(o - O).max() <= 1e-15 * np.abs(O).max()


def test_small_transforms_of_the_kernel():
    """the hand-written 3- and 5-point transforms (efast_dft<3>, efast_dft<5>) against the matrix form"""
    rng = np.random.default_rng(5)
    v = rng.normal(size=5) + 1j * rng.normal(size=5)
    # R = 3
    a, b, c = v[:3]
    h = 0.86602540378443865
    s, df = b + c, b - c
    m, n = a - 0.5 * s, -1j * h * df
    got3 = np.array([a + s, m + n, m - n])
    assert np.abs(got3 - np.fft.fft(v[:3])).max() < 1e-15 * 10
    # R = 5
    c1, c2, s1, s2 = 0.30901699437494742, -0.80901699437494742, 0.95105651629515357, 0.58778525229247313
    a, b, c, d, e = v
    t1, t2, t3, t4 = b + e, c + d, b - e, c - d
    m1, m2 = a + c1 * t1 + c2 * t2, a + c2 * t1 + c1 * t2
    n1, n2 = s1 * t3 + s2 * t4, s2 * t3 - s1 * t4
    got5 = np.array([a + t1 + t2, m1 - 1j * n1, m2 - 1j * n2, m2 + 1j * n2, m1 + 1j * n1])
    assert np.abs(got5 - np.fft.fft(v)).max() < 1e-15 * 10
