"""CPU oracle (NumPy) for the Sobol path of GlobalSensitivity.jl  --  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, what the reference computes for
    gsa(f, Sobol(order, nboot, conf_level), A, B; batch=true)
so that the CUDA path can be checked against it.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product
path (`globalsensitivity.jl_b200/`) never does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this file against the literal
golden vectors in the reference's own test-suite (`test/sobol_method.jl:41-45` S2 and
`:55-71` confidence intervals, reproduced to <= 1e-15) -- which also pins the Sobol-point
convention of the third-party generator (QuasiMonteCarlo.jl -> Sobol.jl, versions unpinned
by the reference: `Project.toml:34`, no Manifest).

Shapes follow Julia: a design is `d x N` (one sample per COLUMN), model outputs are a vector
of length `cols` or a matrix `nout x cols`.

Reference lines followed (all in /root/reference/src/sobol_sensitivity.jl):
  fuse_designs                  :94-108
  gsa(f, ::Sobol, A, B)         :110-168   (batch branch :142-149)
  gsa_sobol_all_y_analysis      :169-405
  gsa(f, ::Sobol, p_range)      :407-417
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                      "globalsensitivity.jl_b200", "data", "joe_kuo_21201.bin")

EI_ESTIMATORS = ("Homma1996", "Sobol2007", "Jansen1999", "Janon2014")


# --------------------------------------------------------------------------------------
# Sobol sequence (third-party in the reference: Sobol.jl via QuasiMonteCarlo.jl).
# Published algorithm: Bratley & Fox Alg. 659 with the Joe-Kuo (2008) `new-joe-kuo-6.21201`
# direction numbers, Gray-code order (Antonov-Saleev).  Call sites in the reference:
# src/sobol_sensitivity.jl:408-413, README.md:55-56, test/sobol_method.jl:29-30.
# --------------------------------------------------------------------------------------
def direction_numbers(ndim: int) -> np.ndarray:
    """V[j, b] (uint32), j < ndim, b < 32: 32-bit direction numbers, MSB-aligned."""
    if not 1 <= ndim <= 21201:
        raise ValueError("Sobol dimension must be in 1..21201")
    rec = np.fromfile(_TABLE, dtype="<u4").reshape(21201, 19)
    V = np.zeros((ndim, 32), dtype=np.uint64)
    for j in range(ndim):
        m = [0] * 32
        if j == 0:
            m = [1] * 32
        else:
            p = int(rec[j, 0])
            deg = p.bit_length() - 1
            for i in range(deg):
                m[i] = int(rec[j, 1 + i])
            for i in range(deg, 32):
                v = m[i - deg] ^ (m[i - deg] << deg)
                for k in range(1, deg):
                    if (p >> (deg - k)) & 1:
                        v ^= m[i - k] << k
                m[i] = v
        for i in range(32):
            V[j, i] = (m[i] << (31 - i)) & 0xFFFFFFFF
    return V.astype(np.uint32)


def sobol_ints(first_index: int, n: int, V: np.ndarray) -> np.ndarray:
    """uint32 Sobol integers, shape (ndim, n), for 0-based sequence indices
    first_index .. first_index+n-1 (index 0 = the all-zero point)."""
    t = np.arange(first_index, first_index + n, dtype=np.uint64)
    g = t ^ (t >> np.uint64(1))
    x = np.zeros((V.shape[0], n), dtype=np.uint32)
    for b in range(int(first_index + n).bit_length()):
        mask = ((g >> np.uint64(b)) & np.uint64(1)).astype(bool)
        x[:, mask] ^= V[:, b:b + 1]
    return x


def sobol_first_index(n: int) -> int:
    """QuasiMonteCarlo.sample(n, d, SobolSample()) skips so that the first emitted point has
    0-based index 2^floor(log2 n)  (pinned by the reference's CI goldens)."""
    return 1 << (int(n).bit_length() - 1)


def generate_design_matrices(n: int, lb, ub, num_mats: int = 2, col0: int = 0,
                             ncols: Optional[int] = None):
    """QuasiMonteCarlo.generate_design_matrices(n, lb, ub, SobolSample(), num_mats):
    one (num_mats*d)-dimensional unrandomised Sobol point set split by rows into num_mats
    d x n matrices, each scaled as (ub-lb)*u + lb with a separately rounded mul and add."""
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    d = lb.shape[0]
    if ncols is None:
        ncols = n - col0
    V = direction_numbers(num_mats * d)
    x = sobol_ints(sobol_first_index(n) + col0, ncols, V)
    u = x.astype(np.float64) * 2.0 ** -32
    out = []
    for m in range(num_mats):
        um = u[m * d:(m + 1) * d]
        out.append((ub - lb)[:, None] * um + lb[:, None])
    return out


# --------------------------------------------------------------------------------------
# Julia Base numerics used by the analysis
# --------------------------------------------------------------------------------------
def jl_sum(x: np.ndarray) -> float:
    """Base.sum on a vector: pairwise, sequential below 1024 elements
    (Base.mapreduce_impl, pairwise_blocksize = 1024)."""
    x = np.ascontiguousarray(x, dtype=np.float64)

    def rec(lo, hi):
        if hi - lo <= 1024:
            # sequential left-to-right (np.cumsum is strictly sequential)
            return float(np.cumsum(x[lo:hi])[-1])
        mid = (lo + hi) >> 1
        return rec(lo, mid) + rec(mid, hi)

    if x.size == 0:
        return 0.0
    return rec(0, x.size)


def _sum(x, axis=None):
    if axis is None:
        return jl_sum(np.ravel(x))
    return np.sum(x, axis=axis)


def jl_var(x: np.ndarray) -> float:
    """Statistics.var: corrected, two-pass."""
    m = jl_sum(x) / x.size
    return jl_sum((x - m) ** 2) / (x.size - 1)


def norm_quantile(p: float) -> float:
    """Distributions.quantile(Normal(0,1), p)."""
    from scipy.special import ndtri
    return float(ndtri(p))


# --------------------------------------------------------------------------------------
# The path itself
# --------------------------------------------------------------------------------------
@dataclass
class SobolResult:
    """Same field order as the reference's mutable struct (:85-92)."""
    S1: np.ndarray
    S1_Conf_Int: Optional[np.ndarray]
    S2: Optional[np.ndarray]
    S2_Conf_Int: Optional[np.ndarray]
    ST: np.ndarray
    ST_Conf_Int: Optional[np.ndarray]


def fuse_designs(A: np.ndarray, B: np.ndarray, second_order: bool = False) -> np.ndarray:
    """:94-108  -> d x n*(d+2) or d x n*(2d+2): [A, B, AB_1..AB_d (, BA_1..BA_d)]."""
    d = A.shape[0]
    blocks = [A, B]
    for i in range(d):
        Ab = A.copy()
        Ab[i, :] = B[i, :]
        blocks.append(Ab)
    if second_order:
        for i in range(d):
            Ba = B.copy()
            Ba[i, :] = A[i, :]
            blocks.append(Ba)
    return np.concatenate(blocks, axis=1)


def all_points(A: np.ndarray, B: np.ndarray, nboot: int, second_order: bool) -> np.ndarray:
    """:116-134  replicate chunking + fusion.  Trailing N mod nboot columns are dropped."""
    d, N = A.shape
    n = N // nboot
    parts = [fuse_designs(A[:, n * i:n * (i + 1)], B[:, n * i:n * (i + 1)], second_order)
             for i in range(nboot)]
    return np.concatenate(parts, axis=1)


def gsa_sobol_all_y_analysis(all_y: np.ndarray, d: int, n: int, order: Sequence[int] = (0, 1),
                             nboot: int = 1, conf_level: float = 0.95,
                             Ei_estimator: str = "Jansen1999") -> SobolResult:
    """:169-405.  `all_y` is a vector (scalar model) or an `nout x cols` matrix."""
    if Ei_estimator not in EI_ESTIMATORS:
        raise ValueError(f"unknown Ei_estimator {Ei_estimator!r}")
    all_y = np.asarray(all_y, dtype=np.float64)
    multi = all_y.ndim == 2
    second = 2 in order
    step = 2 * d + 2 if second else d + 2
    Y = all_y if multi else all_y[None, :]
    nout = Y.shape[0]

    Var = np.zeros((nboot, nout))
    Vi = np.zeros((nboot, nout, d))
    Ei = np.zeros((nboot, nout, d))
    Vij = np.zeros((nboot, nout, d, d))
    for r in range(nboot):
        i0 = r * step
        for o in range(nout):
            y = Y[o]
            blk = lambda b: y[(i0 + b) * n:(i0 + b + 1) * n]
            Var[r, o] = jl_var(y[i0 * n:(i0 + 2) * n])                      # :183 / :241
            fA, fB = blk(0), blk(1)                                        # :185-186
            fAi = [blk(2 + k) for k in range(d)]                           # :187
            fBi = [blk(2 + d + k) for k in range(d)] if second else None   # :189
            for k in range(d):
                Vi[r, o, k] = jl_sum(fB * (fAi[k] - fA)) / n               # :192
            if second:
                for k in range(d):
                    for j in range(k + 1, d):
                        Vij[r, o, k, j] = jl_sum(fBi[k] * fAi[j] - fA * fB) / n   # :197
            for k in range(d):
                if Ei_estimator == "Homma1996":                            # :203-209
                    mA = jl_sum(fA) / n
                    Ei[r, o, k] = (jl_sum((fA - mA) ** 2) / (n - 1)
                                   - jl_sum(fA * fAi[k]) / n + mA ** 2)
                elif Ei_estimator == "Sobol2007":                          # :211
                    Ei[r, o, k] = jl_sum(fA * (fA - fAi[k])) / n
                elif Ei_estimator == "Jansen1999":                         # :213
                    Ei[r, o, k] = jl_sum((fA - fAi[k]) ** 2) / (2 * n)
                else:                                                      # Janon2014 :215-235
                    s2 = jl_sum(fA ** 2 + fAi[k] ** 2)
                    s1 = jl_sum(fA + fAi[k])
                    sp = jl_sum(fA * fAi[k])
                    mh = (1.0 / n) * jl_sum((fA + fAi[k]) / 2)
                    Ei[r, o, k] = ((s2 / (2 * n) - (s1 / (2 * n)) ** 2)
                                   * (1.0 - ((1.0 / n) * sp - mh ** 2)
                                      / ((1.0 / n) * jl_sum((fA ** 2 + fAi[k] ** 2) / 2) - mh ** 2)))
    # :318-350
    Si = Vi / Var[:, :, None]
    Ti = Ei / Var[:, :, None]
    Sij = None
    if second:
        M = np.zeros_like(Vij)
        for k in range(d):
            for j in range(k + 1, d):
                M[:, :, k, j] = Vi[:, :, k] + Vi[:, :, j]
        Sij = (Vij - M) / Var[:, :, None, None]

    S1ci = STci = S2ci = None
    if nboot > 1:                                                          # :351-379
        z = norm_quantile((1 + conf_level) / 2)

        def mean_ci(x):  # x: (nboot, ...)
            m = np.sum(x, axis=0) / nboot
            sd = np.sqrt(np.sum((x - m) ** 2, axis=0) / (nboot - 1))
            return m, z * sd / np.sqrt(nboot)

        S1, S1ci = mean_ci(Si)
        ST, STci = mean_ci(Ti)
        if second:
            S2, S2ci = mean_ci(Sij)
    else:
        S1, ST = Si[0], Ti[0]
        if second:
            S2 = Sij[0]

    # result shapes (:397-404): scalar -> (d,), (d,d); multi -> (nout,d), (d,d,nout)
    def shp1(x):
        return None if x is None else (x if multi else x[0])

    def shp2(x):
        if x is None:
            return None
        return np.transpose(x, (1, 2, 0)) if multi else x[0]

    return SobolResult(shp1(S1), shp1(S1ci), shp2(S2) if second else None,
                       shp2(S2ci) if second else None, shp1(ST), shp1(STci))


def gsa_sobol(f: Callable[[np.ndarray], np.ndarray], A: np.ndarray, B: np.ndarray,
              order: Sequence[int] = (0, 1), nboot: int = 1, conf_level: float = 0.95,
              Ei_estimator: str = "Jansen1999") -> SobolResult:
    """gsa(f, Sobol(order, nboot, conf_level), A, B; batch=true)   :110-149."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    d, N = A.shape
    n = N // nboot
    pts = all_points(A, B, nboot, 2 in order)
    all_y = f(pts)                                                         # :143
    return gsa_sobol_all_y_analysis(all_y, d, n, order, nboot, conf_level, Ei_estimator)


def gsa_sobol_p_range(f, p_range, samples: int, order=(0, 1), nboot: int = 1,
                      conf_level: float = 0.95, Ei_estimator: str = "Jansen1999") -> SobolResult:
    """gsa(f, ::Sobol, p_range; samples)  :407-417 -- a 2*nboot*d-dimensional design; matrices
    1..nboot concatenated column-wise form A, nboot+1..2nboot form B."""
    lb = [float(p[0]) for p in p_range]
    ub = [float(p[1]) for p in p_range]
    AB = generate_design_matrices(samples, lb, ub, 2 * nboot)
    A = np.concatenate(AB[:nboot], axis=1)
    B = np.concatenate(AB[nboot:], axis=1)
    return gsa_sobol(f, A, B, order, nboot, conf_level, Ei_estimator)


# --------------------------------------------------------------------------------------
# Batch models used by the reference's tests / BASELINE configs
# --------------------------------------------------------------------------------------
def ishigami_batch(X, a=7.0, b=0.1):
    """test/sobol_method.jl:4-8."""
    return np.sin(X[0]) + a * np.sin(X[1]) ** 2 + b * X[2] ** 4 * np.sin(X[0])


def linear_batch(X, a=7.0, b=0.1):
    """test/sobol_method.jl:15-19."""
    return a * X[0] + b * X[1]


def ishi_linear_batch(X):
    """test/sobol_method.jl:153-158 (2 x cols)."""
    return np.stack([ishigami_batch(X), linear_batch(X)], axis=0)


def sobol_g_batch(X, a):
    """Sobol G-function prod_i (|4 x_i - 2| + a_i) / (1 + a_i), left-to-right product."""
    a = np.asarray(a, dtype=np.float64)
    y = np.ones(X.shape[1])
    for i in range(X.shape[0]):
        y = y * ((np.abs(4.0 * X[i] - 2.0) + a[i]) / (1.0 + a[i]))
    return y


def linear_multi_batch(X, W):
    """y = W x, W is nout x d; accumulated k = 0..d-1 in order."""
    W = np.asarray(W, dtype=np.float64)
    y = np.zeros((W.shape[0], X.shape[1]))
    for k in range(X.shape[0]):
        y += W[:, k:k + 1] * X[k][None, :]
    return y


# --------------------------------------------------------------------------------------
# "Exact" evaluation of the reference's formulas, used to ADJUDICATE the few tests whose tolerance is looser
# than 1e-12: the model, every per-sample term and every sum are evaluated in extended precision (x87 long
# double, 64-bit significand) with exactly rounded `math.fsum`-style sums where it matters, and rounded to
# fp64 once at the end.  If the CUDA path is closer to this than the fp64 oracle is, the difference is the
# oracle's (i.e. the reference's own fp64 rounding), not the kernel's.
# --------------------------------------------------------------------------------------
def gsa_sobol_exact(f: Callable[[np.ndarray], np.ndarray], A, B, order: Sequence[int] = (0, 1), nboot: int = 1,
                    conf_level: float = 0.95, Ei_estimator: str = "Jansen1999", all_y=None, d=None, n=None) -> SobolResult:
    """`f` must accept and return np.longdouble arrays (NumPy ufuncs do).  With `all_y` given (fp64 data), the
    analysis alone is evaluated in extended precision."""
    L = np.longdouble
    second = 2 in order
    if all_y is None:
        A = np.asarray(A, dtype=np.float64)
        B = np.asarray(B, dtype=np.float64)
        d, N = A.shape
        n = N // nboot
        all_y = np.asarray(f(all_points(A, B, nboot, second).astype(L)), dtype=L)
    else:
        all_y = np.asarray(all_y).astype(L)
    multi = all_y.ndim == 2
    Y = all_y if multi else all_y[None, :]
    nout = Y.shape[0]
    step = 2 * d + 2 if second else d + 2
    Yr = Y.reshape(nout, nboot, step, n)                       # [o, r, block, sample]
    fA, fB, fAi = Yr[:, :, 0], Yr[:, :, 1], Yr[:, :, 2:2 + d]
    nn = L(n)
    both = Yr[:, :, 0:2].reshape(nout, nboot, 2 * n)
    m = both.sum(axis=-1, dtype=L) / (2 * nn)
    Var = ((both - m[..., None]) ** 2).sum(axis=-1, dtype=L) / (2 * nn - 1)           # two-pass, extended precision
    Vi = (fB[:, :, None, :] * (fAi - fA[:, :, None, :])).sum(axis=-1, dtype=L) / nn
    a = fA[:, :, None, :]
    if Ei_estimator == "Jansen1999":
        Ei = ((a - fAi) ** 2).sum(axis=-1, dtype=L) / (2 * nn)
    elif Ei_estimator == "Sobol2007":
        Ei = (a * (a - fAi)).sum(axis=-1, dtype=L) / nn
    elif Ei_estimator == "Homma1996":
        mA = fA.sum(axis=-1, dtype=L) / nn
        Ei = (((fA - mA[..., None]) ** 2).sum(axis=-1, dtype=L) / (nn - 1))[..., None] - (a * fAi).sum(axis=-1, dtype=L) / nn \
            + (mA ** 2)[..., None]
    else:
        s2 = (a ** 2 + fAi ** 2).sum(axis=-1, dtype=L)
        s1 = (a + fAi).sum(axis=-1, dtype=L)
        sp = (a * fAi).sum(axis=-1, dtype=L)
        mh = s1 / (2 * nn)
        # centred form of the same expression (the raw moments cancel when mean^2 >> var, even in extended precision)
        Ei = (s2 / (2 * nn) - mh ** 2) * (1 - (sp / nn - mh ** 2) / (s2 / (2 * nn) - mh ** 2))
    Si, Ti = Vi / Var[..., None], Ei / Var[..., None]          # [o, r, k]
    Sij = None
    if second:
        fBi = Yr[:, :, 2 + d:2 + 2 * d]
        Sij = np.zeros((nout, nboot, d, d), dtype=L)
        fafb = fA * fB
        for k in range(d):
            for j in range(k + 1, d):
                vkj = (fBi[:, :, k] * fAi[:, :, j] - fafb).sum(axis=-1, dtype=L) / nn
                Sij[:, :, k, j] = (vkj - (Vi[:, :, k] + Vi[:, :, j])) / Var
    S1ci = STci = S2ci = None
    if nboot > 1:
        z = L(norm_quantile((1 + conf_level) / 2))

        def mean_ci(x):                                        # x: [o, r, ...]
            mu = x.sum(axis=1, dtype=L) / nboot
            sd = np.sqrt(((x - mu[:, None]) ** 2).sum(axis=1, dtype=L) / (nboot - 1))
            return mu, z * sd / np.sqrt(L(nboot))

        S1, S1ci = mean_ci(Si)
        ST, STci = mean_ci(Ti)
        if second:
            S2, S2ci = mean_ci(Sij)
    else:
        S1, ST = Si[:, 0], Ti[:, 0]
        if second:
            S2 = Sij[:, 0]
    f64 = lambda x: None if x is None else np.asarray(x, dtype=np.float64)
    v1 = lambda x: None if x is None else (f64(x) if multi else f64(x)[0])
    v2 = lambda x: None if x is None else (np.transpose(f64(x), (1, 2, 0)) if multi else f64(x)[0])
    return SobolResult(v1(S1), v1(S1ci), v2(S2) if second else None, v2(S2ci) if second else None, v1(ST), v1(STci))
