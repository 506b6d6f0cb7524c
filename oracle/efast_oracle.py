"""CPU oracle (NumPy) for the eFAST path of GlobalSensitivity.jl  --  TEST INFRASTRUCTURE ONLY (same rules as sobol_oracle.py).

Restates  gsa(f, eFAST(num_harmonics), p_range; samples, batch=true)  (reference src/eFAST_sensitivity.jl:75-166) and
gsa_efast_all_y_analysis (:167-213).  Third-party arithmetic absent from /root/reference: FFTW.jl's `rfft` (Project.toml
compat "1.3"): restated with numpy.fft.rfft (pocketfft) -- the DFT is a defined function of its input, the two libraries agree
to a few ulp of the largest bin.

Parity status: PINNED TO 1e-4 ONLY.  The reference draws the phase of every search curve from `rand(rng)` (:111), so its golden
vectors (test/eFAST_method.jl:34-38, :66-70, :98-113) hold for ANY phase to the tests' own atol = 1e-4; tests/test_efast.py checks
this restatement against those literals for several phases.  Here `phi` is an explicit input, so the CUDA path and this oracle can
be compared on identical inputs at 1e-12.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np


@dataclass
class eFASTResult:
    """`struct eFASTResult` (:70-73): S1, ST are 1 x d (scalar model) or nout x d matrices."""
    S1: np.ndarray
    ST: np.ndarray


def efast_frequencies(num_params: int, samples: int, num_harmonics: int = 4) -> np.ndarray:
    """omega (:83-96): omega[0] for the parameter of interest, the complementary set for the others."""
    omega = [(samples - 1) // (2 * num_harmonics)]
    m = omega[0] // (2 * num_harmonics)
    if m >= num_params - 1:
        # floor.(range(1, stop = m, length = d - 1)) in exact integer arithmetic: 1 + k (m - 1) / (d - 2)
        omega += [1 + (k * (m - 1)) // (num_params - 2) if num_params > 2 else 1 for k in range(num_params - 1)]
    else:
        omega += [(k % m) + 1 for k in range(num_params - 1)]
    return np.array(omega, dtype=np.int64)


def efast_design(p_range, samples: int, num_harmonics: int = 4, phi: Optional[Sequence[float]] = None) -> np.ndarray:
    """ps (:98-135): d x (samples*d); block i holds the search curve in which parameter i carries omega[0].
    `phi[i]` replaces the reference's `2rand(rng)` (:111)."""
    d = len(p_range)
    omega = efast_frequencies(d, samples, num_harmonics)
    s = (2.0 / samples) * np.arange(samples)
    phi = np.zeros(d) if phi is None else np.asarray(phi, dtype=np.float64)
    ps = np.empty((d, samples * d))
    for i in range(d):
        ot = np.empty(d, dtype=np.int64)
        ot[i] = omega[0]
        ot[:i] = omega[1:i + 1]
        ot[i + 1:] = omega[i + 1:]
        for j in range(d):
            lo, hi = float(p_range[j][0]), float(p_range[j][1])
            blk = slice(i * samples, (i + 1) * samples)
            if lo == hi:
                ps[j, blk] = lo
            else:
                arg = ot[j] * s + phi[i]
                q = 0.5 + (1.0 / np.pi) * np.arcsin(sinpi(arg))
                ps[j, blk] = lo + q * (hi - lo)          # quantile(Uniform(lo, hi), q) = lo + q (hi - lo)
    return ps


def sinpi(x: np.ndarray) -> np.ndarray:
    """Julia's sinpi: exact argument reduction (x mod 2) before the multiplication by pi."""
    r = np.mod(x, 2.0)
    r = np.where(r > 1.0, r - 2.0, r)                   # (-1, 1]
    r = np.where(r > 0.5, 1.0 - r, np.where(r < -0.5, -1.0 - r, r))
    return np.sin(np.pi * r)


def efast_analysis(all_y: np.ndarray, num_params: int, samples: int, num_harmonics: int = 4) -> eFASTResult:
    """gsa_efast_all_y_analysis (:167-213): all_y is a vector of length samples*d (scalar model) or nout x samples*d."""
    Y = np.asarray(all_y, dtype=np.float64)
    Y = Y[None, :] if Y.ndim == 1 else Y
    nout = Y.shape[0]
    omega0 = (samples - 1) // (2 * num_harmonics)
    S1 = np.empty((nout, num_params))
    ST = np.empty((nout, num_params))
    for i in range(num_params):
        ft = np.fft.rfft(Y[:, i * samples:(i + 1) * samples], axis=1)
        buf = np.abs(ft[:, 1:samples // 2]) ** 2                              # |F_j|^2, j = 1 .. samples/2 - 1   (:186, :200)
        z = buf.sum(axis=1)
        S1[:, i] = buf[:, np.arange(1, num_harmonics + 1) * omega0 - 1].sum(axis=1) / z    # bins h*omega0       (:188, :202)
        ST[:, i] = 1.0 - buf[:, :omega0 // 2].sum(axis=1) / z                               # bins 1..omega0/2    (:189, :203)
    return eFASTResult(S1, ST)


def gsa_efast(f: Callable[[np.ndarray], np.ndarray], p_range, samples: int, num_harmonics: int = 4,
              phi: Optional[Sequence[float]] = None) -> eFASTResult:
    ps = efast_design(p_range, samples, num_harmonics, phi)
    return efast_analysis(f(ps), len(p_range), samples, num_harmonics)
