"""Host-side mirror of the reference's eFAST interface (src/eFAST_sensitivity.jl) over libgsa_cuda.so.

Reference API                                                    here
-------------------------------------------------------------------------------------------------------------
eFAST(; num_harmonics = 4)                              :63-67   eFAST(num_harmonics=4)
eFASTResult (S1, ST)                                    :70-73   eFASTResult (same fields; 1 x d or nout x d)
gsa(f, ::eFAST, p_range; samples, batch, rng)           :75-166  gsa_efast(f, eFAST(...), p_range, samples=..., batch=..., phi=...)
gsa_efast_all_y_analysis(method, all_y, d, ...)         :167-213 gsa_efast_all_y_analysis(method, all_y, d, samples)

`f` is a DeviceModel (design generated in registers, model evaluated and spectrum analysed on the GPU: nothing but the ranges goes
in) or any Python callable (batch=True: it receives the d x samples*d design built on the GPU; the spectral analysis runs on the GPU).
`phi` (d phases in [0, 2)) replaces the reference's `2 rand(rng)` per search curve (:111); default: drawn from `rng`
(numpy.random.default_rng() if None), so results vary from call to call exactly as the reference's do.  Only [lb, ub] ranges are
supported (Distributions.jl objects stay on the reference's own code)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from ._lib import EfastResultC, LOC_DEVICE, LOC_HOST, check, lib
from .sobol import DeviceModel, GsaHandle, _TorchStream, _is_torch, default_handle


@dataclass
class eFAST:
    """`struct eFAST <: GSAMethod` (reference :63-67)."""
    num_harmonics: int = 4


@dataclass
class eFASTResult:
    """`struct eFASTResult` (reference :70-73)."""
    S1: Any
    ST: Any


def _ranges(p_range):
    lb = np.ascontiguousarray([float(p[0]) for p in p_range], dtype=np.float64)
    ub = np.ascontiguousarray([float(p[1]) for p in p_range], dtype=np.float64)
    return lb, ub


def _phases(d, phi, rng):
    if phi is None:
        rng = rng if rng is not None else np.random.default_rng()
        phi = 2.0 * rng.random(d)                                        # :111
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    if phi.shape != (d,):
        raise ValueError("phi must have one phase per parameter")
    return phi


def _result(nout, d):
    s1 = np.zeros((nout, d), order="F")
    st = np.zeros((nout, d), order="F")
    dp = C.POINTER(C.c_double)
    return EfastResultC(s1.ctypes.data_as(dp), st.ctypes.data_as(dp)), s1, st


def efast_design(method: eFAST, p_range, samples: int, *, phi=None, rng=None, device: bool = False, handle: Optional[GsaHandle] = None):
    """The d x (samples*d) search-curve design (:98-135), built on the GPU."""
    h = handle or default_handle()
    lb, ub = _ranges(p_range)
    d = lb.size
    phi = _phases(d, phi, rng)
    if device:
        import torch
        P = torch.empty((samples * d, d), dtype=torch.float64, device="cuda")
        with _TorchStream(h, True):
            check(lib().gsa_efast_design(h.ptr, lb.ctypes.data, ub.ctypes.data, d, samples, method.num_harmonics, phi.ctypes.data, P.data_ptr(), LOC_DEVICE))
        return P.t()
    P = np.empty((d, samples * d), dtype=np.float64, order="F")
    check(lib().gsa_efast_design(h.ptr, lb.ctypes.data, ub.ctypes.data, d, samples, method.num_harmonics, phi.ctypes.data, P.ctypes.data, LOC_HOST))
    return P


def gsa_efast_all_y_analysis(method: eFAST, all_y, num_params: int, samples: int, *, handle: Optional[GsaHandle] = None) -> eFASTResult:
    """gsa_efast_all_y_analysis (:167-213) on the GPU: all_y is a vector of samples*d values or an nout x samples*d matrix."""
    h = handle or default_handle()
    if _is_torch(all_y):
        import torch
        t = all_y.detach().to(torch.float64)
        multi = t.dim() == 2
        tt = t.t().contiguous() if multi else t.contiguous()
        nout, cols = (int(t.shape[0]) if multi else 1), int(t.shape[-1])
        ptr, loc, keep = tt.data_ptr(), (LOC_DEVICE if tt.is_cuda else LOC_HOST), tt
    else:
        a = np.asarray(all_y, dtype=np.float64)
        multi = a.ndim == 2
        a = np.asfortranarray(a) if multi else np.ascontiguousarray(a)
        nout, cols = (a.shape[0] if multi else 1), a.shape[-1]
        ptr, loc, keep = a.ctypes.data, LOC_HOST, a
    if cols != samples * num_params:
        raise ValueError(f"all_y has {cols} columns, expected samples*d = {samples * num_params}")
    rc, s1, st = _result(nout, num_params)
    with _TorchStream(h, loc == LOC_DEVICE):
        check(lib().gsa_efast_analyze_y(h.ptr, method.num_harmonics, ptr, loc, nout, num_params, samples, C.byref(rc)))
    del keep
    return eFASTResult(s1, st)


def gsa_efast(f, method: eFAST, p_range, *, samples: int, batch: bool = False, phi=None, rng=None,
              handle: Optional[GsaHandle] = None) -> eFASTResult:
    """gsa(f, method::eFAST, p_range; samples, batch) (reference :75-166)."""
    if not isinstance(method, eFAST):
        raise TypeError("method must be eFAST")
    h = handle or getattr(f, "handle", None) or default_handle()
    lb, ub = _ranges(p_range)
    d = lb.size
    phi = _phases(d, phi, rng)
    if isinstance(f, DeviceModel):
        nout = f.nout(d)
        rc, s1, st = _result(nout, d)
        check(lib().gsa_efast_run(h.ptr, method.num_harmonics, f.model_id, f.params.ctypes.data, f.params.size, lb.ctypes.data,
                                  ub.ctypes.data, phi.ctypes.data, d, samples, C.byref(rc)))
        return eFASTResult(s1, st)
    if not callable(f):
        raise TypeError("f must be a DeviceModel or a callable")
    P = efast_design(method, p_range, samples, phi=phi, handle=h)
    if batch:
        all_y = f(P)                                                                                     # :137
    else:
        ys = [np.asarray(f(P[:, c]), dtype=np.float64) for c in range(P.shape[1])]                       # :144
        all_y = np.stack(ys, axis=-1) if ys[0].ndim else np.asarray(ys)
    return gsa_efast_all_y_analysis(method, all_y, d, samples, handle=h)
