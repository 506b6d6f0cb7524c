"""ctypes view of oracle/libgsa_ref.so (the C twin of sobol_oracle.py).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgsa_ref.so")
_TABLE = os.path.join(_HERE, "..", "globalsensitivity.jl_b200", "data", "joe_kuo_21201.bin")

MODELS = {"ishigami": 0, "sobol_g": 1, "linear": 2, "ishi_linear": 3}
ESTIMATORS = {"Homma1996": 0, "Sobol2007": 1, "Jansen1999": 2, "Janon2014": 3}

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("gsa_ref.c", "gsa_ref_streamed.c", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(src) for src in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libgsa_ref.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ref_num_threads.restype = C.c_int
        L.ref_set_num_threads.restype = None
        L.ref_set_num_threads.argtypes = [C.c_int]
        L.ref_norm_quantile.restype = C.c_double
        L.ref_norm_quantile.argtypes = [C.c_double]
        L.ref_all_points_cols.restype = C.c_int64
        L.ref_all_points_cols.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int]
        L.ref_gsa_sobol.restype = C.c_int64
        L.ref_gsa_sobol.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int64, C.c_int, C.c_int,
                                    C.c_int, C.c_double] + [_dp] * 6
        L.ref_analyze_y.restype = None
        L.ref_analyze_y.argtypes = [_dp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double] + [_dp] * 6
        L.ref_sobol_design.restype = None
        L.ref_sobol_design.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, _dp, _dp, C.c_int64, C.c_int64, _dp]
        L.ref_all_points.restype = None
        L.ref_all_points.argtypes = [_dp, _dp, C.c_int, C.c_int64, C.c_int, C.c_int, _dp]
        L.ref_model_batch.restype = None
        L.ref_model_batch.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int64, _dp]
        L.ref_model_nout.restype = C.c_int
        L.ref_model_nout.argtypes = [C.c_int, C.c_int, C.c_int]
        # gsa_ref_streamed.c: the streamed exact-sum oracle
        L.ref_stream_new.restype = C.c_void_p
        L.ref_stream_new.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.ref_stream_free.restype = None
        L.ref_stream_free.argtypes = [C.c_void_p]
        L.ref_stream_push.restype = None
        L.ref_stream_push.argtypes = [C.c_void_p, _dp, _dp, C.c_int64, C.c_int64]
        L.ref_stream_finish.restype = C.c_int64
        L.ref_stream_finish.argtypes = [C.c_void_p, C.c_double] + [_dp] * 6
        L.ref_gsa_sobol_streamed.restype = C.c_int64
        L.ref_gsa_sobol_streamed.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                             C.c_double, C.c_int64] + [_dp] * 6
        L.ref_gsa_sobol_streamed_generated.restype = C.c_int64
        L.ref_gsa_sobol_streamed_generated.argtypes = [C.c_int, _dp, C.c_int, C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int64, C.c_int,
                                                       C.c_int, C.c_int, C.c_double, C.c_int64] + [_dp] * 6
        L.ref_analyze_y_exact.restype = None
        L.ref_analyze_y_exact.argtypes = [_dp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double] + [_dp] * 6
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f(a):
    """(d, N) Julia-shaped array -> Fortran-ordered float64 (d contiguous per sample)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def num_threads() -> int:
    return lib().ref_num_threads()


def use_all_cores() -> int:
    """OpenMP threads = the cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().ref_set_num_threads(n)
    return lib().ref_num_threads()


def norm_quantile(p: float) -> float:
    return lib().ref_norm_quantile(p)


_table = None


def sobol_design(n, lb, ub, num_mats=2, col0=0, ncols=None):
    global _table
    if _table is None:
        _table = np.fromfile(_TABLE, dtype="<u4")
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    d = lb.size
    ncols = n - col0 if ncols is None else ncols
    out = np.empty((num_mats, ncols, d), dtype=np.float64)
    lib().ref_sobol_design(_table.ctypes.data, n, d, num_mats, _p(lb), _p(ub), col0, ncols, _p(out))
    return [out[m].T for m in range(num_mats)]          # each (d, ncols), Fortran-ordered view


def _alloc_result(nout, d, nboot, second):
    multi = nout > 1
    s1 = np.zeros((nout, d), order="F")
    st = np.zeros((nout, d), order="F")
    s1c = np.zeros((nout, d), order="F") if nboot > 1 else None
    stc = np.zeros((nout, d), order="F") if nboot > 1 else None
    s2 = np.zeros((d, d, nout), order="F") if second else None
    s2c = np.zeros((d, d, nout), order="F") if second and nboot > 1 else None

    def fin():
        sq1 = (lambda x: None if x is None else (x if multi else x[0]))
        sq2 = (lambda x: None if x is None else (x if multi else x[:, :, 0]))
        return SimpleNamespace(S1=sq1(s1), S1_Conf_Int=sq1(s1c), S2=sq2(s2), S2_Conf_Int=sq2(s2c),
                               ST=sq1(st), ST_Conf_Int=sq1(stc))
    return (s1, s1c, s2, s2c, st, stc), fin


def gsa_sobol(model, params, A, B, order=(0, 1), nboot=1, conf_level=0.95, Ei_estimator="Jansen1999"):
    A, B = _f(A), _f(B)
    d, N = A.shape
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel(order="F"))
    mid = MODELS[model]
    nout = lib().ref_model_nout(mid, params.size, d)
    second = 2 in order
    bufs, fin = _alloc_result(nout, d, nboot, second)
    rows = lib().ref_gsa_sobol(mid, _p(params), params.size, _p(A), _p(B), d, N, nboot, int(second),
                               ESTIMATORS[Ei_estimator], conf_level, *[_p(b) for b in bufs])
    if rows < 0:
        raise MemoryError("oracle could not materialise all_points")
    res = fin()
    res.rows = rows
    return res


def analyze_y(Y, d, n, order=(0, 1), nboot=1, conf_level=0.95, Ei_estimator="Jansen1999", exact=False):
    """gsa_sobol_all_y_analysis on a given all_y.  exact=True: the same per-sample terms summed EXACTLY
    (gsa_ref_streamed.c) instead of with Julia's pairwise sum."""
    Y = np.asarray(Y, dtype=np.float64)
    nout = 1 if Y.ndim == 1 else Y.shape[0]
    Yf = np.asfortranarray(Y.reshape(nout, -1))
    second = 2 in order
    bufs, fin = _alloc_result(nout, d, nboot, second)
    fn = lib().ref_analyze_y_exact if exact else lib().ref_analyze_y
    fn(_p(Yf), nout, d, n, nboot, int(second), ESTIMATORS[Ei_estimator], conf_level, *[_p(b) for b in bufs])
    res = fin()
    if Y.ndim == 2 and nout == 1:   # a 1 x cols matrix is still "multioutput" in the reference
        res.S1, res.ST = res.S1[None, :], res.ST[None, :]
    return res


def all_points(A, B, nboot=1, second_order=False):
    A, B = _f(A), _f(B)
    d, N = A.shape
    cols = lib().ref_all_points_cols(d, N, nboot, int(second_order))
    P = np.empty((cols, d))
    lib().ref_all_points(_p(A), _p(B), d, N, nboot, int(second_order), _p(P))
    return P.T


def model_batch(model, params, X):
    X = _f(X)
    d, cols = X.shape
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel(order="F"))
    mid = MODELS[model]
    nout = lib().ref_model_nout(mid, params.size, d)
    Y = np.empty((cols, nout))
    lib().ref_model_batch(mid, _p(params), params.size, _p(X), d, cols, _p(Y))
    return Y.T if nout > 1 else Y[:, 0]


# ---------------------------------------------------------------------------------------------------------------
# streamed exact-sum oracle (gsa_ref_streamed.c): full model per swapped coordinate, per-sample terms formed as the
# reference forms them, summed exactly; never holds more than a chunk of the design
# ---------------------------------------------------------------------------------------------------------------
def gsa_sobol_streamed(model, params, A, B, order=(0, 1), nboot=1, conf_level=0.95, Ei_estimator="Jansen1999", chunk=1 << 16):
    A, B = _f(A), _f(B)
    d, N = A.shape
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel(order="F"))
    mid = MODELS[model]
    nout = lib().ref_model_nout(mid, params.size, d)
    second = 2 in order
    bufs, fin = _alloc_result(nout, d, nboot, second)
    rows = lib().ref_gsa_sobol_streamed(mid, _p(params), params.size, _p(A), _p(B), d, N, nboot, int(second),
                                        ESTIMATORS[Ei_estimator], conf_level, chunk, *[_p(b) for b in bufs])
    if rows < 0:
        raise MemoryError("streamed oracle: allocation failed")
    res = fin()
    res.rows = rows
    return res


def gsa_sobol_streamed_generated(model, params, lb, ub, N, order=(0, 1), nboot=1, conf_level=0.95, Ei_estimator="Jansen1999",
                                 p_range_design=False, chunk=1 << 16):
    """The same with the design generated chunk by chunk inside the oracle: generate_design_matrices(N, lb, ub, 2), or, with
    p_range_design=True, the 2*nboot*d-dimensional design of gsa(f, ::Sobol, p_range; samples=N//nboot) (:407-417)."""
    global _table
    if _table is None:
        _table = np.fromfile(_TABLE, dtype="<u4")
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    d = lb.size
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel(order="F"))
    mid = MODELS[model]
    nout = lib().ref_model_nout(mid, params.size, d)
    second = 2 in order
    bufs, fin = _alloc_result(nout, d, nboot, second)
    rows = lib().ref_gsa_sobol_streamed_generated(mid, _p(params), params.size, _table.ctypes.data, _p(lb), _p(ub),
                                                  2 if p_range_design else 1, d, N, nboot, int(second), ESTIMATORS[Ei_estimator],
                                                  conf_level, chunk, *[_p(b) for b in bufs])
    if rows < 0:
        raise MemoryError("streamed oracle: allocation failed")
    res = fin()
    res.rows = rows
    return res


class Stream:
    """Incremental form: push column chunks of the SAME A/B the GPU path sees (e.g. copied back from the device), finish()."""

    def __init__(self, model, params, d, N, order=(0, 1), nboot=1, Ei_estimator="Jansen1999"):
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel(order="F"))
        mid = MODELS[model]
        self.d, self.N, self.nboot, self.second = d, N, nboot, 2 in order
        self.nout = lib().ref_model_nout(mid, self.params.size, d)
        self.h = lib().ref_stream_new(mid, _p(self.params), self.params.size, d, N, nboot, int(self.second), ESTIMATORS[Ei_estimator])
        if not self.h:
            raise MemoryError("ref_stream_new failed")

    def push(self, A_chunk, B_chunk, col0):
        A, B = _f(A_chunk), _f(B_chunk)
        assert A.shape == B.shape and A.shape[0] == self.d
        lib().ref_stream_push(self.h, _p(A), _p(B), col0, A.shape[1])

    def finish(self, conf_level=0.95):
        bufs, fin = _alloc_result(self.nout, self.d, self.nboot, self.second)
        lib().ref_stream_finish(self.h, conf_level, *[_p(b) for b in bufs])
        return fin()

    def close(self):
        if self.h:
            lib().ref_stream_free(self.h)
            self.h = None

    def __del__(self):
        self.close()
