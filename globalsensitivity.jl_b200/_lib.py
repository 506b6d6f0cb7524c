"""ctypes binding of libgsa_cuda.so (the C ABI declared in include/gsa_cuda.h).

The product path has no CPU fallback: if the CUDA library is missing this module raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsa_cuda.so")

GSA_OK = 0
LOC_HOST, LOC_DEVICE = 0, 1
EI_ESTIMATORS = {"Homma1996": 0, "Sobol2007": 1, "Jansen1999": 2, "Janon2014": 3}


class GsaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgsa_cuda error {code}: {msg}")
        self.code = code


class SobolOpts(C.Structure):
    _fields_ = [("second_order", C.c_int), ("nboot", C.c_int), ("conf_level", C.c_double),
                ("ei_estimator", C.c_int), ("input_location", C.c_int)]


_dp = C.POINTER(C.c_double)


class EfastResultC(C.Structure):
    _fields_ = [("S1", C.POINTER(C.c_double)), ("ST", C.POINTER(C.c_double))]


class SobolResultC(C.Structure):
    _fields_ = [("S1", _dp), ("S1_conf_int", _dp), ("S2", _dp), ("S2_conf_int", _dp), ("ST", _dp), ("ST_conf_int", _dp)]


_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
_ip = C.POINTER(C.c_int)
_op = C.POINTER(SobolOpts)
_rp = C.POINTER(SobolResultC)

# name -> argtypes; every entry point returns int.  Must list every symbol of include/gsa_cuda.h.
SIGNATURES = {
    "gsa_version": [],
    "gsa_last_error": [C.c_char_p, _sz],
    "gsa_create": [_i, C.POINTER(_vp)],
    "gsa_destroy": [_vp],
    "gsa_set_stream": [_vp, _vp, _i],
    "gsa_synchronize": [_vp],
    "gsa_host_register": [_vp, _sz],
    "gsa_host_unregister": [_vp],
    "gsa_model_lookup": [C.c_char_p, _ip],
    "gsa_model_nout": [_vp, _i, _i, _i, _ip],
    "gsa_model_register_fatbin": [_vp, _vp, _sz, C.c_char_p, _i, _ip],
    "gsa_model_register_separable": [_vp, _vp, _sz, C.c_char_p, _i, _ip],
    "gsa_sobol_design": [_vp, _i64, _i, _i, _vp, _vp, _i64, _i64, C.POINTER(_vp), _i],
    "gsa_sobol_nstat": [_op, _i, _ip],
    "gsa_sobol_run": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _rp],
    "gsa_sobol_partials": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _i64, _i64, _vp],
    "gsa_sobol_run_sharded": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _i64, _i64, _rp],
    "gsa_sobol_run_sharded_generated": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _i64, _i64, _rp],
    "gsa_sobol_run_generated": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _rp],
    "gsa_sobol_partials_generated": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _i64, _i64, _vp],
    "gsa_sobol_partials_y": [_vp, _op, _vp, _i, _i, _i64, _vp],
    "gsa_sobol_combine": [_vp, _i, _sz, _i, _vp],
    "gsa_sobol_finalize": [_op, _vp, _i, _i, _i64, _rp],
    "gsa_sobol_analyze_y": [_vp, _op, _vp, _i, _i, _i64, _rp],
    "gsa_sobol_all_points": [_vp, _op, _vp, _vp, _i, _i64, _vp, _i],
    "gsa_multi_create": [C.POINTER(C.c_int), _i, C.POINTER(_vp)],
    "gsa_multi_destroy": [_vp],
    "gsa_multi_device_count": [_vp, _ip],
    "gsa_multi_handle": [_vp, _i, C.POINTER(_vp)],
    "gsa_multi_model_register_fatbin": [_vp, _vp, _sz, C.c_char_p, _i, _ip],
    "gsa_multi_model_register_separable": [_vp, _vp, _sz, C.c_char_p, _i, _ip],
    "gsa_multi_sobol_run": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _rp],
    "gsa_multi_sobol_run_generated": [_vp, _op, _i, _vp, _i, _vp, _vp, _i, _i64, _rp],
    "gsa_efast_design": [_vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _i],
    "gsa_efast_run": [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i64, C.POINTER(EfastResultC)],
    "gsa_efast_analyze_y": [_vp, _i, _vp, _i, _i, _i, _i64, C.POINTER(EfastResultC)],
    "gsa_last_timing": [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _ip],
    "gsa_peer_export": [_vp, _sz, _vp],
    "gsa_peer_connect": [_vp, _i, _i, _vp],
    "gsa_peer_allreduce": [_vp, _vp, _sz, _i, _vp],
    "gsa_peer_status": [_vp, _ip],
    "gsa_peer_unmap": [_vp],
    "gsa_peer_disconnect": [_vp],
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python globalsensitivity.jl_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback for the Sobol path.")
        L = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library does not export it
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().gsa_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != GSA_OK:
        raise GsaError(rc, last_error())
