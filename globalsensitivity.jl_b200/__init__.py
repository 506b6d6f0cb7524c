"""B200-native Sobol sensitivity path of GlobalSensitivity.jl (host-side mirror of the reference API).

    from gsa_b200 import gsa, Sobol, DeviceModel            (see gsa_loader.py for the import name)
    res = gsa(DeviceModel("ishigami", [7.0, 0.1]), Sobol(order=[0, 1, 2], nboot=20), A, B, batch=True)
    res.S1, res.ST, res.S2, res.S1_Conf_Int, ...

All numerics run in libgsa_cuda.so (hand-written sm_100a CUDA behind a C ABI, include/gsa_cuda.h).
"""
from .sobol import (DeviceModel, GsaHandle, GsaMulti, PeerGroup, Sobol, SobolResult, default_handle, generate_design_matrices, sample, gsa,
                    gsa_sobol_all_y_analysis, nstat, sobol_partials, sobol_partials_generated, sobol_finalize, shard_columns,
                    reduce_and_finalize, gsa_sharded, ShardedStep, pin_host, unpin_host)
from . import sobol
from .efast import eFAST, eFASTResult, efast_design, gsa_efast, gsa_efast_all_y_analysis
from ._lib import GsaError, LIB_PATH, lib

__all__ = ["gsa", "Sobol", "SobolResult", "DeviceModel", "GsaHandle", "GsaMulti", "PeerGroup", "generate_design_matrices", "sample",
           "gsa_sobol_all_y_analysis", "sobol_partials", "sobol_partials_generated", "sobol_finalize", "shard_columns", "nstat", "reduce_and_finalize", "gsa_sharded", "ShardedStep", "pin_host", "unpin_host", "default_handle",
           "eFAST", "eFASTResult", "efast_design", "gsa_efast", "gsa_efast_all_y_analysis", "GsaError", "LIB_PATH", "lib"]
