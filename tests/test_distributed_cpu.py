"""World-size-2 (gloo, CPU) test of the multi-rank host logic: contiguous column shards, ONE all-reduce of the additive
per-replicate partial sums, host epilogue.  The per-shard sums come from the test-side NumPy construction
(tests/stat_records.py), since there is no GPU here; on the GPU box the same plumbing is fed by the fused kernel
(tests/test_gpu_parity.py::test_sharded_partials_are_additive)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sobol_oracle as so


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _offset_model(X):
    """mean^2 / var ~ 1e16: the (hi, lo) pairs of the record matter"""
    return 1.0e5 + 1.0e-3 * so.ishi_linear_batch(X)


def _worker(rank, world, port, nboot, est, out_q, offset=False):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gsa_loader
    import stat_records as sr
    g = gsa_loader.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, N = 4, 3001
    A, B = so.generate_design_matrices(N, -np.pi * np.ones(d), np.pi * np.ones(d))
    col0, ncols = g.shard_columns(N, rank, world)
    P = sr.shard_records(_offset_model if offset else so.ishi_linear_batch, A, B, nboot, True, est, col0, ncols)     # this rank's shard only
    res = g.reduce_and_finalize(g.Sobol(order=[0, 1, 2], nboot=nboot), torch.from_numpy(P), d, N, Ei_estimator=est)
    out_q.put((rank, {f: getattr(res, f) for f in ("S1", "ST", "S2", "S1_Conf_Int", "ST_Conf_Int", "S2_Conf_Int")}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nboot,est", [(1, "Jansen1999"), (3, "Jansen1999"), (3, "Janon2014"), (2, "Homma1996")])
def test_two_rank_allreduce_matches_oracle(nboot, est):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nboot, est, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, N = 4, 3001
    A, B = so.generate_design_matrices(N, -np.pi * np.ones(d), np.pi * np.ones(d))
    ref = so.gsa_sobol(so.ishi_linear_batch, A, B, order=(0, 1, 2), nboot=nboot, Ei_estimator=est)
    for rank in range(world):
        for f, v in got[rank].items():
            r = getattr(ref, f)
            assert (v is None) == (r is None)
            if v is not None:
                assert np.max(np.abs(v - r) / np.maximum(1.0, np.abs(r))) <= 1e-12, (rank, f)     # SURVEY 8 tolerance
    for f in got[0]:
        if got[0][f] is not None:
            assert np.array_equal(got[0][f], got[1][f])          # every rank returns identical results


def test_two_rank_reduce_keeps_double_double_pairs():
    """mean^2 >> var across ranks: the all-gather + ordered double-double combine must reproduce the two-pass
    variance of the whole design; adding hi and lo words independently (a plain all_reduce) would not."""
    world, port, nboot, est = 2, _free_port(), 2, "Jansen1999"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nboot, est, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, N = 4, 3001
    A, B = so.generate_design_matrices(N, -np.pi * np.ones(d), np.pi * np.ones(d))
    Y = _offset_model(so.all_points(A, B, nboot, True))             # the fp64 model values the shards saw; their analysis is exact below
    ex = so.gsa_sobol_exact(None, None, None, order=(0, 1, 2), nboot=nboot, Ei_estimator=est, all_y=Y, d=d, n=N // nboot)
    for rank in range(world):
        for f in ("ST", "ST_Conf_Int"):                        # ST (Jansen) isolates Var; S1 is noise in the reference's own formula here
            v, r = got[rank][f], getattr(ex, f)
            assert np.max(np.abs(v - r) / np.maximum(1.0, np.abs(r))) <= 1e-12, (rank, f)
