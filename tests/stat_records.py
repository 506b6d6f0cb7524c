"""Test-side NumPy construction of the library's sufficient-statistics records (layout documented in
globalsensitivity.jl_b200/csrc/gsa_common.cuh) from model outputs.  Used to test the host-side finalize and the
multi-rank all-reduce plumbing on CPU without a GPU."""
import math

import numpy as np


def _dd(values):
    hi = math.fsum(values)
    lo = math.fsum(list(values) + [-hi])
    return hi, lo


def nstat(d, est, second):
    ne = 3 if est == "Janon2014" else 1
    return 8 + d + ne * d + (d * (d - 1) // 2 if second else 0)


def records(yA, yB, yAB, yBA, est="Jansen1999"):
    """yA, yB: (n,), yAB: (d, n), yBA: (d, n) or None  ->  one record (nstat,)."""
    d = yAB.shape[0]
    second = yBA is not None
    rec = np.zeros(nstat(d, est, second))
    both = np.concatenate([yA, yB])
    rec[0], rec[1] = _dd(both.tolist())
    rec[2], rec[3] = _dd([float(v) ** 2 for v in both.tolist()]) if False else _dd_sq(both)
    rec[4], rec[5] = _dd(yA.tolist())
    rec[6], rec[7] = _dd_sq(yA)
    for k in range(d):
        rec[8 + k] = math.fsum((yB * (yAB[k] - yA)).tolist())
        if est == "Jansen1999":
            rec[8 + d + k] = math.fsum(((yA - yAB[k]) ** 2).tolist())
        elif est == "Sobol2007":
            rec[8 + d + k] = math.fsum((yA * (yA - yAB[k])).tolist())
        else:
            rec[8 + d + k] = math.fsum((yA * yAB[k]).tolist())
            if est == "Janon2014":
                rec[8 + 2 * d + k] = math.fsum((yA ** 2 + yAB[k] ** 2).tolist())
                rec[8 + 3 * d + k] = math.fsum((yA + yAB[k]).tolist())
    if second:
        ne = 3 if est == "Janon2014" else 1
        p = 8 + d + ne * d
        for k in range(d):
            for j in range(k + 1, d):
                rec[p] = math.fsum((yBA[k] * yAB[j] - yA * yB).tolist())
                p += 1
    return rec


def _dd_sq(v):
    """double-double sum of squares, exact products via Fractions-free splitting (Dekker through fsum)."""
    from fractions import Fraction
    s = sum((Fraction(float(x)) ** 2 for x in v.tolist()), Fraction(0))
    hi = float(s)
    lo = float(s - Fraction(hi))
    return hi, lo


def shard_records(model_batch, A, B, nboot, second, est, col0, ncols):
    """Records (nboot, nout, nstat) of the column shard [col0, col0+ncols) of the d x N design (zeros elsewhere)."""
    d, N = A.shape
    n = N // nboot
    probe = np.asarray(model_batch(A[:, :1]))
    nout = 1 if probe.ndim == 1 else probe.shape[0]
    out = np.zeros((nboot, nout, nstat(d, est, second)))
    for r in range(nboot):
        lo, hi = max(r * n, col0), min((r + 1) * n, col0 + ncols)
        if hi <= lo:
            continue
        a, b = A[:, lo:hi], B[:, lo:hi]
        ev = lambda X: np.asarray(model_batch(X)).reshape(nout, -1)
        yA, yB = ev(a), ev(b)
        yAB, yBA = [], []
        for k in range(d):
            x = a.copy(); x[k] = b[k]; yAB.append(ev(x))
            if second:
                x = b.copy(); x[k] = a[k]; yBA.append(ev(x))
        for o in range(nout):
            out[r, o] = records(yA[o], yB[o], np.stack([y[o] for y in yAB]),
                                np.stack([y[o] for y in yBA]) if second else None, est)
    return out
