"""Cost of the fused tail (csrc/tail.cuh) against the classic launch sequence, per kernel family: kernel / device / wall time of the
public call with GSA_NO_TAIL unset and set.  One JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
h = g.default_handle()
CASES = {
    "c1": dict(model=("ishigami", [7.0, 0.1]), d=3, N=600000, order=[0, 1], nboot=1, lb=-np.pi, ub=np.pi),
    "c2": dict(model=("ishigami", [7.0, 0.1]), d=3, N=1 << 20, order=[0, 1, 2], nboot=1000, lb=-np.pi, ub=np.pi),
    "c3": dict(model=("sobol_g", [0, 1, 4.5, 9, 99, 99, 99, 99]), d=8, N=1 << 22, order=[0, 1, 2], nboot=1, lb=0.0, ub=1.0),
    "c5s": dict(model=("sobol_g", [0, 1, 4.5, 9] + [99.0] * 96), d=100, N=1 << 20, order=[0, 1], nboot=1, lb=0.0, ub=1.0),
    "c5b": dict(model=("sobol_g", [0, 1, 4.5, 9] + [99.0] * 96), d=100, N=1 << 20, order=[0, 1], nboot=100, lb=0.0, ub=1.0),
    "c5_8gpu_shard": dict(model=("sobol_g", [0, 1, 4.5, 9] + [99.0] * 96), d=100, N=1 << 23, order=[0, 1], nboot=1, lb=0.0, ub=1.0),
}
for name, c in CASES.items():
    d, N = c["d"], c["N"]
    model, method = g.DeviceModel(*c["model"]), g.Sobol(order=c["order"], nboot=c["nboot"])
    A, B = g.generate_design_matrices(N, np.full(d, c["lb"]), np.full(d, c["ub"]), 2, device=True, handle=h)
    torch.cuda.synchronize()
    out = {"case": name}
    for mode in ("tail", "classic") + tuple(f"dbg{b}" for b in (1, 2, 4, 7) if "--dbg" in sys.argv) + (("tail_tpc1", "classic_tpc1") if "--tpc" in sys.argv else ()):
        os.environ.pop("GSA_NO_TAIL", None)
        os.environ.pop("GSA_TAIL_DBG", None)
        os.environ.pop("GSA_TILES_PER_CTA", None)
        if mode.endswith("_tpc1"):
            os.environ["GSA_TILES_PER_CTA"] = "1"
            if mode.startswith("classic"):
                os.environ["GSA_NO_TAIL"] = "1"
        if mode == "classic":
            os.environ["GSA_NO_TAIL"] = "1"
        elif mode.startswith("dbg"):
            os.environ["GSA_TAIL_DBG"] = mode[3:]
        ks, ts, ws = [], [], []
        for i in range(12):
            t0 = time.perf_counter()
            res = g.gsa(model, method, A, B, batch=True, handle=h)
            w = time.perf_counter() - t0
            t, k, nl = h.last_timing()
            if i >= 2:
                ks.append(k); ts.append(t); ws.append(w)
        out[mode] = {"kernel_us": round(1e3 * float(np.median(ks)), 1), "device_us": round(1e3 * float(np.median(ts)), 1),
                     "wall_us": round(1e6 * float(np.median(ws)), 1), "launches": nl}
    os.environ.pop("GSA_NO_TAIL", None)
    os.environ.pop("GSA_TAIL_DBG", None)
    print(json.dumps(out), flush=True)
    del A, B
