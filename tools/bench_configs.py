"""Times every BASELINE.json config (and the estimator-only kernel on the same shapes) with device-resident inputs.
    python tools/bench_configs.py [--only c3] [--reps 5]
Prints one JSON line per measurement: kernel time (CUDA events around the dominant kernel, from gsa_last_timing),
algorithmic bytes, achieved GB/s and the fraction of the measured HBM peak."""
import argparse, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def W_c4(nout=256, d=20):
    o, i = np.meshgrid(np.arange(1, nout + 1), np.arange(1, d + 1), indexing="ij")
    return 1.0 + ((o * 31 + i * 17) % 13) / 13.0


CONFIGS = {
    "c1": dict(model=("ishigami", [7.0, 0.1]), d=3, N=600000, order=[0, 1], nboot=1, lb=-np.pi, ub=np.pi),
    "c2": dict(model=("ishigami", [7.0, 0.1]), d=3, N=1 << 20, order=[0, 1, 2], nboot=1000, lb=-np.pi, ub=np.pi),
    "c3": dict(model=("sobol_g", [0, 1, 4.5, 9, 99, 99, 99, 99]), d=8, N=1 << 22, order=[0, 1, 2], nboot=1, lb=0.0, ub=1.0),
    "c4": dict(model=("linear", W_c4()), d=20, N=1 << 22, order=[0, 1], nboot=1, lb=0.0, ub=1.0),
    "p20": dict(model=("sobol_g", [0, 1, 4.5, 9] + [99.0] * 16), d=20, N=1 << 21, order=[0, 1, 2], nboot=1, lb=0.0, ub=1.0),   # second order, d > 8
    # the replicate ("bootstrap") path at sizes that are not launch-latency-bound
    "b5": dict(model=("sobol_g", [0, 1, 4.5, 9] + [99.0] * 96), d=100, N=1 << 24, order=[0, 1], nboot=100, lb=0.0, ub=1.0),
    "b2": dict(model=("ishigami", [7.0, 0.1]), d=3, N=1 << 25, order=[0, 1, 2], nboot=1000, lb=-np.pi, ub=np.pi),
    "c5": dict(model=("sobol_g", [0, 1, 4.5, 9] + [99.0] * 96), d=100, N=1 << 24, order=[0, 1], nboot=1, lb=0.0, ub=1.0),
}


def torch_model(name, params, X):  # X: (d, cols) cuda tensor -> y (cols,)   (used only to synthesise Y for the K3 timing)
    if name == "ishigami":
        return torch.sin(X[0]) + params[0] * torch.sin(X[1]) ** 2 + params[1] * X[2] ** 4 * torch.sin(X[0])
    if name == "sobol_g":
        y = torch.ones(X.shape[1], dtype=torch.float64, device=X.device)
        for i in range(X.shape[0]):
            y *= (torch.abs(4 * X[i] - 2) + params[i]) / (1 + params[i])
        return y
    raise ValueError(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-y", action="store_true")
    args = ap.parse_args()
    h = g.default_handle()
    for name, c in CONFIGS.items():
        if args.only and name not in args.only.split(","):
            continue
        d, N = c["d"], c["N"]
        model = g.DeviceModel(c["model"][0], c["model"][1])
        method = g.Sobol(order=c["order"], nboot=c["nboot"])
        second = 2 in c["order"]
        step = 2 * d + 2 if second else d + 2
        n = N // c["nboot"]
        rows = c["nboot"] * n * step
        A, B = g.generate_design_matrices(N, np.full(d, c["lb"]), np.full(d, c["ub"]), 2, device=True, handle=h)
        try:
            ks, ts = [], []
            for _ in range(args.reps):
                p = g.sobol_partials(model, method, A, B, N, 0, handle=h)
                torch.cuda.synchronize()
                t, k, nl = h.last_timing()
                ks.append(k); ts.append(t)
            k, t = float(np.median(ks)), float(np.median(ts))
            res = g.sobol_finalize(method, p, d, N)
            bytes_ = 16.0 * d * c["nboot"] * n
            print(json.dumps({"config": name, "path": "fused", "model": c["model"][0], "d": d, "N": N, "nboot": c["nboot"], "second": second,
                              "rows": rows, "kernel_ms": k, "call_ms": t, "launches": nl, "rows_per_s_kernel": rows / (k * 1e-3),
                              "rows_per_s_call": rows / (t * 1e-3), "alg_bytes": bytes_, "GBps": bytes_ / (k * 1e-3) / 1e9,
                              "frac_hbm": bytes_ / (k * 1e-3) / 1e9 / PEAK, "S1_head": np.ravel(res.S1)[:3].tolist()}), flush=True)
        except Exception as e:  # noqa
            print(json.dumps({"config": name, "path": "fused", "error": str(e)}), flush=True)
        if args.no_y or c["model"][0] == "linear" or rows * 8 > 60e9:
            del A, B
            continue
        # estimator-only kernel on a Y of the same shape (random values: only the timing matters here; parity is in tests/)
        del A, B
        torch.cuda.empty_cache()
        Y = torch.rand(rows, dtype=torch.float64, device="cuda")
        ks, ts = [], []
        for _ in range(args.reps):
            res = g.gsa_sobol_all_y_analysis(method, Y, d, n, handle=h)
            t, k, nl = h.last_timing()
            ks.append(k); ts.append(t)
        k, t = float(np.median(ks)), float(np.median(ts))
        bytes_ = 8.0 * rows
        print(json.dumps({"config": name, "path": "estimator_y", "d": d, "N": N, "nboot": c["nboot"], "second": second, "rows": rows,
                          "kernel_ms": k, "call_ms": t, "launches": nl, "rows_per_s_kernel": rows / (k * 1e-3), "alg_bytes": bytes_,
                          "GBps": bytes_ / (k * 1e-3) / 1e9, "frac_hbm": bytes_ / (k * 1e-3) / 1e9 / PEAK,
                          "S1_head": np.ravel(res.S1)[:3].tolist()}), flush=True)
        del Y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
