"""Tuning sweep of the d = 20 second-order kernels (run on a B200): one subprocess per setting of the knobs
GSA_PAIRS_YCH / GSA_PAIRS_YST (Y source) and GSA_PAIRS_VARIANT (product source); prints kernel time and HBM fraction."""
import json, os, subprocess, sys

CHILD = r'''
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import gsa_loader
g = gsa_loader.load()
h = g.default_handle(0)
torch.manual_seed(0)
which, d, n = sys.argv[1], 20, 1 << 21
peak = 6533.5
if which == "y":
    Y = torch.rand((2 * d + 2) * n, dtype=torch.float64, device="cuda")
    ks = []
    for i in range(8):
        r = g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2]), Y, d, n, handle=h)
        if i >= 2: ks.append(h.last_timing()[1])
    alg = 8.0 * (2 * d + 2) * n
else:
    if which == "c3": d, n = 8, 1 << 22
    A, B = g.generate_design_matrices(n, np.zeros(d), np.ones(d), 2, device=True, handle=h)
    m = g.DeviceModel("sobol_g", list(np.linspace(0, 9, d)))
    ks = []
    for i in range(8):
        r = g.gsa(m, g.Sobol(order=[0, 1, 2]), A, B, batch=True, handle=h)
        if i >= 2: ks.append(h.last_timing()[1])
    alg = 16.0 * d * n
k = float(np.median(ks))
print(json.dumps({"kernel_ms": k, "gbs": alg / k / 1e6, "frac": alg / k / 1e6 / peak, "S1_0": float(np.ravel(r.S1)[0]), "S2_01": float(np.asarray(r.S2)[0, 1])}))
'''

def run(which, env):
    e = dict(os.environ, **env)
    p = subprocess.run([sys.executable, "-c", CHILD, which], env=e, capture_output=True, text=True, timeout=300)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "FAILED " + p.stderr[-300:]
    print(which, env, line, flush=True)

if __name__ == "__main__":
    for ch, st in ((16, 3), (16, 4), (16, 6), (32, 2), (32, 3)):
        run("y", {"GSA_PAIRS_YCH": str(ch), "GSA_PAIRS_YST": str(st)})
    run("c3", {})
    run("p", {})
    for v in ("4,2", "1,6", "2,3", "4,3", "2,6", "1,8"):
        run("p", {"GSA_PAIRS_VARIANT": v})
