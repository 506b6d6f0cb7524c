"""Times the estimator-on-Y kernel (K3) on a random Y: python tools/ky_probe.py d log2n [second]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, json
import gsa_loader
g = gsa_loader.load()
d, log2n = int(sys.argv[1]), int(sys.argv[2])
second = len(sys.argv) > 3 and sys.argv[3] == "1"
n = 1 << log2n
step = 2 * d + 2 if second else d + 2
h = g.default_handle()
Y = torch.rand(n * step, dtype=torch.float64, device="cuda")
method = g.Sobol(order=[0, 1, 2] if second else [0, 1])
ks = []
for _ in range(5):
    g.gsa_sobol_all_y_analysis(method, Y, d, n, handle=h)
    ks.append(h.last_timing()[1])
k = float(np.median(ks))
print(f"K3 d={d} n=2^{log2n} second={second} TS={os.environ.get('GSA_KY_TS','-')} thr={os.environ.get('GSA_KY_THREADS','-')}: {k:.3f} ms  {8.0*n*step/k/1e6:.0f} GB/s")
