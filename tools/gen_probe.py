"""Times the fused-generation path (no A/B in memory): python tools/gen_probe.py log2n [d]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
N = 1 << log2n
a = ([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d]
model, method = g.DeviceModel("sobol_g", a), g.Sobol()
h = g.default_handle()
lb, ub = np.zeros(d), np.ones(d)
ks = []
for _ in range(5):
    p = g.sobol_partials_generated(model, method, lb, ub, N, 0, N, handle=h)
    torch.cuda.synchronize()
    ks.append(h.last_timing()[1])
k = float(np.median(ks))
res = g.sobol_finalize(method, p, d, N)
print(f"GEN d={d} N=2^{log2n} T={os.environ.get('GSA_PRODUCT_T','auto')}: {k:.3f} ms  {N*(d+2)/k/1e6:.1f} Grows/s  (HBM-read equivalent {16.0*d*N/k/1e6:.0f} GB/s)  S1[0]={res.S1[0]:.6f}")
