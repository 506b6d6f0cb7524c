"""End-to-end rows/s of gsa(model, Sobol(), A_host, B_host) for pageable vs pinned host designs: python tools/e2e_probe.py log2n"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
d, N = 100, 1 << log2n
a = [0.0, 1.0, 4.5, 9.0] + [99.0] * 96
model, method = g.DeviceModel("sobol_g", a), g.Sobol()
A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True)
An, Bn = A.t().contiguous().cpu().numpy().T, B.t().contiguous().cpu().numpy().T     # pageable, Julia layout
Ap, Bp = A.t().contiguous().cpu().pin_memory().t(), B.t().contiguous().cpu().pin_memory().t()
del A, B
for name, (X, Y) in (("pageable", (An, Bn)), ("pinned", (Ap, Bp))):
    g.gsa(model, method, X, Y, batch=True)
    t0 = time.perf_counter()
    for _ in range(3):
        res = g.gsa(model, method, X, Y, batch=True)
    dt = (time.perf_counter() - t0) / 3
    print(f"E2E {name} [threads={os.environ.get('GSA_STAGING_THREADS','auto')}]: {dt*1e3:.1f} ms  {N*(d+2)/dt/1e9:.2f} Grows/s  {16.0*d*N/dt/1e9:.1f} GB/s  S1[0]={res.S1[0]:.6f}")
