"""Times K1 (gsa_sobol_design) on the device: python tools/k1_probe.py [log2n] [d]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = 1 << log2n
h = g.default_handle()
for it in range(4):
    A, B = g.generate_design_matrices(n, np.zeros(d), np.ones(d), 2, device=True, handle=h)
    torch.cuda.synchronize()
    ms = h.last_timing()[0]
    print(f"K1 d={d} n=2^{log2n} call {it}: {ms:.3f} ms  {16.0 * d * n / ms / 1e6:.1f} GB/s")
    del A, B
