"""Times one fused configuration with device-resident inputs: python tools/fused_probe.py model d log2n second nboot"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
name, d, log2n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
second = len(sys.argv) > 4 and sys.argv[4] == "1"
nboot = int(sys.argv[5]) if len(sys.argv) > 5 else 1
N = 1 << log2n
params = {"sobol_g": ([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d], "ishigami": [7.0, 0.1]}[name]
model, method = g.DeviceModel(name, params), g.Sobol(order=[0, 1, 2] if second else [0, 1], nboot=nboot)
h = g.default_handle()
A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, handle=h)
ks = []
for _ in range(5):
    g.sobol_partials(model, method, A, B, N, 0, handle=h)
    torch.cuda.synchronize()
    ks.append(h.last_timing()[1])
k = float(np.median(ks))
step = 2 * d + 2 if second else d + 2
print(f"FUSED {name} d={d} N=2^{log2n} second={second} nboot={nboot} [MMA={os.environ.get('GSA_PAIRS_MMA','-')} GEN={os.environ.get('GSA_FORCE_GENERIC','-')}]: {k:.3f} ms  {N*step/k/1e6:.1f} Grows/s  {16.0*d*N/k/1e6:.0f} GB/s")
