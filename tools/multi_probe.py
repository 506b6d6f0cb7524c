import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
print("devices", torch.cuda.device_count(), flush=True)
m = g.GsaMulti(ndev=int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count())
print("created", flush=True)
rng = np.random.default_rng(0)
A, B = rng.random((3, 10000)) * 6 - 3, rng.random((3, 10000)) * 6 - 3
res = m.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=4), A, B)
print("S1", res.S1, flush=True)
d, N = 100, 100003
a = np.array([0.0, 1.0, 4.5, 9.0] + [99.0] * 96)
A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2)   # host arrays (K1 on the GPU, copied back)
print("second run", A.flags["F_CONTIGUOUS"], A.shape, flush=True)
res = m.gsa(g.DeviceModel("sobol_g", a), g.Sobol(nboot=3), A, B)
print("S1", res.S1[:3], flush=True)
m.close()
print("closed", flush=True)
