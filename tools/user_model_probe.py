"""Debug helper: builds a user __device__ model, links it (gsa_model_register_fatbin) and runs one small analysis."""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import gsa_loader
g = gsa_loader.load()
SRC = r'''
extern "C" __device__ void my_model(const double* x, int d, const double* p, int np, double* y, int nout) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += p[i] * x[i] * x[i];
    y[0] = s + x[0] * x[1];
    if (nout > 1) y[1] = sin(x[0]) + x[d - 1];
}
'''
tmp = tempfile.mkdtemp()
open(os.path.join(tmp, "um.cu"), "w").write(SRC)
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-ccbin", "/usr/bin/g++", "-rdc=true", "-fatbin", "-gencode",
                       "arch=compute_100a,code=sm_100a", "-o", os.path.join(tmp, "um.fatbin"), os.path.join(tmp, "um.cu")])
d, N = 5, 2000
second = len(sys.argv) > 1 and sys.argv[1] == "2"
p = np.array([0.5, 1.0, 1.5, 2.0, 2.5])
rng = np.random.default_rng(0)
A, B = rng.random((d, N)), rng.random((d, N))
m = g.DeviceModel.from_fatbin(os.path.join(tmp, "um.fatbin"), "my_model", nout=2, params=p)
print("registered", m.model_id, flush=True)
res = g.gsa(m, g.Sobol(order=[0, 1, 2] if second else [0, 1]), A, B, batch=True)
print("S1", res.S1, flush=True)
