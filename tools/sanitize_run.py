"""Small invocations of every kernel family (ragged sizes, every block count, all estimators) without the oracle: a quick
exercise run, and the workload for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_run.py` where the tool is
available (it is closed on the round-1 GPU pool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
h = g.default_handle()
rng = np.random.default_rng(0)


def design(N, d):
    return g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, handle=h)


def sg(d):
    return g.DeviceModel("sobol_g", ([0.0, 1.0, 4.5, 9.0] + [99.0] * d)[:d])


# product-form first order (config 5 shape), with replicates and a ragged tail; generated design
A, B = design(10007, 100)
for est in ("Jansen1999", "Sobol2007", "Homma1996"):
    g.gsa(sg(100), g.Sobol(nboot=3), A, B, batch=True, Ei_estimator=est, handle=h)
g.gsa(sg(100), g.Sobol(nboot=2), [[0.0, 1.0]] * 100, samples=3001, batch=True)
# pair kernels: tile source (d = 8, 4), lane-per-coordinate source (d = 20, 32), Y source (d = 20, 3 outputs; d = 8)
for d, N in ((8, 4099), (4, 1001), (20, 3003), (32, 1030), (12, 37)):
    A, B = design(N, d)
    for est in ("Jansen1999", "Homma1996"):
        g.gsa(sg(d), g.Sobol(order=[0, 1, 2], nboot=2), A, B, batch=True, Ei_estimator=est, handle=h)
for d, n, nout in ((20, 1501, 3), (8, 777, 1), (32, 300, 1), (9, 5, 2)):
    Y = torch.rand((nout, (2 * d + 2) * n * 2), dtype=torch.float64, device="cuda")
    os.environ["GSA_PAIRS_MMA"] = "1"
    g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2], nboot=2), Y if nout > 1 else Y[0], d, n, handle=h)
    os.environ.pop("GSA_PAIRS_MMA")
# linear Gram path: every block count class, all estimators, S2
for d, nout, N in ((20, 40, 5003), (32, 3, 1201), (1, 2, 100), (7, 300, 2049)):
    A, B = design(N, d)
    W = rng.normal(size=(nout, d))
    for est in ("Jansen1999", "Sobol2007", "Homma1996", "Janon2014"):
        g.gsa(g.DeviceModel("linear", W), g.Sobol(order=[0, 1, 2], nboot=2), A, B, batch=True, Ei_estimator=est, handle=h)
# Ishigami register kernels (one tile per replicate: records straight into the partial sums), estimator on Y, design kernel
A, B = g.generate_design_matrices(50000, -np.pi * np.ones(3), np.pi * np.ones(3), 2, device=True, handle=h)
g.gsa(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=50), A, B, batch=True, handle=h)
Y = torch.rand(102 * 4000, dtype=torch.float64, device="cuda")
g.gsa_sobol_all_y_analysis(g.Sobol(), Y, 100, 4000, handle=h)
# host (streamed) inputs through the new kernels
A, B = design(20011, 8)
g.gsa(sg(8), g.Sobol(order=[0, 1, 2]), A.cpu().numpy(), B.cpu().numpy(), batch=True, handle=h)
# eFAST: FFT route (even packed size with every radix, odd size with the real-input first pass) and the pruned DFT (a prime size)
for S in (3000, 405, 1024, 4099):
    Y = torch.rand((2, S * 5), dtype=torch.float64, device="cuda")
    g.gsa_efast_all_y_analysis(g.eFAST(4), Y, 5, S, handle=h)
g.gsa_efast(g.DeviceModel("ishigami", [7.0, 0.1]), g.eFAST(), [[-np.pi, np.pi]] * 4, samples=3000, phi=np.linspace(0.1, 1.9, 4), handle=h)
torch.cuda.synchronize()
print("sanitize_run done")
