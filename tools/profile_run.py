"""One pass over every hot kernel at sizes larger than L2, for ncu:
    ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/profile_run.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
h = g.default_handle()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1

def run(fn):
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()

# config 5 shape: K1 (two dims per thread), fused product kernel, fused product kernel with generated design
d, N = 100, 1 << 21
a = [0.0, 1.0, 4.5, 9.0] + [99.0] * 96
lb, ub = np.zeros(d), np.ones(d)
model, method = g.DeviceModel("sobol_g", a), g.Sobol()
A, B = g.generate_design_matrices(N, lb, ub, 2, device=True, handle=h)
run(lambda: g.sobol_partials(model, method, A, B, N, 0, handle=h))
run(lambda: g.sobol_partials_generated(model, method, lb, ub, N, 0, N, handle=h))
del A, B
# estimator-only kernels on Y: d = 100 (shared-memory kernel), d = 8 order 2 (register kernel with cp.async ring)
Y = torch.rand((d + 2) * (1 << 20), dtype=torch.float64, device="cuda")
run(lambda: g.gsa_sobol_all_y_analysis(method, Y, d, 1 << 20, handle=h))
del Y
Y = torch.rand(18 * (1 << 22), dtype=torch.float64, device="cuda")
run(lambda: g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2]), Y, 8, 1 << 22, handle=h))
del Y
# config 3: Sobol G d = 8 order 2 (thread-per-sample evaluation into a warp tile + FP64 mma pair sums)
A, B = g.generate_design_matrices(1 << 22, np.zeros(8), np.ones(8), 2, device=True, handle=h)
m3 = g.DeviceModel("sobol_g", [0, 1, 4.5, 9, 99, 99, 99, 99])
run(lambda: g.sobol_partials(m3, g.Sobol(order=[0, 1, 2]), A, B, 1 << 22, 0, handle=h))
del A, B
# second order beyond the register kernels: Sobol G d = 20 (lane-per-coordinate mma source) and the estimator on its Y (warp tiles)
A, B = g.generate_design_matrices(1 << 21, np.zeros(20), np.ones(20), 2, device=True, handle=h)
m20 = g.DeviceModel("sobol_g", [0, 1, 4.5, 9] + [99.0] * 16)
run(lambda: g.sobol_partials(m20, g.Sobol(order=[0, 1, 2]), A, B, 1 << 21, 0, handle=h))
del A, B
Y = torch.rand(42 * (1 << 21), dtype=torch.float64, device="cuda")
run(lambda: g.gsa_sobol_all_y_analysis(g.Sobol(order=[0, 1, 2]), Y, 20, 1 << 21, handle=h))
del Y
# config 2: Ishigami d = 3, order 2, 1000 replicates
A, B = g.generate_design_matrices(1 << 20, -np.pi * np.ones(3), np.pi * np.ones(3), 2, device=True, handle=h)
run(lambda: g.sobol_partials(g.DeviceModel("ishigami", [7.0, 0.1]), g.Sobol(order=[0, 1, 2], nboot=1000), A, B, 1 << 20, 0, handle=h))
del A, B
# config 4: linear, d = 20, 256 outputs (FP64 mma Gram kernel)
o, i = np.meshgrid(np.arange(1, 257), np.arange(1, 21), indexing="ij")
W = 1.0 + ((o * 31 + i * 17) % 13) / 13.0
A, B = g.generate_design_matrices(1 << 22, np.zeros(20), np.ones(20), 2, device=True, handle=h)
run(lambda: g.sobol_partials(g.DeviceModel("linear", W), method, A, B, 1 << 22, 0, handle=h))
del A, B
# eFAST: design in registers + model, centring, Stockham FFT passes (efast_fft.cuh), power, indices
run(lambda: g.gsa_efast(g.DeviceModel("sobol_g", list(np.linspace(0, 9, 20))), g.eFAST(), [[0.0, 1.0]] * 20, samples=1 << 17, phi=np.linspace(0.1, 1.9, 20), handle=h))
print("profile_run done")
