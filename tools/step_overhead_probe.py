"""Where the per-step host overhead of bench.py's step goes (tiny N: the kernels are ~10 us)."""
import os, sys, time, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
d, N = 100, 1 << 14
a = [0.0, 1.0, 4.5, 9.0] + [99.0] * 96
model, method = g.DeviceModel("sobol_g", a), g.Sobol()
h = g.default_handle()
A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, handle=h)
ns = g.nstat(method, d)
partials = torch.empty((1, 1, ns), dtype=torch.float64, device="cuda")
host = torch.empty(partials.shape, dtype=torch.float64).pin_memory()
T = {"partials_call": 0.0, "d2h_sync": 0.0, "finalize": 0.0}
reps = 500
for it in range(reps + 50):
    if it == 50:
        T = {k: 0.0 for k in T}
        torch.cuda.synchronize(); t_all = time.perf_counter()
    t0 = time.perf_counter()
    g.sobol_partials(model, method, A, B, N, 0, handle=h, out=partials)
    t1 = time.perf_counter()
    host.copy_(partials, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()
    res = g.sobol_finalize(method, host.numpy(), d, N)
    t3 = time.perf_counter()
    T["partials_call"] += t1 - t0; T["d2h_sync"] += t2 - t1; T["finalize"] += t3 - t2
tot = time.perf_counter() - t_all
out = {k: round(1e6 * v / reps, 1) for k, v in T.items()}
out["step_us"] = round(1e6 * tot / reps, 1)
out["device_call_ms_last"] = h.last_timing()[0]
print(json.dumps(out))
