"""User `__device__` models: timing of the three routes a model can take, on one GPU:
    generic     gsa_model_register_fatbin: the full-vector model on the generic kernel, O(d^2) per sample
    separable   gsa_model_register_separable: the per-coordinate function on the O(d) product kernel
    built-in    the same function as a built-in model (upper bound: everything inlined)
for a product-form transcendental model  f(x) = prod_k (1 + p_k sin(pi x_k)) / (1 + 2 p_k / pi)  at d = 20 and d = 100.
Prints one JSON line per measurement; run under ncu with --only <route>,<d> for the FP64-pipe evidence."""
import json, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gsa_loader
g = gsa_loader.load()
SRC = r'''
extern "C" __device__ double trig_factor(int k, double x, const double* p, int np) {
    return (1.0 + p[k] * sin(3.141592653589793 * x)) / (1.0 + 0.6366197723675814 * p[k]);
}
extern "C" __device__ void trig_model(const double* x, int d, const double* p, int np, double* y, int nout) {
    double r = 1.0;
    for (int k = 0; k < d; ++k) r *= (1.0 + p[k] * sin(3.141592653589793 * x[k])) / (1.0 + 0.6366197723675814 * p[k]);
    y[0] = r;
}
extern "C" __device__ double sobolg_factor(int k, double x, const double* p, int np) { return (fabs(4.0 * x - 2.0) + p[k]) / (1.0 + p[k]); }
'''
SRC_LTO = r'''
extern "C" __device__ double gsa_user_factor(int k, double x, const double* p, int np) {
    return (1.0 + p[k] * sin(3.141592653589793 * x)) / (1.0 + 0.6366197723675814 * p[k]);
}
extern "C" __device__ void gsa_user_model(const double* x, int d, const double* p, int np, double* y, int nout) {
    double r = 1.0;
    for (int k = 0; k < d; ++k) r *= (1.0 + p[k] * sin(3.141592653589793 * x[k])) / (1.0 + 0.6366197723675814 * p[k]);
    y[0] = r;
}
'''
only = None
if "--only" in sys.argv:
    only = sys.argv[sys.argv.index("--only") + 1].split(",")
tmp = tempfile.mkdtemp()
open(os.path.join(tmp, "um.cu"), "w").write(SRC)
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-ccbin", "/usr/bin/g++", "-O3", "-rdc=true", "-fatbin", "-gencode",
                       "arch=compute_100a,code=sm_100a", "-o", os.path.join(tmp, "um.fatbin"), os.path.join(tmp, "um.cu")])
fat = open(os.path.join(tmp, "um.fatbin"), "rb").read()
# the same factor under the default symbol name, with LTO-IR next to the SASS: nvJitLink inlines it into the kernel (no call)
open(os.path.join(tmp, "um_lto.cu"), "w").write(SRC_LTO)
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-ccbin", "/usr/bin/g++", "-O3", "-rdc=true", "-fatbin", "-gencode", "arch=compute_100a,code=lto_100a",
                       "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(tmp, "um_lto.fatbin"), os.path.join(tmp, "um_lto.cu")])
fat_lto = open(os.path.join(tmp, "um_lto.fatbin"), "rb").read()
h = g.default_handle()
PEAK_DFMA = 36.6e12
for d, N in ((20, 1 << 20), (100, 1 << 18)):
    p = np.linspace(0.1, 0.9, d)
    A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, handle=h)
    routes = {"generic": g.DeviceModel.from_fatbin(fat, "trig_model", params=p),
              "generic_lto": g.DeviceModel.from_fatbin(fat_lto, "gsa_user_model", params=p),
              "separable": g.DeviceModel.from_fatbin(fat, "trig_factor", params=p, kind="product"),
              "separable_lto": g.DeviceModel.from_fatbin(fat_lto, "gsa_user_factor", params=p, kind="product"),
              "separable_sobolg": g.DeviceModel.from_fatbin(fat, "sobolg_factor", params=p, kind="product"),
              "builtin_sobolg": g.DeviceModel("sobol_g", p)}
    results = {}
    for name, m in routes.items():
        if only and (name != only[0] or str(d) != only[1]):
            continue
        ks = []
        for i in range(4):
            res = g.gsa(m, g.Sobol(), A, B, batch=True, handle=h)
            ks.append(h.last_timing()[1])
        k = float(np.median(ks[1:]))
        results[name] = res
        rows = N * (d + 2)
        evals = N * (d + 2) * d if name.startswith("generic") else N * 2 * d          # per-coordinate function evaluations
        if name == "separable_lto" and "separable" in results:
            e = float(max(np.abs(res.S1 - results["separable"].S1).max(), np.abs(res.ST - results["separable"].ST).max()))
            print(json.dumps({"d": d, "lto_vs_call_max_abs_diff": e}), flush=True)
        print(json.dumps({"route": name, "d": d, "N": N, "kernel_ms": k, "rows_per_s": rows / (k * 1e-3), "hbm_gbs": 16.0 * d * N / (k * 1e-3) / 1e9,
                          "factor_evals_per_s": evals / (k * 1e-3), "S1_head": np.ravel(res.S1)[:3].tolist()}), flush=True)
    if "generic" in results and "separable" in results:
        e = float(max(np.abs(results["generic"].S1 - results["separable"].S1).max(), np.abs(results["generic"].ST - results["separable"].ST).max()))
        print(json.dumps({"d": d, "generic_vs_separable_max_abs_diff": e}), flush=True)
    del A, B
