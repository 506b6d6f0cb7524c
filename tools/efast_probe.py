"""eFAST timing on a B200: public call and spectrum step, FFT route vs pruned DFT (GSA_EFAST_DFT=1)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import gsa_loader
g = gsa_loader.load()
h = g.default_handle(0)
for nm, params, dd, S in (("ishigami", [7.0, 0.1], 4, 15000), ("sobol_g", list(np.linspace(0, 9, 20)), 20, 1 << 17), ("sobol_g", list(np.linspace(0, 9, 20)), 20, 15000)):
    rng = [[-np.pi, np.pi]] * dd if nm == "ishigami" else [[0.0, 1.0]] * dd
    m = g.DeviceModel(nm, params)
    for dft in ("0", "1"):
        os.environ["GSA_EFAST_DFT"] = dft
        ws, ks = [], []
        for i in range(6):
            t0 = time.perf_counter()
            r = g.gsa_efast(m, g.eFAST(), rng, samples=S, phi=np.linspace(0.1, 1.9, dd), handle=h)
            w = time.perf_counter() - t0
            if i >= 1:
                ws.append(w); ks.append(h.last_timing()[1])
        print(json.dumps({"model": nm, "d": dd, "samples": S, "route": "dft" if dft == "1" else "fft", "call_wall_ms": 1e3 * float(np.median(ws)),
                          "spectrum_ms": float(np.median(ks)), "launches": h.last_timing()[2], "ST0": float(np.ravel(r.ST)[0])}), flush=True)
