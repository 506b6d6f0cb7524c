#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Sobol path (BASELINE.json):

    Saltelli rows evaluated + reduced per second, rows = nboot * (N / nboot) * (d + 2),
    on  gsa(sobol_g, Sobol(order=[0,1]), A, B; batch=true)  with d = 100, N = 2^26  (config 5),
    strong-scaled over 1/2/4/8 B200 (contiguous column shards, ONE all-reduce of the partial sums).

A "step" is one full pass of the path over the design through ONE public call per rank (gsa_sobol_run_sharded): fused
design-swap + model + estimator kernel on this rank's shard whose last CTA also sums the tile records, (N>1) all-reduces the
nboot*nout*nstat partial sums over NVLink peer memory and writes them to pinned host memory (one kernel launch per step), and
the host epilogue that yields S1/ST.

  value      inputs (A, B shards) resident in HBM before the timed region; CUDA events, max over ranks
  e2e        the same metric through the public API with HOST (pinned) A and B: the chunked H2D copy of the
             whole shard is inside the timed region, every step
  roofline   the fused kernel alone: 16*d bytes per sample / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline   the CPU restatement of the reference (oracle/gsa_ref.c, OpenMP, all host cores) on a bounded
             sample of the same workload (Julia is not in the image, so the reference itself cannot run)

  configs    (N = 1) BASELINE configs 1-4, K1 in steady state and the estimator-only kernel, each with its own kernel time,
             algorithmic bytes and fraction of the measured HBM peak -- measured inside this run, with its clocks

`--impl reference` times only that CPU restatement and prints a line of the same shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

METRIC = "saltelli_rows_per_s"
UNIT = "rows/s"
D = 100
SOBOL_G_A = [0.0, 1.0, 4.5, 9.0] + [99.0] * (D - 4)     # Saltelli's standard G-function coefficients, padded with 99


def workload(log2n: int):
    N = 1 << log2n
    return {"workload": f"sobol_g d={D} S1/ST (order=[0,1], nboot=1, Jansen1999) N=2^{log2n}, unrandomised Sobol design "
                        f"(BASELINE config 5)", "d": D, "N": N, "step": D + 2, "rows": N * (D + 2),
            "a": "[0,1,4.5,9,99 x 96]", "l2": "inputs_larger_than_L2"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle, i.e. the restatement of the reference's own structure (materialise all_points, batch model,
# slice-and-sum analysis).  all_points for config 5 would be 5.5 TB, so a bounded sample of the workload is timed.
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample(log2n_cpu: int, reps: int):
    import gsa_ref
    gsa_ref.build()
    gsa_ref.use_all_cores()
    n = 1 << log2n_cpu
    lb, ub = np.zeros(D), np.ones(D)
    A, B = gsa_ref.sobol_design(n, lb, ub)
    a = np.array(SOBOL_G_A)
    best, rows = None, 0
    for _ in range(reps):
        t0 = time.perf_counter()
        r = gsa_ref.gsa_sobol("sobol_g", a, A, B)
        dt = time.perf_counter() - t0
        rows = r.rows
        best = dt if best is None else min(best, dt)
    return {"value": rows / best, "unit": UNIT, "cores": gsa_ref.num_threads(), "kind": "port",
            "sample": f"same model/estimator, N=2^{log2n_cpu} ({rows} rows, all_points materialised: "
                      f"{8 * D * rows / 1e9:.1f} GB), best of {reps}", "seconds": best}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.cpu_log2n)          # the config this arm really runs: a bounded sample of the headline workload
    w["sampled_from"] = f"N=2^{args.log2n} (the reference's all_points for the full config would be 5.5 TB)"
    w["extrapolated"] = True
    times, rows = [], 0
    import gsa_ref
    gsa_ref.build()
    gsa_ref.use_all_cores()
    n = 1 << args.cpu_log2n
    A, B = gsa_ref.sobol_design(n, np.zeros(D), np.ones(D))
    a = np.array(SOBOL_G_A)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = gsa_ref.gsa_sobol("sobol_g", a, A, B)
        dt = time.perf_counter() - t0
        rows = r.rows
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    val = rows * len(times) / tot
    cb = {"value": val, "unit": UNIT, "cores": gsa_ref.num_threads(), "kind": "port",
          "sample": f"each step = the CPU restatement on N=2^{args.cpu_log2n} samples ({rows} rows) of the same workload"}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": w, "cpu_baseline": cb,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "Julia is not installed in the image: this is oracle/gsa_ref.c (C + OpenMP restatement of "
                    "src/sobol_sensitivity.jl:94-405), pinned to the reference's golden vectors"}
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def wait_first_sample(self, timeout_s: float = 5.0):
        """nvidia-smi takes ~1 s to print its first line: a short timed region would otherwise end before it (no samples)"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def stop(self, t0, t1):
        """terminate nvidia-smi and summarise the samples taken inside the timed region [t0, t1] (perf_counter clock)"""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        rows = list(self.rows)
        inside = [r[1:] for r in rows if t0 <= r[0] <= t1]
        if not inside:   # region shorter than the sampling period: take the samples closest to it
            inside = [r[1:] for r in sorted(rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]]
        self.rows = inside
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
            if any(len(r) >= 8 and r[col].lower() == "active" for r in self.rows):
                reasons.append(name)
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# host memory placement: a rank's host shard lives on the NUMA node its GPU hangs off
# ------------------------------------------------------------------------------------------------------------------
def gpu_numa_node(idx: int):
    """NUMA node of GPU `idx` from sysfs (None if the platform does not say: VMs often report -1)"""
    bus = None
    try:
        import torch
        p = torch.cuda.get_device_properties(idx)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:                                   # noqa: BLE001 - older torch: ask nvidia-smi
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            if out:
                dom, rest = out.split(":", 1)
                bus = f"{int(dom, 16):04x}:{rest.lower()}"
        except Exception:                               # noqa: BLE001
            bus = None
    if not bus:
        return None
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            n = int(f.read())
        return n if n >= 0 else None
    except Exception:                                   # noqa: BLE001 - best effort
        return None


def bind_to_node(addr: int, nbytes: int, node) -> str:
    """mbind(MPOL_BIND) of the page range BEFORE it is first touched; returns how the memory ended up placed"""
    if node is None or node >= 64:
        return "default (GPU NUMA node unknown)"
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    page = 4096
    start = addr & ~(page - 1)
    length = ((addr + nbytes + page - 1) & ~(page - 1)) - start
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(237, ctypes.c_void_p(start), ctypes.c_ulong(length), ctypes.c_int(2), ctypes.byref(mask), ctypes.c_ulong(65),
                      ctypes.c_uint(0))
    return f"mbind to node {node}" if rc == 0 else f"default (mbind failed, errno {ctypes.get_errno()})"


def W_c4(nout=256, d=20):
    o, i = np.meshgrid(np.arange(1, nout + 1), np.arange(1, d + 1), indexing="ij")
    return 1.0 + ((o * 31 + i * 17) % 13) / 13.0


# measured on a B200 at 1965 MHz (tools/dmma_peak.cu -> profiles/r2_dmma_peak.jsonl): ns per warp instruction per SM sub-partition
DMMA_NS, DFMA_NS = 8.2, 1.125
# (DMMA, other FP64 instructions) per warp step in the main loop of the config's kernel
FP64_PIPE_WORK = {"c4_linear_d20_256out_N2^22": (15, 10), "x_sobolg_d20_N2^21_order2": (6, 57)}


def kernel_sass_hash(lib, mangled):
    """sha256[:16] of the SASS of one kernel of `lib` (cuobjdump), or None: identifies the kernel code independently of the rest of
    the library (profiles/r2_traffic.json names the kernel its DRAM traffic was captured from)."""
    import hashlib, re, subprocess
    try:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, lib], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    lines = [l.strip() for l in out.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()[:16] if lines else None


def configs_block(g, h, peak, reps=5):
    """BASELINE configs 1-4 (fused path), K1 in steady state, the estimator-only kernel: device-resident inputs, kernel time from
    CUDA events around the dominant kernel (gsa_last_timing), the public call's wall time beside it."""
    import torch
    out = {}
    cfgs = {
        "c1_ishigami_d3_N600000": dict(model=("ishigami", [7.0, 0.1]), d=3, N=600000, order=[0, 1], nboot=1, lb=-np.pi, ub=np.pi),
        "c2_ishigami_d3_N2^20_order2_nboot1000": dict(model=("ishigami", [7.0, 0.1]), d=3, N=1 << 20, order=[0, 1, 2], nboot=1000, lb=-np.pi, ub=np.pi),
        "c3_sobolg_d8_N2^22_order2": dict(model=("sobol_g", [0, 1, 4.5, 9, 99, 99, 99, 99]), d=8, N=1 << 22, order=[0, 1, 2], nboot=1, lb=0.0, ub=1.0),
        "c4_linear_d20_256out_N2^22": dict(model=("linear", W_c4()), d=20, N=1 << 22, order=[0, 1], nboot=1, lb=0.0, ub=1.0),
        # not a BASELINE config: the second-order path beyond d = 8 (lane-per-coordinate source of the pair kernel)
        "x_sobolg_d20_N2^21_order2": dict(model=("sobol_g", list(np.linspace(0, 9, 20))), d=20, N=1 << 21, order=[0, 1, 2], nboot=1, lb=0.0, ub=1.0),
    }
    for name, c in cfgs.items():
        d, N = c["d"], c["N"]
        model, method = g.DeviceModel(*c["model"]), g.Sobol(order=c["order"], nboot=c["nboot"])
        step = 2 * d + 2 if 2 in c["order"] else d + 2
        n = N // c["nboot"]
        rows = c["nboot"] * n * step
        A, B = g.generate_design_matrices(N, np.full(d, c["lb"]), np.full(d, c["ub"]), 2, device=True, handle=h)
        torch.cuda.synchronize()
        ks, ts, ws = [], [], []
        for i in range(reps + 2):
            t0 = time.perf_counter()
            res = g.gsa(model, method, A, B, batch=True, handle=h)          # the public call, result on the host
            w = time.perf_counter() - t0
            t, k, nl = h.last_timing()
            if i >= 2:
                ks.append(k); ts.append(t); ws.append(w)
        k, t, w = float(np.median(ks)), float(np.median(ts)), float(np.median(ws))
        alg = 16.0 * d * c["nboot"] * n
        out[name] = {"rows": rows, "kernel_ms": k, "device_ms": t, "call_wall_ms": 1e3 * w, "launches": nl, "algorithmic_bytes": alg,
                     "achieved_gbs": alg / (k * 1e-3) / 1e9, "frac": alg / (k * 1e-3) / 1e9 / peak, "rows_per_s_call": rows / w,
                     "S1_head": np.ravel(res.S1)[:3].tolist()}
        if name in FP64_PIPE_WORK:
            # the kernels whose sums run as mma.sync.m8n8k4.f64: DMMA shares the FP64 pipe with DFMA/DADD/DMUL (measured: the times
            # add, profiles/r2_dmma_peak.jsonl), so their second roofline is that pipe: per warp step (4 samples) and SM sub-partition,
            # DMMA_NS per DMMA + DFMA_NS per other FP64 instruction (counts from the SASS main loop, profiles/r2_sass_summary.txt)
            nd, nf = FP64_PIPE_WORK[name]
            sms = torch.cuda.get_device_properties(0).multi_processor_count
            floor_ms = (c["nboot"] * n / 4.0) / (4 * sms) * (nd * DMMA_NS + nf * DFMA_NS) * 1e-6
            out[name]["fp64_pipe"] = {"dmma_per_warp_step": nd, "other_fp64_per_warp_step": nf, "floor_ms": floor_ms, "frac": floor_ms / k,
                                      "ceilings": "36.9 TFLOP/s DMMA, 33.7 TFLOP/s DFMA, one shared pipe (profiles/r2_dmma_peak.jsonl)"}
        del A, B
        torch.cuda.empty_cache()
    # eFAST: the reference's test size, and a larger one; wall time of the public call (design in registers + model +
    # pruned-DFT analysis), spectrum kernel alone beside it
    for name, nm, params, dd, S in (("efast_ishigami_d4_S15000", "ishigami", [7.0, 0.1], 4, 15000), ("efast_sobolg_d20_S2^17", "sobol_g", list(np.linspace(0, 9, 20)), 20, 1 << 17)):
        rng = [[-np.pi, np.pi]] * dd if nm == "ishigami" else [[0.0, 1.0]] * dd
        m = g.DeviceModel(nm, params)
        ws, ks = [], []
        for i in range(reps + 1):
            t0 = time.perf_counter()
            r = g.gsa_efast(m, g.eFAST(), rng, samples=S, phi=np.linspace(0.1, 1.9, dd), handle=h)
            w = time.perf_counter() - t0
            if i >= 1:
                ws.append(w); ks.append(h.last_timing()[1])
        out[name] = {"points": S * dd, "call_wall_ms": 1e3 * float(np.median(ws)), "spectrum_kernel_ms": float(np.median(ks)),
                     "points_per_s_call": S * dd / float(np.median(ws)), "ST_head": np.ravel(r.ST)[:3].tolist()}
    # K1: first touch of fresh memory vs steady state (the second and later calls into the same buffers)
    for d, N in ((100, 1 << 22), (8, 1 << 24)):
        mats = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, handle=h)
        torch.cuda.synchronize()
        first = h.last_timing()[0]
        import ctypes as C
        ptrs = (C.c_void_p * 2)(*[m.t().data_ptr() for m in mats])
        lbv, ubv = np.zeros(d), np.ones(d)
        ts = []
        for _ in range(reps):
            g.sobol.check(g.lib().gsa_sobol_design(h.ptr, N, d, 2, lbv.ctypes.data, ubv.ctypes.data, 0, N, ptrs, 1))
            h.synchronize()
            ts.append(h.last_timing()[0])
        t = float(np.median(ts))
        alg = 16.0 * d * N
        out[f"k1_design_d{d}_N2^{N.bit_length() - 1}"] = {"first_touch_ms": first, "steady_ms": t, "algorithmic_bytes": alg,
                                                          "achieved_gbs": alg / (t * 1e-3) / 1e9, "frac": alg / (t * 1e-3) / 1e9 / peak}
        del mats
        torch.cuda.empty_cache()
    # estimator-only kernel on a precomputed Y (random values: only the timing matters here; parity is in tests/)
    for name, d, n, order in (("y_c3_shape_d8_order2", 8, 1 << 22, [0, 1, 2]), ("y_c5_shape_d100", 100, 1 << 22, [0, 1]), ("y_d20_order2", 20, 1 << 21, [0, 1, 2])):
        step = 2 * d + 2 if 2 in order else d + 2
        Y = torch.rand(step * n, dtype=torch.float64, device="cuda")
        ks = []
        for i in range(reps + 1):
            g.gsa_sobol_all_y_analysis(g.Sobol(order=order), Y, d, n, handle=h)
            if i >= 1:
                ks.append(h.last_timing()[1])
        k = float(np.median(ks))
        alg = 8.0 * step * n
        out[name] = {"rows": step * n, "kernel_ms": k, "algorithmic_bytes": alg, "achieved_gbs": alg / (k * 1e-3) / 1e9,
                     "frac": alg / (k * 1e-3) / 1e9 / peak}
        del Y
        torch.cuda.empty_cache()
    return out


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import gsa_loader
    g = gsa_loader.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = workload(args.log2n)
    N, d = w["N"], D
    col0, ncols = g.shard_columns(N, rank, world)
    model = g.DeviceModel("sobol_g", SOBOL_G_A)
    method = g.Sobol(order=[0, 1], nboot=1)
    h = g.default_handle(local)
    lb, ub = np.zeros(d), np.ones(d)

    # ---- setup (untimed): this rank's shard of the design, generated in HBM by K1 -------------------------------
    A, B = g.generate_design_matrices(N, lb, ub, 2, device=True, col0=col0, ncols=ncols, handle=h)
    torch.cuda.synchronize()
    k1_ms = h.last_timing()[0]
    ns = g.nstat(method, d)

    # the single collective of the path (a few KB) runs over NVLink peer memory (csrc/peer.cu), inside the fused kernel's last CTA;
    # if the mailboxes cannot be mapped every rank falls back to partials + torch.distributed (NCCL) + finalize
    pg, collective = None, "none"
    if world > 1:
        collective = "nccl"
        if args.collective == "peer":
            try:
                pg = g.PeerGroup(h, ns)
                okf = torch.ones(1, device="cuda")
            except Exception as e:                          # noqa: BLE001 - any rank failing sends every rank to NCCL
                sys.stderr.write(f"[bench rank {rank}] peer mailboxes unavailable ({e}); using NCCL\n")
                pg, okf = None, torch.zeros(1, device="cuda")
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if float(okf) < 1.0:
                pg = None
            else:
                collective = "peer-memory exchange fused into the kernel's last CTA (gsa_sobol_run_sharded)"
    one_call = world == 1 or pg is not None
    if one_call:
        sstep = g.ShardedStep(model, method, A, B, N, col0, handle=h)
        step = sstep.run                                    # ONE public call: kernel (+ exchange) + epilogue
    else:
        def step():
            return g.gsa_sharded(model, method, A, B, N, col0, handle=h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                     # nvidia-smi needs ~1 s to start: launch it before the warm-up
        sampler.wait_first_sample()
    for _ in range(args.warmup):
        res = step()
    launches_per_step = h.last_timing()[2]
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    tw0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        res = step()
        kernel_ms.append(h.last_timing()[1])               # (waits for the step's closing event: the result is already on the host)
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    clocks = None
    if rank == 0:
        clocks = sampler.stop(tw0, tw1)
    t = torch.tensor([ms_total, float(np.mean(kernel_ms))], dtype=torch.float64, device="cuda")
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kms = float(t[0]), float(t[1])
    per_rank_kernel_ms = [round(float(x[1]), 4) for x in per_rank]
    value = w["rows"] * args.steps / (ms_total * 1e-3)

    # ---- extra: the same analysis with the design GENERATED inside the fused kernel (gsa(model, Sobol(), p_range; samples)):
    #      no A/B in memory, no HBM reads.  Reported beside the headline, not instead of it. --------------------------------
    gen, res_g = None, None
    if not args.no_gen and one_call:
        gstep = g.ShardedStep(model, method, None, None, N, col0, handle=h, p_range=(lb, ub), ncols=ncols)
        for _ in range(args.warmup):
            res_g = gstep.run()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            res_g = gstep.run()
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gen = {"value": w["rows"] * args.steps / (float(tg[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(tg[0]) / args.steps,
               "hbm_bytes_read": 0, "note": "K1 fused into the product kernel: the Sobol points are generated in registers "
               "(bit-identical to the materialised design), gsa_sobol_run_sharded_generated",
               "max_abs_diff_vs_resident": float(max(np.abs(res_g.S1 - res.S1).max(), np.abs(res_g.ST - res.ST).max()))}

    # ---- N > 1: the single-process entry (gsa_multi_*: what a plain Julia ccall uses) on ALL devices, checked against the
    #      one-process-per-GPU result of this run -------------------------------------------------------
    multi = None
    if world > 1 and not args.no_multi:
        barrier()
        if rank == 0:
            try:
                m = g.GsaMulti(list(range(world)))
                t0 = time.perf_counter()
                rm = m.gsa_p_range(model, method, [[0.0, 1.0]] * d, N)
                dtm = time.perf_counter() - t0
                t0 = time.perf_counter()
                rm = m.gsa_p_range(model, method, [[0.0, 1.0]] * d, N)
                dtm = min(dtm, time.perf_counter() - t0)
                ref_r = res_g if res_g is not None else res
                diff = float(max(np.abs(rm.S1 - ref_r.S1).max(), np.abs(rm.ST - ref_r.ST).max()))
                multi = {"entry": "gsa_multi_sobol_run_generated over %d devices from ONE process (rank 0)" % world,
                         "max_abs_diff_vs_sharded": diff, "ok": bool(diff <= 1e-12), "rows_per_s": w["rows"] / dtm, "ms": 1e3 * dtm}
                m.close()
            except Exception as e:                          # noqa: BLE001 - reported, not fatal for the headline
                multi = {"error": str(e), "ok": False}
        barrier()

    # ---- e2e: the public call with HOST inputs; H2D of the whole shard inside the timed region --------------------------
    e2e = None
    if not args.no_e2e:
        nbytes = 8 * d * ncols
        hostA = np.empty((ncols, d), dtype=np.float64)
        hostB = np.empty((ncols, d), dtype=np.float64)
        node = gpu_numa_node(local)
        placement = [bind_to_node(arr.ctypes.data, nbytes, node) for arr in (hostA, hostB)][0]
        tA, tB = torch.from_numpy(hostA), torch.from_numpy(hostB)
        tA.copy_(A.t())                                     # fill the host copies (untimed; first touch under the policy above)
        tB.copy_(B.t())
        torch.cuda.synchronize()
        for arr in (hostA, hostB):
            g.pin_host(arr)                                 # gsa_host_register: what a Julia caller does with GSACuda.pin!(A)
        del A, B
        if one_call:
            del sstep
        torch.cuda.empty_cache()
        hA, hB = tA.t(), tB.t()                             # (d, ncols) views in Julia layout
        h.set_stream(None)
        if world == 1:
            e2e_step = lambda: g.gsa(model, method, hA, hB, batch=True, handle=h)          # the call a user makes
        elif pg is not None:
            e2e_step = g.ShardedStep(model, method, hA, hB, N, col0, handle=h, torch_stream=False).run     # gsa_sobol_run_sharded
        else:
            e2e_step = lambda: g.gsa_sharded(model, method, hA, hB, N, col0, handle=h)

        def timed(nsteps, nwarm):
            for _ in range(nwarm):
                r = e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(nsteps):
                r = e2e_step()
            mine = time.perf_counter() - t0
            barrier()
            tt = torch.tensor([time.perf_counter() - t0, mine], dtype=torch.float64, device="cuda")
            allr = [tt.clone() for _ in range(world)]
            if world > 1:
                dist.all_gather(allr, tt)
            return r, max(float(x[0]) for x in allr), [float(x[1]) for x in allr]

        res_e, dt, per = timed(args.e2e_steps, args.e2e_warmup)
        e2e = {"value": w["rows"] * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * d * N,
               "d2h_bytes_per_step": int(ns * 8 * world), "steps": args.e2e_steps, "warmup": args.e2e_warmup,
               "ms_per_step": 1e3 * dt / args.e2e_steps, "host_memory": "pinned (cudaHostRegister), " + placement,
               "aggregate_h2d_gbs": 2 * 8 * d * N * args.e2e_steps / dt / 1e9,
               "per_gpu_h2d_gbs": [round(2 * nbytes * args.e2e_steps / p / 1e9, 1) for p in per],
               "entry": "gsa_sobol_run" if world == 1 else "gsa_sobol_run_sharded (one call per rank)",
               "max_abs_diff_vs_resident": float(max(np.abs(res_e.S1 - res.S1).max(), np.abs(res_e.ST - res.ST).max()))}
        for arr in (hostA, hostB):
            g.unpin_host(arr)
        # the same call on PAGEABLE memory (what a plain Julia Matrix{Float64} is): staged through the library's pinned buffers
        res_p, dtp, _ = timed(1, 0)
        e2e["pageable"] = {"value": w["rows"] / dtp, "unit": UNIT, "ms_per_step": 1e3 * dtp, "aggregate_h2d_gbs": 2 * 8 * d * N / dtp / 1e9,
                           "host_memory": "pageable (numpy), staged by the library",
                           "max_abs_diff_vs_resident": float(max(np.abs(res_p.S1 - res.S1).max(), np.abs(res_p.ST - res.ST).max()))}
        del hA, hB, tA, tB, hostA, hostB
        # the platform's own ceiling: all ranks copy 2 GiB of torch-pinned memory to their GPU at the same time (no kernels, no library)
        if world > 1:
            src = torch.empty(1 << 28, dtype=torch.float64).pin_memory()
            dst = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
            dst.copy_(src, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            barrier()
            tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            e2e["platform_concurrent_h2d_gbs"] = 3 * world * src.numel() * 8 / float(tc[0]) / 1e9
            del src, dst

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 16.0 * d * ncols                          # per launch on one GPU: A and B shards read once
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch from the committed ncu capture of the SHIPPED kernel (same binary: checked by its hash)
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tr = json.load(f)["fused_product_kernel"]
            import hashlib
            with open(g.LIB_PATH, "rb") as f:
                sha = hashlib.sha256(f.read()).hexdigest()[:16]
            same = tr.get("lib_sha256_16") in (None, sha)
            if not same and tr.get("kernel_sass_sha256_16"):     # a rebuilt library: the same kernel code? (compare its SASS)
                same = kernel_sass_hash(g.LIB_PATH, tr.get("kernel_mangled", "")) == tr["kernel_sass_sha256_16"]
            if same:
                traffic, traffic_src = tr["dram_bytes_per_algorithmic_byte"] * alg_bytes, tr["source"]
            else:
                traffic_src = "profiles/r2_traffic.json was captured from another build of libgsa_cuda.so: not reported"
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"bound": "hbm", "kernel": "fused_product_kernel<T,CPL> (design swap + Sobol-G + Jansen sums; its last CTA also "
                    "does stage 2, the exchange and the host copy)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peak_src, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                    "per_rank_kernel_ms": per_rank_kernel_ms, "k1_design_first_touch_ms": k1_ms,
                    "k1_design_first_touch_gbs": 16.0 * d * ncols / (k1_ms * 1e-3) / 1e9,
                    "host_overhead_ms_per_step": ms_total / args.steps - kms}
        configs = None
        if world == 1 and not args.no_configs:
            try:
                configs = configs_block(g, h, peak)
            except Exception as e:                          # noqa: BLE001 - the headline line must still be printed
                configs = {"error": str(e)}
        cb = None if args.no_cpu else cpu_sample(args.cpu_log2n, 2)
        Vi = 1.0 / (3.0 * (1.0 + np.array(SOBOL_G_A)) ** 2)
        V = np.prod(1.0 + Vi) - 1.0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": w, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "launches_per_step": launches_per_step,
                "collective": collective, "roofline": roofline, "generated_design": gen, "multi_single_process": multi,
                "configs": configs, "cpu_baseline": cb, "clocks": clocks,
                "check": {"S1_0": float(res.S1[0]), "S1_0_analytic": float(Vi[0] / V), "ST_0": float(res.ST[0]),
                          "ST_0_analytic": float(Vi[0] * np.prod(1 + Vi) / (1 + Vi[0]) / V),
                          "parity": "tests/test_gpu_parity.py::test_config5_full_size_vs_streamed_oracle (1e-12 at N=2^26)"}}
        emit(line)
    if pg is not None:
        if pg.status() != 0:
            raise SystemExit("peer all-reduce timed out waiting for a rank")
        pg.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner, library
    chatter of any rank) goes to stderr from here on."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=26, help="N = 2^log2n samples in total (BASELINE config 5: 26)")
    ap.add_argument("--cpu-log2n", type=int, default=17, help="bounded sample size of the CPU baseline")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-warmup", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gen", action="store_true")
    ap.add_argument("--no-multi", action="store_true", help="N > 1: skip the single-process (gsa_multi) cross-check")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the per-config block (BASELINE configs 1-4, K1, estimator on Y)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N > 1: how the partial sums are all-reduced")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
