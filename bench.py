#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Sobol path (BASELINE.json):

    Saltelli rows evaluated + reduced per second, rows = nboot * (N / nboot) * (d + 2),
    on  gsa(sobol_g, Sobol(order=[0,1]), A, B; batch=true)  with d = 100, N = 2^26  (config 5),
    strong-scaled over 1/2/4/8 B200 (contiguous column shards, ONE all-reduce of the partial sums).

A "step" is one full pass of the path over the design: fused design-swap + model + estimator kernel on this
rank's shard, tile reduction, (N>1) all-reduce of nboot*nout*nstat doubles (one kernel per rank over NVLink peer memory that also writes the sums to pinned host memory; NCCL + a D2H copy as fallback)
and the host epilogue that yields S1/ST.

  value      inputs (A, B shards) resident in HBM before the timed region; CUDA events, max over ranks
  e2e        the same metric through the public API with HOST (pinned) A and B: the chunked H2D copy of the
             whole shard is inside the timed region, every step
  roofline   the fused kernel alone: 16*d bytes per sample / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline   the CPU restatement of the reference (oracle/gsa_ref.c, OpenMP, all host cores) on a bounded
             sample of the same workload (Julia is not in the image, so the reference itself cannot run)

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
    w = workload(args.log2n)
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
    partials = torch.empty((1, 1, ns), dtype=torch.float64, device="cuda")
    host_part = torch.empty(partials.shape, dtype=torch.float64).pin_memory()

    # the single collective of the path (a few KB): one kernel per rank over NVLink peer memory (csrc/peer.cu) that also
    # writes the sums into pinned host memory; torch.distributed's NCCL all-reduce + a D2H copy if the mailboxes cannot be mapped
    pg, collective = None, "none"
    if world > 1:
        collective = "nccl"
        if args.collective == "peer":
            try:
                pg = g.PeerGroup(h, partials.numel())
                okf = torch.ones(1, device="cuda")
            except Exception as e:                          # noqa: BLE001 - any rank failing sends every rank to NCCL
                sys.stderr.write(f"[bench rank {rank}] peer mailboxes unavailable ({e}); using NCCL\n")
                pg, okf = None, torch.zeros(1, device="cuda")
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if float(okf) < 1.0:
                pg = None
            else:
                collective = "peer-memory kernel (gsa_peer_allreduce)"

    def reduce_to_host():
        if pg is not None:
            pg.allreduce(partials, ns, host_out=host_part)
        else:
            if world > 1:
                dist.all_reduce(partials)
            host_part.copy_(partials, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return g.sobol_finalize(method, host_part.numpy(), d, N)

    def step():
        g.sobol_partials(model, method, A, B, N, col0, handle=h, out=partials)
        return reduce_to_host()

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
        kernel_ms.append(h.last_timing()[1])               # events already complete: the step synchronised
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
    gen = None
    if not args.no_gen:
        def gen_step():
            g.sobol_partials_generated(model, method, lb, ub, N, col0, ncols, handle=h, out=partials)
            return reduce_to_host()

        for _ in range(args.warmup):
            res_g = gen_step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            res_g = gen_step()
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gen = {"value": w["rows"] * args.steps / (float(tg[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(tg[0]) / args.steps,
               "hbm_bytes_read": 0, "note": "K1 fused into the product kernel: the Sobol points are generated in registers "
               "(bit-identical to the materialised design), gsa_sobol_partials_generated",
               "max_abs_diff_vs_resident": float(max(np.abs(res_g.S1 - res.S1).max(), np.abs(res_g.ST - res.ST).max()))}

    # ---- e2e: public API with HOST (pinned) inputs; H2D of the whole shard inside the timed region ---------------
    e2e = None
    if not args.no_e2e:
        cudart = torch.cuda.cudart()
        nbytes = 8 * d * ncols
        hostA = np.empty((ncols, d), dtype=np.float64)
        hostB = np.empty((ncols, d), dtype=np.float64)
        for arr in (hostA, hostB):
            rc = cudart.cudaHostRegister(arr.ctypes.data, nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed: {rc}")
        tA, tB = torch.from_numpy(hostA), torch.from_numpy(hostB)
        tA.copy_(A.t())                                     # fill the host copies (untimed)
        tB.copy_(B.t())
        torch.cuda.synchronize()
        del A, B
        torch.cuda.empty_cache()
        hA, hB = tA.t(), tB.t()                             # (d, ncols) views in Julia layout

        def e2e_step():
            if world == 1:
                return g.gsa(model, method, hA, hB, batch=True, handle=h)          # the call a user makes
            g.sobol_partials(model, method, hA, hB, N, col0, handle=h, out=partials)
            return reduce_to_host()

        for _ in range(args.e2e_warmup):
            res_e = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res_e = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": w["rows"] * args.e2e_steps / float(dt[0]), "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * d * N,
               "d2h_bytes_per_step": int(partials.numel() * 8 * world), "steps": args.e2e_steps, "warmup": args.e2e_warmup,
               "ms_per_step": 1e3 * float(dt[0]) / args.e2e_steps, "host_memory": "pinned (cudaHostRegister)",
               "max_abs_diff_vs_resident": float(max(np.abs(res_e.S1 - res.S1).max(), np.abs(res_e.ST - res.ST).max()))}
        for arr in (hostA, hostB):
            cudart.cudaHostUnregister(arr.ctypes.data)

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 16.0 * d * ncols                          # per launch on one GPU: A and B shards read once
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch from the committed ncu capture (measured at N=2^22, scaled by the algorithmic bytes)
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)["fused_product_kernel"]
            traffic, traffic_src = tr["dram_bytes_per_algorithmic_byte"] * alg_bytes, tr["source"]
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"bound": "hbm", "kernel": "fused_product_kernel<T,CPL> (design swap + Sobol-G + Jansen sums)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peak_src, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                    "per_rank_kernel_ms": per_rank_kernel_ms, "k1_design_ms": k1_ms, "k1_design_gbs": 16.0 * d * ncols / (k1_ms * 1e-3) / 1e9}
        cb = None if args.no_cpu else cpu_sample(args.cpu_log2n, 2)
        Vi = 1.0 / (3.0 * (1.0 + np.array(SOBOL_G_A)) ** 2)
        V = np.prod(1.0 + Vi) - 1.0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": w, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "collective": collective, "roofline": roofline, "generated_design": gen, "cpu_baseline": cb, "clocks": clocks,
                "check": {"S1_0": float(res.S1[0]), "S1_0_analytic": float(Vi[0] / V), "ST_0": float(res.ST[0]),
                          "ST_0_analytic": float(Vi[0] * np.prod(1 + Vi) / (1 + Vi[0]) / V)}}
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
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N > 1: how the partial sums are all-reduced")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
