"""Tabulates the key metrics of every kernel in an ncu report: python tools/ncu_summary.py prof.ncu-rep > summary.txt"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"), ("launch__occupancy_limit_registers", "occ_regs"), ("launch__occupancy_limit_shared_mem", "occ_smem")]
print("# ncu --set full --clock-control none --import-source on; source:", rep)
for r in rows[2:]:
    print("\n##", r[idx["Kernel Name"]][:150])
    out = []
    for k, short in keys:
        if k in idx:
            out.append(f"{short}={r[idx[k]]}{(' ' + units[idx[k]]) if units[idx[k]] not in ('', '%') else ''}")
    print("   " + "  ".join(out))
    try:
        t = float(r[idx["gpu__time_duration.sum"]]); u = units[idx["gpu__time_duration.sum"]]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[u if u in ("ns", "us", "ms", "s") else "us"]
        def gb(key):
            v = float(r[idx[key]]); un = units[idx[key]]
            return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[un]
        print(f"   DRAM read {gb('dram__bytes_read.sum') / t_s:.0f} GB/s, write {gb('dram__bytes_write.sum') / t_s:.0f} GB/s (under ncu)")
    except Exception:
        pass
