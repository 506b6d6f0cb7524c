"""profiles/r2_traffic.json from an ncu report of bench.py's headline kernel: DRAM bytes per algorithmic byte of the SHIPPED build
(the sha256 of libgsa_cuda.so is recorded; bench.py only quotes the figure when it runs the same binary).
    python tools/make_traffic_json.py gpurun_out/r2_prod.ncu-rep <log2n> > profiles/r2_traffic.json"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rep, log2n = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
best = None
for r in rows[2:]:
    if "fused_product_kernel" not in r[idx["Kernel Name"]]:
        continue
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    best = (r[idx["Kernel Name"]], rd, wr)
alg = 16.0 * 100 * (1 << log2n)
with open(os.path.join(ROOT, "globalsensitivity.jl_b200", "libgsa_cuda.so"), "rb") as f:
    sha = hashlib.sha256(f.read()).hexdigest()[:16]
name, rd, wr = best
# the kernel's own identity: its mangled name and a hash of its SASS (bench.py falls back to this when the library was rebuilt)
sys.path.insert(0, ROOT)
from bench import kernel_sass_hash
LIB = os.path.join(ROOT, "globalsensitivity.jl_b200", "libgsa_cuda.so")
syms = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
import re
def norm(n):   # ncu prints "void f<8, 14, 1>(ProductArgs)", c++filt "void gsa::f<8, 14, true>(gsa::ProductArgs)"
    return n.replace("gsa::", "").replace("true", "1").replace("false", "0").replace(" ", "")
mangled = ""
for m in sorted(set(re.findall(r"_ZN3gsa\d+fused_product_kernel\w+", syms))):
    if norm(subprocess.run(["c++filt", m], capture_output=True, text=True).stdout.strip()) == norm(name):
        mangled = m
        break
print(json.dumps({"fused_product_kernel": {
    "dram_bytes_per_algorithmic_byte": (rd + wr) / alg, "lib_sha256_16": sha, "kernel": name, "kernel_mangled": mangled,
    "kernel_sass_sha256_16": kernel_sass_hash(LIB, mangled) if mangled else None, "log2n": log2n,
    "source": f"ncu --set full --clock-control none, bench.py --log2n {log2n} (one launch of {name[:60]}): dram__bytes_read.sum "
              f"{rd / 1e9:.6f} GB + dram__bytes_write.sum {wr / 1e9:.6f} GB for {alg / 1e9:.6f} GB algorithmic; {os.path.basename(rep)}"}}, indent=1))
