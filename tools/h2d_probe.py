"""H2D bandwidth from (a) cudaHostAlloc'ed and (b) cudaHostRegister'ed host memory, at a size given in GiB."""
import sys, time, ctypes
import numpy as np, torch
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(gib * (1 << 30))
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
rt = torch.cuda.cudart()
def run(host_t, label):
    chunk = 1 << 30
    torch.cuda.synchronize(); t = time.perf_counter()
    for off in range(0, n, chunk):
        dev[:min(chunk, n - off)].copy_(host_t[off:off + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{label}: {n / dt / 1e9:.1f} GB/s over {gib} GiB", flush=True)
t0 = time.perf_counter()
a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
print("cudaHostAlloc(pin_memory) took %.1f s" % (time.perf_counter() - t0), flush=True)
a.fill_(1)
run(a, "cudaHostAlloc"); run(a, "cudaHostAlloc")
del a
t0 = time.perf_counter()
b = np.empty(n, dtype=np.uint8)
rc = rt.cudaHostRegister(b.ctypes.data, n, 0)
print("cudaHostRegister rc", int(rc), "took %.1f s" % (time.perf_counter() - t0), flush=True)
tb = torch.from_numpy(b); tb.fill_(1)
run(tb, "cudaHostRegister"); run(tb, "cudaHostRegister")
rt.cudaHostUnregister(b.ctypes.data)
cat = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
print("THP:", cat)
