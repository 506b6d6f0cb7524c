"""SASS evidence for profiles/ (VERDICT r1 missing #6): disassembles libgsa_cuda.so with `cuobjdump -sass` and writes, per
hot kernel, the counts of the mnemonics that matter on this path (FP64 pipe: DFMA/DMUL/DADD; FP64 tensor-core route: DMMA; streaming
loads: LDG.E.*; cp.async: LDGSTS; TMA/tcgen05: UBLKCP/UTMA*/UTC*MMA -- expected ABSENT, the path has no tcgen05-eligible contraction)
plus a short excerpt of the inner loop.  Runs on the build machine (no GPU needed):
    python tools/sass_evidence.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SO = os.path.join(ROOT, "globalsensitivity.jl_b200", "libgsa_cuda.so")
WANT = [("fused_product_kernel<8, 13, false, true>", "config 5 (8-byte loads)"), ("fused_product_kernel_v2<8, 14, true>", "config 5 (16-byte pair loads)"),
        ("fused_product_kernel<8, 13, true, true>", "config 5, generated design"), ("pairs_mma_kernel<2, 1, 8, 3, true>", "config 3 (tile source)"),
        ("pairs_mma_kernel<0, 3, 2, 3, true>", "Sobol G d=20 order 2"), ("pairs_mma_kernel<1, 3, 16, 4, true>", "estimator on Y, d=20 order 2"),
        ("linear_gram_kernel<5, 2, 6>", "config 4"), ("small_kernel<gsa::IshigamiSource<3, true>, 3, true, true>", "config 2"),
        ("small_kernel<gsa::IshigamiSource<3, false>, 3, false, true>", "config 1"), ("sobol_design_kernel2", "K1"),
        ("estimator_y_kernel<false>", "estimator on Y, d=100"),
        ("fused_generic_kernel<gsa::SobolGModel, false, true>", "generic kernel (large instantiation)"), ("peer_allreduce_kernel", "K5 stand-alone"),
        ("reduce_tiles_kernel", "stage 2 stand-alone")]
KEYS = ["DFMA", "DMUL", "DADD", "DMMA", "DSETP", "MUFU", "LDG", "LDGSTS", "LDS", "STS", "STG", "SHFL", "ATOMG", "RED", "BAR", "UBLKCP", "UTMALDG", "UTCQMMA", "UTCHMMA", "HMMA", "IMMA"]
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs, cur, name = {}, None, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = funcs.setdefault(name, [])
        continue
    if cur is not None and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        cur.append(line.strip())
print("# cuobjdump -sass of", os.path.relpath(SO, ROOT), "(sm_100a) -- instruction counts per kernel; 0 of UBLKCP/UTMA*/UTC*MMA/HMMA by design (fp64 path)")
for pat, what in WANT:
    hits = [n for n in funcs if pat in n]
    if not hits:
        print(f"\n## {pat}: NOT FOUND")
        continue
    n = hits[0]
    ins = funcs[n]
    cnt = collections.Counter()
    ldg = collections.Counter()
    for l in ins:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", l)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k:
                cnt[k] += 1
        if op == "LDG":
            ldg["LDG" + m.group(2)] += 1
    print(f"\n## {n[:140]}   [{what}]")
    print("   instructions:", len(ins), " ", "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k] or k in ("DFMA", "DMMA", "UBLKCP", "UTCQMMA", "HMMA")))
    if ldg:
        print("   loads:", "  ".join(f"{k}={v}" for k, v in ldg.most_common(4)))
    # excerpt: the first 12 FP64 / DMMA / load instructions in a row-dense region
    idx = [i for i, l in enumerate(ins) if re.search(r"\b(DMMA|DFMA|DMUL)\b", l)]
    if idx:
        mid = idx[len(idx) // 2]
        for l in ins[max(0, mid - 4):mid + 8]:
            print("      " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l)[:150])
    # loops that carry DMMA: instruction mix per trip (the FP64-pipe work per warp step quoted in DESIGN.md / bench.py comes from here)
    if cnt["DMMA"]:
        addr = []
        for l in ins:
            m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m:
                addr.append((int(m.group(1), 16), m.group(2)))
        for a, t in addr:
            m2 = re.search(r"BRA\S*\s.*0x([0-9a-f]+)", t)
            if not m2 or int(m2.group(1), 16) >= a:
                continue
            lo = int(m2.group(1), 16)
            mix = collections.Counter()
            for b, u in addr:
                if lo <= b <= a:
                    w = u.split()
                    op = (w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0]
                    mix[op] += 1
            if mix["DMMA"] and sum(mix.values()) < 1300:
                print(f"   loop {lo:#x}..{a:#x}: {sum(mix.values())} instructions  " + "  ".join(f"{k}={v}" for k, v in mix.most_common(12)))
