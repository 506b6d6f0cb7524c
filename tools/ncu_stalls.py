"""Summarise the per-instruction stall samples of an ncu `--page source --csv` dump, per kernel.
    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys


def main(path, top=12):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    for x in rows:
        if x and x[0] == "Kernel Name":
            cur = {"name": x[1], "hdr": None, "rows": []}
            sections.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = x
        elif cur is not None and len(x) == len(cur["hdr"]):
            cur["rows"].append(x)
    for s in sections:
        idx = {h: i for i, h in enumerate(s["hdr"])}
        stalls = [h for h in s["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        tot = {k: sum(int(x[idx[k]] or 0) for x in s["rows"]) for k in stalls}
        allsamp = sum(tot.values()) or 1
        print("==", s["name"][:100], "instructions:", len(s["rows"]))
        print("   ", {k: f"{100 * v / allsamp:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > allsamp})
        rs = sorted(s["rows"], key=lambda x: -int(x[idx["# Samples"]] or 0))
        for x in rs[:top]:
            why = {k[6:]: x[idx[k]] for k in stalls if int(x[idx[k]] or 0) * 10 > int(x[idx["# Samples"]] or 1)}
            print(f"   {x[idx['# Samples']]:>7} {x[idx['Source']].strip()[:70]:70s} {why}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
