"""Provenance of tests/golden/reference_literals.json: re-reads the numeric literals at the cited line ranges of the reference's
own test file and checks (or, with --write, refreshes) the fixture.  Runs only where /root/reference exists (the build
container); the GPU box and the tests use the committed JSON.

    python tests/golden/make_reference_literals.py [--write] [--reference /root/reference]
"""
import argparse
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
FLOAT = re.compile(r"(?<![\w.])[-+]?(?:\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)")


def literals(path, first, last):
    with open(path) as f:
        lines = f.readlines()[first - 1:last]
    txt = "".join(l.split("#")[0] for l in lines)
    txt = re.sub(r"atol\s*=\s*[-+\deE.]+", "", txt)          # tolerances are not data
    return [float(x) for x in FLOAT.findall(txt)]


def flatten(v):
    if isinstance(v, (list, tuple)):
        for x in v:
            yield from flatten(x)
    else:
        yield float(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    fixture = os.path.join(HERE, "reference_literals.json")
    data = json.load(open(fixture))
    bad = 0
    for key, entry in data.items():
        if not isinstance(entry, dict) or "src" not in entry:
            continue
        rel, rng = entry["src"].rsplit(":", 1)
        first, last = (int(x) for x in rng.split("-"))
        got = literals(os.path.join(args.reference, rel), first, last)
        want = list(flatten(entry["value"]))
        # the fixture may hold a subset (e.g. only the numbers of one vector inside the cited lines): every fixture number must
        # appear, in order, among the literals of the cited lines
        it = iter(got)
        ok = all(any(w == g for g in it) for w in want)
        print(f"{key:34s} {entry['src']:28s} {len(want):3d} numbers  {'ok' if ok else 'MISMATCH'}")
        bad += not ok
    if bad:
        raise SystemExit(f"{bad} fixture entries do not match the reference's literals")
    if args.write:
        json.dump(data, open(fixture, "w"), indent=1)


if __name__ == "__main__":
    main()
