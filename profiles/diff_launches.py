"""Compare two ncu launch lists of the same step (e.g. a fused-epilogue option on / off): total per kernel and the
GEMM launches paired in order.  usage: python profiles/diff_launches.py A.csv B.csv"""
import csv
import sys
from collections import defaultdict


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
        rows.append((r["Kernel Name"].split("(")[0].replace("void mvd::", "").replace("mvd::", ""), r["Grid Size"], ns / 1e3))
    return rows


a, b = load(sys.argv[1]), load(sys.argv[2])
for name, rows in (("A", a), ("B", b)):
    by = defaultdict(float)
    for k, _, us in rows:
        by[k.split("<")[0]] += us
    print(f"{name}: {len(rows)} launches, {sum(r[2] for r in rows) / 1e3:.3f} ms  " +
          "  ".join(f"{k} {v / 1e3:.3f}" for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:7]))
ga, gb = [x for x in a if "gemm" in x[0]], [x for x in b if "gemm" in x[0]]
if len(ga) == len(gb):
    d = defaultdict(lambda: [0, 0.0, 0.0])
    for x, y in zip(ga, gb):
        key = (y[0], y[1], round(y[2] / 5) * 5)
        d[key][0] += 1
        d[key][1] += x[2]
        d[key][2] += y[2]
    for k, (n, fa, fb) in sorted(d.items(), key=lambda kv: -abs(kv[1][1] - kv[1][2])):
        if abs(fa - fb) > 3:
            print(f"  {k[0]:26s} grid {k[1]:14s} ~{k[2]:4d} us x{n:3d}: A {fa:8.1f}  B {fb:8.1f}  A-B {fa - fb:+7.1f} us")
    print(f"  GEMM total A-B: {sum(v[1] - v[2] for v in d.values()):+.1f} us")
