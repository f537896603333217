"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel (and per
grid size for the big kernels). Usage: python profiles/summarize_launches.py gpurun_out/launches_r1.csv"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"].split("(")[0]
        rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), ns))
    total = sum(r[3] for r in rows)
    by = defaultdict(lambda: [0, 0.0])
    for n, g, b, ns in rows:
        by[n][0] += 1
        by[n][1] += ns
    print(f"launches: {len(rows)}  total kernel time: {total / 1e6:.3f} ms")
    print(f"{'kernel':70s} {'n':>5s} {'ms':>9s} {'share':>7s}")
    for n, (c, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:5d} {ns / 1e6:9.3f} {100 * ns / total:6.1f}%")
    print()
    big = defaultdict(lambda: [0, 0.0])
    for n, g, b, ns in rows:
        if "gemm_conv" in n or "attn_" in n:
            big[(n[:40], g)][0] += 1
            big[(n[:40], g)][1] += ns
    print("largest (kernel, grid) groups:")
    for (n, g), (c, ns) in sorted(big.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  {n:40s} grid {g:18s} x{c:3d}  {ns / 1e6:8.3f} ms  ({ns / c / 1e3:8.1f} us each)")


if __name__ == "__main__":
    main(sys.argv[1])
