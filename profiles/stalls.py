"""Top stalled SASS instructions of an ncu report (source page). usage: python profiles/stalls.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {k: hdr.index(k) for k in ["Source", "# Samples", "stall_long_sb", "stall_wait", "stall_short_sb", "Instructions Executed",
                                "stall_math", "stall_not_selected", "stall_selected", "stall_barrier", "stall_mio"]}
data = []
for r in rows[2:]:
    try:
        data.append({k: (r[i] if k == "Source" else int(r[i])) for k, i in ix.items()})
    except Exception:
        pass
tot = sum(d["# Samples"] for d in data)
print("kernel:", rows[0][1][:80])
print("total samples", tot, {k: sum(d[k] for d in data) for k in ix if k.startswith("stall")})
for d in sorted(data, key=lambda d: -d["# Samples"])[:n]:
    print(f"{d['# Samples']:6d} {100 * d['# Samples'] / tot:5.1f}% long={d['stall_long_sb']:5d} wait={d['stall_wait']:5d} "
          f"short={d['stall_short_sb']:5d} math={d['stall_math']:4d} ex={d['Instructions Executed']:8d}  {d['Source'][:84]}")
