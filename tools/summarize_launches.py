"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((int(r["ID"]), name, ns))
tot = sum(r[2] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for _, n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"{len(rows)} launches, total {tot/1e6:.3f} ms (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'n':>6s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:6d} {ns/1e6:10.3f} {ns/c/1e3:10.1f} {100*ns/tot:6.1f}%")
