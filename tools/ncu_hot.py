"""Top stalled SASS instructions of one kernel launch from an .ncu-rep source page.
   python tools/ncu_hot.py rep.ncu-rep <kernel-regex> [launch-skip] [top-n]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print(lines[0][:150])
print(f"total samples {tot}, {len(rows)} SASS instructions")
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
for i, r in enumerate(rows):
    r["_i"] = i
hot = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:topn]
for r in sorted(hot, key=lambda r: r["_i"]):
    n = int(r["# Samples"] or 0)
    st = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{r['_i']:5d} {100*n/tot:5.1f}%  {r['Source'].strip()[:70]:70s} {st}")
