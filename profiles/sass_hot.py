"""Hot SASS instructions of the first kernel of an ncu report: samples, executed count and dominant stall reasons.
    python profiles/sass_hot.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"]) for r in rows)
print("total samples", tot, "instructions", len(rows))
agg = {}
for c in stall_cols:
    agg[c] = sum(int(r[c] or 0) for r in rows)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"]))[:top]
for i in sorted(idx):
    r = rows[i]
    st = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r['# Samples']):7d} {100*int(r['# Samples'])/tot:5.1f}% exec {int(r['Instructions Executed']):9d}  {r['Source'].strip()[:70]:70s} {st}")
