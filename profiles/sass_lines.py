"""Stall samples of the first kernel in an ncu report, aggregated by the source line of the kernel file (inlined helpers
are charged to the line that calls them): joins ncu's SASS page with `nvdisasm -gi` of the object the report was taken
from (same build!).    python profiles/sass_lines.py report.ncu-rep scenedino_b200/build/field_bin.o field_bin.cu [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, fname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
kname = lines[0].split('","')[1].split("(")[0].split("::")[-1]
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
# walk the disassembly of that kernel: current line of `fname` (outermost frame of the inline chain)
line_of = []
cur, in_k, pending = None, False, None
for l in dis.splitlines():
    if l.startswith("\t.section") or l.startswith("//-----"):
        in_k = kname in l
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if os.path.basename(m.group(1)) == fname and "inlined at" not in m.group(3):
            cur = int(m.group(2))
        mm = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        for f, n in mm:
            if os.path.basename(f) == fname:
                cur = int(n)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        line_of.append(cur)
print(f"kernel {kname}: {len(rows)} SASS rows in the report, {len(line_of)} in the object")
n = min(len(rows), len(line_of))
agg, ex = {}, {}
for i in range(n):
    agg[line_of[i]] = agg.get(line_of[i], 0) + int(rows[i]["# Samples"])
    ex[line_of[i]] = max(ex.get(line_of[i], 0), int(rows[i]["Instructions Executed"]))
tot = sum(agg.values())
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", fname)).read().splitlines()
print("total samples", tot)
for ln, s in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    text = src[ln - 1].strip()[:100] if ln and ln <= len(src) else ""
    print(f"{s:7d} {100 * s / tot:5.1f}%  exec {ex[ln]:9d}  L{ln}: {text}")
