#!/usr/bin/env python
"""Attribute an ncu SASS profile to CUDA source lines (ncu's own cuda view carries no
metrics for header-resident device code).
Usage: ncu -i rep --page source --csv > src.csv
       ncu_lines.py src.csv path/to/object.o kernel_substring [top]
Joins by instruction order: the n-th SASS instruction of the kernel in `nvdisasm -g`
output carries the //## File/line marker that precedes it."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

src_csv, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, infn = [], ("?", 0), False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
recs = [r for r in rows[2:] if len(r) > iex and r[iex].strip().isdigit()]
print("sass instructions: profile %d, disassembly %d" % (len(recs), len(lines)))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = st = 0
for i, r in enumerate(recs):
    key = lines[i] if i < len(lines) else ("?", 0)
    ex = int(r[iex]); sm = int(r[ismp]) if r[ismp].strip().isdigit() else 0
    a = agg[key]; a[0] += ex; a[1] += sm
    op = (r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]).split(".")[0]
    a[2][op] += ex
    tot += ex; st += sm
for key, (ex, sm, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    opsum = " ".join("%s:%.0f%%" % (o, 100.0 * c / ex) for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print("%-22s %6.2f%% instr %6.2f%% samples   %s" % ("%s:%d" % key, 100.0 * ex / tot, 100.0 * sm / max(st, 1), opsum))
