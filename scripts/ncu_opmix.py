#!/usr/bin/env python
"""Opcode mix + hottest SASS instructions from `ncu --page source --csv` output.
Usage: ncu -i rep --page source --csv > src.csv ; ncu_opmix.py src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samples = Counter(), Counter()
tot = 0
recs = []
for r in rows[2:]:
    if len(r) <= iex or not r[iex].strip().isdigit():
        continue
    ex = int(r[iex]); sm = int(r[ismp]) if r[ismp].strip().isdigit() else 0
    op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
    op = op.split(".")[0]
    ops[op] += ex; samples[op] += sm; tot += ex
    recs.append((ex, sm, r[isrc]))
print("total warp instructions executed: %d, static SASS instructions: %d" % (tot, len(recs)))
st = sum(samples.values()) or 1
for op, c in ops.most_common(top):
    print("  %-10s %6.2f%% of instr   %6.2f%% of stall samples" % (op, 100.0 * c / tot, 100.0 * samples[op] / st))
