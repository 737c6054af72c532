#!/usr/bin/env python
"""Per-instruction executed counts from an ncu report (SASS view), as a compact listing:
   python scripts/sass_profile.py report.ncu-rep [min_count]  -> offset  warp-instr executed  avg lanes  samples  SASS"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
head = rows[1]
ia, isrc, iex, ith, ismp = head.index("Address"), head.index("Source"), head.index("Instructions Executed"), head.index("Avg. Threads Executed"), head.index("# Samples")
base = None
total = 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    ex = int(r[iex]); total += ex
    print("%05x %10d %5s %6s  %s" % (a - base, ex, r[ith], r[ismp], r[isrc].strip()))
print("total", total)
