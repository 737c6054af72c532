#!/usr/bin/env python
"""Tabulate an `ncu --csv --metrics ...` launch list: one line per launch, kernel names shortened.
    python scripts/launch_table.py gpurun_out/pt_wave_launches.csv [--skip N] [--count M]"""
import argparse
import csv
import re
from collections import OrderedDict

ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("--skip", type=int, default=0)
ap.add_argument("--count", type=int, default=10 ** 9)
args = ap.parse_args()
launches = OrderedDict()
for r in csv.reader(open(args.path, errors="replace")):
    if len(r) >= 15 and r[0].isdigit():
        launches.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))


def short(n):
    m = re.search(r"(\w+)<([^>]*)>?", n.replace("unnamed>::", ""))
    base = re.sub(r"\(.*", "", n.replace("void ", "").replace("unnamed>::", ""))
    base = re.sub(r"cbq::|<unnamed>::|\(anonymous namespace\)::", "", base)
    return base[:64]


total = {}
rows = [v for k, v in launches.items()][args.skip:args.skip + args.count]
print("%-66s %10s %12s %6s %6s" % ("kernel", "us", "warp instr", "lanes", "issue%"))
for v in rows:
    us = v.get("gpu__time_duration.sum", 0.0) / 1e3
    print("%-66s %10.1f %12.0f %6.2f %6.1f" % (short(v["name"]), us, v.get("smsp__inst_executed.sum", 0), v.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0),
                                               v.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)))
    key = short(v["name"]).split("<")[0]
    total[key] = total.get(key, 0.0) + us
print("--- totals (us):", {k: round(t, 1) for k, t in total.items()}, "sum", round(sum(total.values()), 1))
