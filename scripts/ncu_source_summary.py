#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix, executed-instruction total, hottest
SASS lines by stall samples.  usage: ncu_source_summary.py file.source.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > ix["Instructions Executed"]]
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:100])
print("SASS lines %d, warp instructions executed %d, stall samples %d" % (len(body), tot_inst, tot_samp))
mix = collections.Counter()
for r in body:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    mix[op] += int(r[ix["Instructions Executed"]] or 0)
print("instruction mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot_inst, 1)) for k, v in mix.most_common(14)))
print("hottest lines (samples, executed, SASS):")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
    print("  %6s %10s  %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:110]))
