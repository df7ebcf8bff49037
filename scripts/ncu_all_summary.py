#!/usr/bin/env python
"""Per-launch table from `ncu --csv --page raw` (wide format: one row per launch, one column per metric).
usage: ncu_all_summary.py ncu_all.csv"""
import csv
import re
import sys

lines = [x for x in open(sys.argv[1]) if not x.startswith("==")]
rows = list(csv.reader(lines))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def col(r, name, default=""):
    i = ix.get(name)
    return r[i] if i is not None and i < len(r) else default


def f(r, name):
    try:
        return float(col(r, name).replace(",", ""))
    except Exception:
        return float("nan")


print("%-4s %-66s %9s %6s %6s %6s %6s %6s %5s %5s %6s %6s %8s" % ("#", "kernel", "us", "dram%", "sm%", "issue%", "occ%", "thocc%", "regs", "ctas", "l2hit%", "l1hit%", "dramMB"))
tot = 0.0
for n, r in enumerate(body):
    name = col(r, "Kernel Name")
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dn::", "").replace("(int)", "").replace("(bool)", "")[:66]
    us = f(r, "gpu__time_duration.sum")
    unit = units[ix["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in ix else "ns"
    us = us / 1e3 if unit in ("ns", "nsecond") else (us * 1e3 if unit in ("ms", "msecond") else us)
    tot += us
    mb = (f(r, "dram__bytes_read.sum") * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(units[ix["dram__bytes_read.sum"]], 1e-6)
          + f(r, "dram__bytes_write.sum") * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(units[ix["dram__bytes_write.sum"]], 1e-6)) \
        if "dram__bytes_read.sum" in ix else float("nan")
    print("%-4d %-66s %9.1f %6.1f %6.1f %6.1f %6.1f %6.1f %5.0f %5.0f %6.1f %6.1f %8.1f" % (
        n, name, us, f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        f(r, "smsp__issue_active.avg.pct"), f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        f(r, "sm__maximum_warps_per_active_cycle_pct"), f(r, "launch__registers_per_thread"), f(r, "launch__grid_size"),
        f(r, "lts__t_sector_hit_rate.pct"), f(r, "l1tex__t_sector_hit_rate.pct"), mb))
print("total %.1f us over %d launches (ncu: serialised, cold-ish caches, base clocks unless --clock-control none)" % (tot, len(body)))
