#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
per kernel: launches, device time, share of the step, DRAM traffic.  usage: ncu_launch_summary.py launches.csv [out.json]"""
import collections
import csv
import json
import sys

lines = [x for x in open(sys.argv[1]) if not x.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    name = name.split("<")[0]
    m, u = row["Metric Name"], row["Metric Unit"]
    v = float(row["Metric Value"].replace(",", ""))
    a = agg.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    if m == "gpu__time_duration.sum":
        a["time_us"] += v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a["launches"] += 1
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        a["dram_read_bytes" if "read" in m else "dram_write_bytes"] += v
tot = sum(a["time_us"] for a in agg.values())
out = {"total_time_us": tot, "launches": sum(a["launches"] for a in agg.values()), "kernels": {}}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]):
    traffic = a["dram_read_bytes"] + a["dram_write_bytes"]
    out["kernels"][k] = {"launches": a["launches"], "time_us": round(a["time_us"], 1), "share": round(a["time_us"] / tot, 4),
                         "dram_bytes": int(traffic), "dram_bytes_per_launch": int(traffic / max(a["launches"], 1)),
                         "dram_GBps": round(traffic / a["time_us"] / 1e3, 1) if a["time_us"] else 0}
    print("%-28s n=%3d %9.1f us %5.1f%%  dram %9.1f MB  %7.1f GB/s" % (k[:28], a["launches"], a["time_us"], 100 * a["time_us"] / tot,
                                                                     traffic / 1e6, traffic / a["time_us"] / 1e3 if a["time_us"] else 0))
print("launches %d, total %.1f us (ncu times are cold-cache and serialised: compare shares, not absolutes)" % (out["launches"], tot))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
