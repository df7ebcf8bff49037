#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / TMEM use (B200_PROFILING.md), from cuobjdump of
the built library.  Usage: python scripts/sass_summary.py [fp16|bf16] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dt = sys.argv[1] if len(sys.argv) > 1 else "fp16"
lib = os.path.join(ROOT, "demonet_b200", "lib", "libdemonet_b200_%s.so" % dt)
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FFMA2", "HFMA2",
             "F2FP", "MUFU.EX2", "LDS", "STS", "LDG", "STG", "ACQBULK", "UBLKCP"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))              # drop the argument list, keep template arguments
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                cur[mn] += 1
print("# SASS evidence per kernel, %s (cuobjdump -sass, sm_100a); columns = instruction counts" % os.path.basename(lib))
print("# UTCHMMA = tcgen05.mma kind::f16, UTMALDG/UTMASTG = TMA tensor load/store, LDTM = tcgen05.ld (TMEM -> registers),")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, F2FP = packed float -> fp16/bf16 conversion")
cols = [m for m in MNEMONICS if any(c[m] for c in counts.values())]
print("%-88s %7s " % ("kernel", "instrs") + " ".join("%8s" % c for c in cols))
fam = collections.OrderedDict()
for name, c in counts.items():
    base = re.sub(r"<.*", "", name).replace("dn::", "")
    f = fam.setdefault(base, {"variants": 0, "c": collections.Counter()})
    f["variants"] += 1
    for k, v in c.items():
        f["c"][k] = max(f["c"][k], v)
for base, f in fam.items():
    print("%-88s %7d " % ("%s (%d variants, max over variants)" % (base, f["variants"]), f["c"]["_total"]) +
          " ".join("%8d" % f["c"][c] for c in cols))
