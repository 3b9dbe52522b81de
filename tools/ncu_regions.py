#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump of the conv kernel: executed warp instructions and stall samples per
code region (2 KB buckets, labelled by the role markers they contain) and the top stall reasons.
    ncu -i rep.ncu-rep --page source --csv > src.csv;  python tools/ncu_regions.py src.csv [units_per_cta]
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
a0 = int(data[0][ix["Address"]], 16)
marks = {"LDTM": "epilogue", "F2FP": "gather", "STTM.x32": "build", "UTCHMMA": "issue", "UBLKCP": "producer"}
b, sm, lab = collections.Counter(), collections.Counter(), collections.defaultdict(set)
stall = collections.Counter()
for r in data:
    a = (int(r[ix["Address"]], 16) - a0) // 0x800
    b[a] += int(r[ix["Instructions Executed"]])
    sm[a] += int(r[ix["# Samples"]])
    for m, name in marks.items():
        if re.search(m, r[ix["Source"]]):
            lab[a].add(name)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ix[h]])
tot, stot = sum(b.values()), sum(sm.values())
print("kernel:", rows[0][1][:80], " code bytes:", int(data[-1][ix["Address"]], 16) - a0)
for k in sorted(b):
    if b[k] > tot * 0.004 or sm[k] > stot * 0.004:
        print("%6s  instr %10d %5.1f%%   samples %6d %5.1f%%   %s" % (hex(k * 0x800), b[k], 100.0 * b[k] / tot, sm[k],
                                                                      100.0 * sm[k] / stot, ",".join(sorted(lab[k]))))
print("total warp instructions", tot)
if len(sys.argv) > 2:
    print("per SMSP per unit: %.1f" % (tot / (148 * 4 * float(sys.argv[2]))))
print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * n / max(1, sum(stall.values()))) for h, n in stall.most_common(8)))
