#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): evidence that the tensor-core kernels really
issue tcgen05 (UTCHMMA / UTCBAR), TMEM loads/stores (LDTM / STTM), bulk copies (UBLKCP) and mbarrier waits (SYNCS).
    python tools/sass_histogram.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "3d-wsis_b200", "wsis_b200", "libwsis_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS",
       "F2FP", "FFMA", "HMMA", "ATOMG", "RED", "BAR", "ELECT"]
print("# SASS opcode histogram per kernel (`cuobjdump -sass %s`)\n" % os.path.relpath(so, ROOT))
print("UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk, "
      "SYNCS = mbarrier ops, LDGSTS = cp.async, ELECT = elect.sync.\n")
print("| kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for k, h in hist.items():
    if sum(h.values()) < 40:
        continue
    print("| `%s` | %d | " % (k[:70], sum(h.values())) + " | ".join(str(h.get(o, 0) or "") for o in KEY) + " |")
