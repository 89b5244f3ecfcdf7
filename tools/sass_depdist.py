#!/usr/bin/env python3
"""Dependency distances inside a SASS address range: for every instruction, how many instructions before it
its youngest register producer sits (in-order issue: < 4-5 means a `wait` stall for a lone warp).
usage: sass_depdist.py file.sass 0xSTART 0xEND"""
import re
import sys
from collections import Counter

lines = []
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
for l in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m and lo <= int(m.group(1), 16) <= hi:
        lines.append(m.group(2))
last = {}
hist = Counter()
stall = 0
for k, ins in enumerate(lines):
    ins2 = re.sub(r"^@!?U?P\d\s+", "", ins)
    parts = ins2.split(None, 1)
    op = parts[0]
    regs = re.findall(r"\bR(\d+)\b", parts[1]) if len(parts) > 1 else []
    wide = 4 if ".128" in op else (2 if ".64" in op or "WIDE" in op else 1)
    if op.startswith(("ST", "BRA", "ISETP", "NANOSLEEP", "BSYNC", "BSSY", "WARPSYNC")):
        dst, src = [], regs
    else:
        dst, src = regs[:1], regs[1:]
    d = min((k - last[r] for r in src if r in last), default=99)
    hist[min(d, 12)] += 1
    lat = 4
    if d < lat:
        stall += lat - d
    for r in dst:
        for w in range(wide):
            last[str(int(r) + w)] = k
print("instructions", len(lines), "est. wait stall cycles (lat 4, 1 issue/cycle)", stall)
print(sorted(hist.items()))
