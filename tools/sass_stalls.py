#!/usr/bin/env python3
"""Sum of the encoded stall counts (control bits 105..108) over a SASS address range: the minimum number of
cycles a lone warp needs to issue the range.  usage: sass_stalls.py file.sass 0xSTART 0xEND"""
import re
import sys
from collections import Counter

lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
L = open(sys.argv[1]).read().split("\n")
tot = 0
n = 0
byop = Counter()
cnt = Counter()
hist = Counter()
for i, l in enumerate(L):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", l)
    if not m or not (lo <= int(m.group(1), 16) <= hi):
        continue
    m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", L[i + 1])
    hiw = int(m2.group(1), 16)
    stall = (hiw >> 41) & 0xf
    ins = re.sub(r"^@!?U?P\d\s+", "", m.group(2))
    op = ins.split()[0].split(".")[0]
    tot += stall
    n += 1
    byop[op] += stall
    cnt[op] += 1
    hist[stall] += 1
print("instructions", n, "sum of stall counts", tot)
print("stall histogram", sorted(hist.items()))
for op, s in byop.most_common(12):
    print("  %-10s n=%4d stall sum=%5d avg=%.2f" % (op, cnt[op], s, s / cnt[op]))
