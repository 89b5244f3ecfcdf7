#!/usr/bin/env python3
"""Integer / DPX pipe microbenchmark on the GPU (bsa_measure_int_peak): the measured
denominators of the integer roofline.  Writes JSON to stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context  # noqa: E402

NAMES = {0: "cell_mix_8op", 1: "VIADDMNMX", 2: "VIMNMX3+LOP3", 3: "LOP3", 4: "IADD3", 5: "IMAD",
         6: "VIADDMNMX.S16x2"}

with Context(0) as ctx:
    out = {}
    for which, name in NAMES.items():
        best = 0.0
        mhz = 0.0
        for _ in range(3):
            ops, m = ctx.measure_int_peak(which)
            if ops > best:
                best, mhz = ops, m
        out[name] = {"lane_ops_per_s": best, "sm_mhz": mhz,
                     "lane_ops_per_clk_per_sm": best / (mhz * 1e6) / 148 if mhz else None}
    print(json.dumps(out, indent=1))
