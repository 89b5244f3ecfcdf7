#!/usr/bin/env python3
"""K3 diagnostics: one long pair alone (every block's warp has a scheduler to itself), per-block timing from
BSA_WAVE_TRACE: the lone-warp time per step and the hand-off lag per block.
usage: BSA_WAVE_TRACE=/tmp/t.csv python tools/wave_probe.py n m [n_pairs]"""
import csv
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(1)
AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)
seqs, off = [], [0]
for p in range(npairs):
    for L in (n, m):
        seqs.append(AA[rng.integers(0, 20, L)])
        off.append(off[-1] + L)
res = np.concatenate(seqs)
off = np.array(off, np.uint64)
trace = os.environ.get("BSA_WAVE_TRACE", "/tmp/wave_trace.csv")
os.environ["BSA_WAVE_TRACE"] = trace
with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    for rep in range(2):
        s, nid, paths = ctx.align_pairs_paths(0, 0, np.arange(0, 2 * npairs, 2), np.arange(1, 2 * npairs, 2))
rows = [{k: int(v) for k, v in r.items()} for r in csv.DictReader(open(trace))]
t0 = min(r["start_ns"] for r in rows)
byp = {}
for r in rows:
    byp.setdefault(r["pair"], []).append(r)
for p, lst in sorted(byp.items()):
    lst.sort(key=lambda r: r["block"])
    ends = np.array([(r["end_ns"] - t0) / 1e3 for r in lst])
    d = np.diff(ends)
    print("pair %d: n=%d m=%d blocks=%d | block 0: %d steps in %.0f us = %.4f us/step (%.0f cycles at 1.965 GHz) | "
          "lag per block: median %.2f us, mean %.2f us = %.1f steps | total %.0f us" %
          (p, lst[0]["n"], lst[0]["m"], len(lst), lst[0]["steps"], ends[0], ends[0] / lst[0]["steps"],
           ends[0] / lst[0]["steps"] * 1965, float(np.median(d)) if len(d) else 0, float(d.mean()) if len(d) else 0,
           (float(d.mean()) if len(d) else 0) / (ends[0] / lst[0]["steps"]), ends[-1]))
