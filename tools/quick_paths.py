#!/usr/bin/env python3
"""Timing probe for the direction-store + traceback path (K2/K3)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, synth  # noqa: E402

with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    res, off = synth.config("cfg2", n=4000)
    ctx.load_sequences(0, res, off)
    rng = np.random.default_rng(1)
    for npairs in (2000, 20000):
        q = rng.integers(0, 4000, npairs); t = rng.integers(0, 4000, npairs)
        for rep in range(2):
            t0 = time.perf_counter()
            s, nid, paths = ctx.align_pairs_paths(0, 0, q, t)
            st = ctx.stats()
        print("K2 %d pairs: %.1f ms wall, kernel %.1f ms, %.1f GCUPS (kernel), launches %d" % (
            npairs, (time.perf_counter() - t0) * 1e3, st["kernel_ms"], st["cells"] / 1e6 / st["kernel_ms"], st["launches"]), flush=True)
    res, off = synth.config("cfg5")
    ctx.load_sequences(1, res, off)
    lens = np.diff(off.astype(np.int64))
    q = np.arange(0, 32, 2); t = np.arange(1, 32, 2)
    for rep in range(2):
        t0 = time.perf_counter()
        s, nid, paths = ctx.align_pairs_paths(1, 1, q, t)
        st = ctx.stats()
    print("K3 16 titin-scale pairs (%d..%d): %.1f ms wall, kernel %.1f ms, %.1f GCUPS" % (
        lens.min(), lens.max(), (time.perf_counter() - t0) * 1e3, st["kernel_ms"], st["cells"] / 1e6 / st["kernel_ms"]), flush=True)
