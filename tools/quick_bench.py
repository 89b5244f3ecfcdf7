#!/usr/bin/env python3
"""Device-leg-only throughput probe for kernel tuning (not the contract bench).
BSA_PROFILE_GROUPS=1 makes the library time every columns-per-lane group separately."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(synth.CONFIGS["cfg2"]); cfg["n"] = n
res, off = synth.generate(**cfg)
import torch
with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    counts = np.arange(n, dtype=np.uint32)
    npairs = n * (n - 1) // 2
    ds = torch.empty(npairs, dtype=torch.int32, device="cuda")
    dn = torch.empty(npairs, dtype=torch.int32, device="cuda")
    for r in range(reps + 1):
        t0 = time.perf_counter()
        ctx.align_all_pairs(0, 0, counts, scores=ds.data_ptr(), n_identical=dn.data_ptr(), device_out=True)
        st = ctx.stats()
        if r:
            print("rep %d: %.1f GCUPS (kernel %.1f ms, wall %.1f ms, swept/cells %.3f, items %d)" % (
                r, st["cells"] / 1e6 / st["kernel_ms"], st["kernel_ms"], (time.perf_counter() - t0) * 1e3,
                st["padded_cells"] / st["cells"], st["items"]), flush=True)
