#!/usr/bin/env python3
"""one-vs-many (BASELINE configs[3] shape) throughput probe: Q queries x D database sequences,
score only (16-bit packed lanes where the score range allows) and score+identity (32-bit)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, synth  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
import torch
qr, qo = synth.config("cfg4q", n=nq)
dr, do = synth.config("cfg4db", n=nd)
with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, qr, qo)
    ctx.load_sequences(1, dr, do)
    ds = torch.empty(nq * nd, dtype=torch.int32, device="cuda")
    dn = torch.empty(nq * nd, dtype=torch.int32, device="cuda")
    for want_i in (False, True):
        for rep in range(3):
            ctx.align_all_pairs(0, 1, None, want_identical=want_i, scores=ds.data_ptr(),
                                n_identical=dn.data_ptr() if want_i else None, device_out=True)
            st = ctx.stats()
        print("one-vs-many %d x %d %s: %.1f GCUPS (kernel %.1f ms, %d launches, swept/cells %.3f)" % (
            nq, nd, "score+identity (32-bit lanes)" if want_i else "score only (16-bit lanes)",
            st["cells"] / 1e6 / st["kernel_ms"], st["kernel_ms"], st["launches"], st["padded_cells"] / st["cells"]), flush=True)
