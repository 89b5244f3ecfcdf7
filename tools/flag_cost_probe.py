#!/usr/bin/env python3
"""How much does the end-of-query (flagged) path of the streaming kernels cost?  One-vs-many, score+identity,
fixed template length, the SAME number of stream residues cut into queries of length L: the shorter the
queries, the larger the share of double steps in which some lane passes a flag."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context  # noqa: E402
import torch  # noqa: E402

rng = np.random.default_rng(5)
AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)
def make(n, L):
    res = AA[rng.integers(0, 20, n * L)]
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    return res, off
with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    for m in (544, 320):
        tres, toff = make(600, m)
        ctx.load_sequences(1, tres, toff)
        for L in (64, 128, 288, 576, 1152, 3456):
            nq = 3456 * 100 // L
            qres, qoff = make(nq, L)
            ctx.load_sequences(0, qres, qoff)
            ds = torch.empty(nq * 600, dtype=torch.int32, device="cuda")
            dn = torch.empty(nq * 600, dtype=torch.int32, device="cuda")
            for rep in range(3):
                ctx.align_all_pairs(0, 1, None, scores=ds.data_ptr(), n_identical=dn.data_ptr(), device_out=True)
                st = ctx.stats()
            print("template %d cols, query length %5d: %.1f GCUPS (kernel %.2f ms, swept/cells %.3f)" % (
                m, L, st["cells"] / 1e6 / st["kernel_ms"], st["kernel_ms"], st["padded_cells"] / st["cells"]), flush=True)
