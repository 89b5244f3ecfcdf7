#!/usr/bin/env python3
"""BASELINE configs[4]: titin-scale pairs (5k-35k residues) through the K3 wavefront kernel with
full traceback.  16 pairs are timed; the largest pair (and one homologous pair) is checked glyph
for glyph against the CPU oracle (3 x 35001^2 bytes of trace on the host)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, synth  # noqa: E402
from bioshell_b200.scoring import ncbi_text  # noqa: E402
from oracle import c_oracle  # noqa: E402

res, off = synth.config("cfg5")      # 16 pairs (2p, 2p+1), U{5000..35000}, even pairs homologous, pair 0 = 34,350 x 35,000
lens = np.diff(off.astype(np.int64))
q = np.arange(0, 32, 2)
t = np.arange(1, 32, 2)
with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    all_ms = []
    for rep in range(int(os.environ.get("BSA_CFG5_REPS", "2"))):
        w0 = time.perf_counter()
        s, nid, paths = ctx.align_pairs_paths(0, 0, q, t)
        wall = time.perf_counter() - w0
        st = ctx.stats()
        all_ms.append(round(st["kernel_ms"], 3))
raw = res.tobytes()
sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
checked = []
for k in (() if os.environ.get("BSA_CFG5_NOCHECK") else (0, 1)):      # NOCHECK: timing only (A/B runs)
    a = raw[int(off[q[k]]):int(off[q[k] + 1])]
    b = raw[int(off[t[k]]):int(off[t[k] + 1])]
    t0 = time.perf_counter()
    ref = c_oracle.align_pair(a, b, sc, ai, -10, -1)
    ok = ref["score"] == s[k] and ref["n_identical"] == nid[k] and ref["path"] == paths[k].decode()
    checked.append({"len_q": len(a), "len_t": len(b), "score": int(s[k]), "n_identical": int(nid[k]),
                    "path_len": len(paths[k]), "bit_exact": bool(ok), "oracle_seconds": time.perf_counter() - t0})
print(json.dumps({"workload": "cfg5: 16 pairs, lengths %d..%d, full traceback" % (lens.min(), lens.max()),
                  "cells": st["cells"], "kernel_ms": st["kernel_ms"], "wall_ms": wall * 1e3,
                  "GCUPS": st["cells"] / 1e6 / st["kernel_ms"], "kernel_ms_all_reps": all_ms,
                  "checked_against_oracle": checked}))
sys.exit(0 if all(c["bit_exact"] for c in checked) else 1)
