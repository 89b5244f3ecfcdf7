#!/usr/bin/env python3
"""Hierarchical clustering at the sizes the identity matrix of BASELINE configs[1..2] would have (SURVEY.md
8f rank 2: "unusable at 100 k" for the reference): n = 50,000 .. 100,000 points, distance matrix made on the
device (BSA_IN_DEVICE), single and average linkage.  Reports the HBM roofline fraction of the whole run
(n^3/6 * 4 algorithmic bytes: the closest_elements scans) and checks what can be checked without an
oracle run: single-link merge distances never decrease, every merge names two live matrix indices.
usage: python tools/bench_hclust_large.py n [linkage ...]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, clustering as cl, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
links = sys.argv[2:] or ["single", "average"]
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except OSError:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)
MODE = "full scan per merge" if os.environ.get("BSA_HC_NN") == "0" else "row cache (nearest neighbour per row)"
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.empty((n, n), dtype=torch.float32, device="cuda")
for r0 in range(0, n, 4096):          # distances like 100 - identity: two decimals, plenty of ties
    r1 = min(n, r0 + 4096)
    d[r0:r1] = torch.round(torch.rand((r1 - r0, n), generator=g, device="cuda") * 8000.0 + 1000.0) / 100.0
torch.cuda.synchronize()
with Context(0) as ctx:
    for name in links:
        link = {"single": cl.single_link, "average": cl.average_link, "complete": cl.complete_link}[name]
        k = n - 1
        mi, mj, md = np.zeros(k, np.uint32), np.zeros(k, np.uint32), np.zeros(k, np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = time.perf_counter()
        ctx._ck(ctx._L.bsa_hclust(ctx._h, n, C.c_void_p(d.data_ptr()), link.code, _lib.IN_DEVICE, p(mi), p(mj), p(md)))
        wall = time.perf_counter() - t0
        st = ctx.stats()
        alg_bytes = sum(4.0 * o * (o - 1) / 2 for o in range(2, n + 1))
        order = n - np.arange(k)                              # live matrix size at each step
        ok = bool((mi < mj).all() and (mj.astype(np.int64) < order).all())
        if name == "single":
            ok = ok and bool((np.diff(md) >= 0).all())
        print(json.dumps({"what": "hclust " + name, "n": n, "matrix_gb": 4.0 * n * n / 1e9, "kernel_ms": st["kernel_ms"],
                          "wall_ms": wall * 1e3, "merges_per_s": k / (st["kernel_ms"] / 1e3), "launches": st["launches"],
                          "properties_ok": ok,
                          "mode": MODE,
                          # full scan: n^3/6 * 4 algorithmic bytes against the HBM roofline.  Row cache: the same figure is
                          # only the scan traffic the cache AVOIDS per second (it can exceed the HBM peak), not a roofline
                          ("roofline" if MODE == "full scan per merge" else "scan_equivalent"):
                              {"bound": "hbm", "achieved": alg_bytes / (st["kernel_ms"] / 1e3) / 1e9, "peak": hbm,
                               "unit": "GB/s", "frac": alg_bytes / (st["kernel_ms"] / 1e3) / 1e9 / hbm,
                               "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}), flush=True)
