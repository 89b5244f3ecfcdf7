#!/usr/bin/env python3
"""Throughput of the GPU hierarchical clustering against its HBM roofline,
with the oracle timed beside it on a smaller n.  The dominant kernel (hclust_argmin_kernel) reads
order^2/2 f32 per merge: n^3/6 * 4 algorithmic bytes per clustering."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, clustering as cl  # noqa: E402
from oracle import c_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rng = np.random.default_rng(1)
a = np.round(rng.random((n, n), dtype=np.float32) * 80 + 10, 2)
a = np.tril(a, -1)
m = (a + a.T).astype(np.float32)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except OSError:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)
MODE = "full scan per merge" if os.environ.get("BSA_HC_NN") == "0" else "row cache (nearest neighbour per row)"
with Context(0) as ctx:
    for link in (cl.single_link, cl.average_link):
        for rep in range(2):
            t0 = time.perf_counter()
            mi, mj, md = cl.hclust_merge_log(n, m, link, ctx)
            wall = time.perf_counter() - t0
            st = ctx.stats()
        alg_bytes = sum(4.0 * o * (o - 1) / 2 for o in range(2, n + 1))
        print(json.dumps({"what": "hclust " + link.name, "n": n, "kernel_ms": st["kernel_ms"], "wall_ms": wall * 1e3,
                          "merges_per_s": (n - 1) / (st["kernel_ms"] / 1e3), "launches": st["launches"],
                          "mode": MODE,
                          # full scan: n^3/6 * 4 algorithmic bytes against the HBM roofline.  Row cache: the same figure is
                          # only the scan traffic the cache AVOIDS per second (it can exceed the HBM peak), not a roofline
                          ("roofline" if MODE == "full scan per merge" else "scan_equivalent"):
                              {"bound": "hbm", "achieved": alg_bytes / (st["kernel_ms"] / 1e3) / 1e9, "peak": hbm,
                               "unit": "GB/s", "frac": alg_bytes / (st["kernel_ms"] / 1e3) / 1e9 / hbm,
                               "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}), flush=True)
nc = min(n, 2000)
t0 = time.perf_counter()
c_oracle.hclust(m[:nc, :nc].copy(), "single")
dt = time.perf_counter() - t0
print(json.dumps({"what": "oracle (1 thread, as the reference)", "n": nc, "seconds": dt,
                  "extrapolated_seconds_for_n": dt * (n / nc) ** 3}), flush=True)
