#!/usr/bin/env python3
"""Scale sanity check: all-vs-all on N synthetic proteins (default 30,000: 4.5e8 pairs), results
kept on the GPU, then a random sample of pairs is compared with the CPU oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, synth  # noqa: E402
from bioshell_b200.scoring import ncbi_text  # noqa: E402
from oracle import c_oracle  # noqa: E402
import torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
res, off = synth.config("cfg3", n=n)
lens = np.diff(off.astype(np.int64))
with Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    counts = np.arange(n, dtype=np.uint32)
    npairs = n * (n - 1) // 2
    ds = torch.empty(npairs, dtype=torch.int32, device="cuda")
    dn = torch.empty(npairs, dtype=torch.int32, device="cuda")
    t0 = time.perf_counter()
    ctx.align_all_pairs(0, 0, counts, scores=ds.data_ptr(), n_identical=dn.data_ptr(), device_out=True)
    st = ctx.stats()
    print("N=%d pairs=%d cells=%.3e: %.1f GCUPS, kernel %.1f s, wall %.1f s, items %d, fallback pairs %d" % (
        n, npairs, st["cells"], st["cells"] / 1e6 / st["kernel_ms"], st["kernel_ms"] / 1e3, time.perf_counter() - t0,
        st["items"], st["fallback_pairs"]), flush=True)
    rng = np.random.default_rng(5)
    t = rng.integers(1, n, 4000)
    q = (rng.random(4000) * t).astype(np.int64)
    # include the very last pairs and the longest sequences
    order = np.argsort(lens)[-8:]
    for a in order:
        for b in order:
            if a < b:
                q = np.append(q, a); t = np.append(t, b)
    q = np.append(q, [n - 2, 0]); t = np.append(t, [n - 1, n - 1])
    k = torch.from_numpy(t * (t - 1) // 2 + q).cuda()
    gs, gn = ds[k].cpu().numpy(), dn[k].cpu().numpy()
sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
S = c_oracle.SeqSet.from_packed(res, off)
ref = c_oracle.align_pair_list(S, S, sc, ai, -10, -1, q, t, int(lens.max()), n_threads=os.cpu_count())
bad = int(np.count_nonzero((gs != ref["score"]) | (gn.astype(np.uint32) != ref["n_identical"])))
print("sampled %d pairs vs oracle: %d mismatches" % (len(q), bad))
sys.exit(1 if bad else 0)
