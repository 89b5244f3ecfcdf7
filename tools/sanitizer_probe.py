#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck) on a GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bioshell_b200 as bs  # noqa: E402
from bioshell_b200 import clustering as cl, synth  # noqa: E402

with bs.Context(0) as ctx:
    ctx.set_scoring("BLOSUM62", -10, -1)
    res, off = synth.generate(40, seed=3, dist=0, lo=1, hi=700)       # K = 1..20 and a multi-pass template
    ctx.load_sequences(0, res, off)
    s, n = ctx.all_vs_all(0)
    print("all_vs_all", len(s), int(s.astype(np.int64).sum()), int(n.astype(np.int64).sum()))
    s16, _ = ctx.one_vs_many(0, 0, want_identical=False)               # 16-bit frame kernels over the aligned stream
    print("one_vs_many score only", len(s16), int(s16.astype(np.int64).sum()))
    res2, off2 = synth.generate(3, seed=4, dist=0, lo=4200, hi=4600)   # wavefront kernel
    ctx.load_sequences(1, res2, off2)
    s2, n2, p2 = ctx.align_pairs_paths(1, 1, [0, 1], [1, 2])
    s3, n3, p3 = ctx.align_pairs_paths(0, 0, np.arange(0, 20), np.arange(20, 40))
    print("paths", s2.tolist(), [len(x) for x in p2], int(s3.sum()))
    rng = np.random.default_rng(0)
    a = np.tril(rng.integers(1, 9, (70, 70)).astype(np.float32), -1)
    mi, mj, md = cl.hclust_merge_log(70, a + a.T, cl.average_link, ctx)
    print("hclust", int(mi.sum()), int(mj.sum()), float(md.sum()))
