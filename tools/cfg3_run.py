#!/usr/bin/env python3
"""BASELINE configs[2]: 100,000 synthetic proteins all-vs-all (4.99995e9 pairs, ~4.1e14 cells),
template-range shards over the ranks of one torchrun job (one process per GPU, no collective on
the data path).  Each rank keeps its slice of the results in HBM; rank-local samples are checked
against the CPU oracle.  Prints one JSON line on rank 0.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg3_run.py [N]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioshell_b200 import Context, synth  # noqa: E402
from bioshell_b200.scoring import ncbi_text  # noqa: E402
from oracle import c_oracle  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
torch.cuda.set_device(local)
if world > 1:
    if os.environ.get("NCCL_DEBUG"):
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res, off = synth.config("cfg3", n=n)
lens = np.diff(off.astype(np.int64))
counts = np.arange(n, dtype=np.uint32)
ctx = Context(local)
ctx.set_scoring("BLOSUM62", -10, -1)
ctx.load_sequences(0, res, off)
b = ctx.plan_shards(0, 0, counts, world)
t0, t1 = int(b[rank]), int(b[rank + 1])
first = np.concatenate([[0], np.cumsum(counts.astype(np.int64))])
n_res = int(first[t1] - first[t0])
ds = torch.empty(n_res, dtype=torch.int32, device="cuda")
dn = torch.empty(n_res, dtype=torch.int32, device="cuda")
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
w0 = time.perf_counter()
ctx.align_all_pairs(0, 0, counts, t0, t1, scores=ds.data_ptr(), n_identical=dn.data_ptr(), device_out=True)
st = ctx.stats()
torch.cuda.synchronize()
wall = time.perf_counter() - w0
# rank-local sample against the oracle
rng = np.random.default_rng(100 + rank)
t = rng.integers(max(t0, 1), t1, 1500)
q = (rng.random(1500) * t).astype(np.int64)
k = torch.from_numpy(first[t] - first[t0] + q).cuda()
gs, gn = ds[k].cpu().numpy(), dn[k].cpu().numpy().astype(np.uint32)
sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
S = c_oracle.SeqSet.from_packed(res, off)
ref = c_oracle.align_pair_list(S, S, sc, ai, -10, -1, q, t, int(lens.max()), n_threads=max(1, (os.cpu_count() or 8) // world))
bad = int(np.count_nonzero((gs != ref["score"]) | (gn != ref["n_identical"])))
v = torch.tensor([st["kernel_ms"], wall * 1e3], dtype=torch.float64, device="cuda")
u = torch.tensor([float(st["cells"]), float(st["pairs"]), float(bad), float(st["fallback_pairs"])], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
if rank == 0:
    ms, wms = v.tolist()
    cells, pairs, bad, fb = u.tolist()
    print(json.dumps({"workload": "cfg3: %d synthetic proteins all-vs-all, scores+identity" % n, "n_gpus": world,
                      "pairs": int(pairs), "cells": cells, "kernel_s_max_over_ranks": ms / 1e3, "wall_s": wms / 1e3,
                      "GCUPS": cells / 1e6 / ms, "pairs_per_s": pairs / (ms / 1e3),
                      "oracle_sample_pairs": 1500 * world, "oracle_mismatches": int(bad), "fallback_pairs": int(fb)}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
