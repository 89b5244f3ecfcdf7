"""N>1 host logic on CPU: two `gloo` ranks derive the same template-range shards with no
exchange, each fills its slice (the oracle stands in for the GPU here -- tests only), and the
concatenation equals the single-rank result bit for bit.  Mirrors what bench.py does under
torchrun with one CUDA context per rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bioshell_b200 import sharding, synth
from bioshell_b200.scoring import ncbi_text
from oracle import c_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res, off = synth.generate(n, seed=42, dist=0, lo=5, hi=90)
    lens = np.diff(off.astype(np.int64))
    counts = np.arange(n, dtype=np.int64)                   # strict upper triangle
    bounds = sharding.plan_shards(lens, lens, counts, world)
    # every rank must hold identical bounds without communicating: check by all_gather
    mine = torch.from_numpy(bounds.copy())
    got = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    assert all(torch.equal(g, mine) for g in got)
    t0, t1 = int(bounds[rank]), int(bounds[rank + 1])
    sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
    S = c_oracle.SeqSet.from_packed(res, off)
    pt = np.repeat(np.arange(t0, t1), counts[t0:t1])
    first = np.concatenate([[0], np.cumsum(counts)])
    pq = np.arange(len(pt)) - (first[pt] - first[t0])
    r = c_oracle.align_pair_list(S, S, sc, ai, -10, -1, pq, pt, int(lens.max()))
    np.savez(os.path.join(out_dir, "part%d.npz" % rank), score=r["score"], nid=r["n_identical"],
             base=sharding.shard_result_offsets(counts, bounds)[rank], cells=r["cells"])
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_ranks_concatenate_to_single_rank_result(tmp_path, world):
    n = 60
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    res, off = synth.generate(n, seed=42, dist=0, lo=5, hi=90)
    sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
    S = c_oracle.SeqSet.from_packed(res, off)
    whole = c_oracle.align_all_pairs(S, S, sc, ai, -10, -1, True)
    score = np.zeros(whole["n_pairs"], np.int32)
    nid = np.zeros(whole["n_pairs"], np.uint32)
    filled = 0
    cells = 0.0
    for r in range(world):
        p = np.load(os.path.join(str(tmp_path), "part%d.npz" % r))
        b = int(p["base"])
        score[b:b + len(p["score"])] = p["score"]
        nid[b:b + len(p["nid"])] = p["nid"]
        filled += len(p["score"])
        cells += float(p["cells"])
    assert filled == whole["n_pairs"]
    assert np.array_equal(score, whole["score"]) and np.array_equal(nid, whole["n_identical"])
    assert cells == whole["cells"]


def test_plan_shards_balances_cells():
    res, off = synth.generate(2000, seed=3, dist=1)
    lens = np.diff(off.astype(np.int64))
    counts = np.arange(2000)
    for world in (2, 4, 8):
        b = sharding.plan_shards(lens, lens, counts, world)
        assert b[0] == 0 and b[-1] == 2000 and np.all(np.diff(b) >= 0)
        qoff = np.concatenate([[0], np.cumsum(lens)])
        work = lens * qoff[counts]
        per = np.array([work[b[i]:b[i + 1]].sum() for i in range(world)], np.float64)
        assert per.max() / per.mean() < 1.05
    # rectangle (one-vs-many): every template has the same query count
    b = sharding.plan_shards(lens[:100], lens, None, 4)
    assert b[-1] == 2000
