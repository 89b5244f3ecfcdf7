"""Frozen oracle vectors (tests/golden/oracle_vectors.json, tools/gen_golden.py): the CPU oracle
must keep reproducing them (no silent drift), and the GPU path must match them through the C ABI."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import clustering as cl
from bioshell_b200.scoring import ncbi_text
from oracle import c_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")) as fh:
        return json.load(fh)


def _pairs(n):
    return [(q, t) for t in range(n) for q in range(t)]


def test_oracle_reproduces_frozen_vectors(gold):
    seqs = [s.encode() for s in gold["sequences"]]
    S = c_oracle.SeqSet(seqs)
    for g, lg in zip(gold["global"], gold["local"]):
        sc, ai = c_oracle.parse_ncbi(ncbi_text(g["matrix"]))
        r = c_oracle.align_all_pairs(S, S, sc, ai, g["gap_open"], g["gap_extend"], True)
        assert r["score"].tolist() == g["scores"] and r["n_identical"].tolist() == g["n_identical"]
        q, t = r["q"][7], r["t"][7]
        assert c_oracle.local_align(seqs[q], seqs[t], sc, ai, g["gap_open"], g["gap_extend"]) == lg["first_200"][7]
    m = np.array(gold["hclust_matrix_lower"], np.float32)
    m = m + m.T
    for h in gold["hclust"]:
        r = c_oracle.hclust(m, h["rule"])
        assert r["mat_i"].tolist() == h["mat_i"] and r["dist"].view(np.uint32).tolist() == h["dist_bits"]


@pytest.mark.gpu
def test_gpu_matches_frozen_vectors(gold, ctx):
    seqs = [s.encode() for s in gold["sequences"]]
    res, off = bs.pack(seqs)
    pairs = _pairs(len(seqs))
    for g, lg in zip(gold["global"], gold["local"]):
        ctx.set_scoring(g["matrix"], g["gap_open"], g["gap_extend"])
        ctx.load_sequences(0, res, off)
        scores, nid = ctx.all_vs_all(0)
        assert scores.tolist() == g["scores"] and nid.tolist() == g["n_identical"]
        q = [p[0] for p in pairs[:120]]
        t = [p[1] for p in pairs[:120]]
        s2, n2, paths = ctx.align_pairs_paths(0, 0, q, t)
        assert [p.decode() for p in paths] == g["paths_first_120"]
        assert s2.tolist() == g["scores"][:120] and n2.tolist() == g["n_identical"][:120]
        q = [p[0] for p in pairs[:200]]
        t = [p[1] for p in pairs[:200]]
        lo = ctx.local_align_pairs(0, 0, q, t)
        for k, ref in enumerate(lg["first_200"]):
            got = dict(score=int(lo["score"][k]), path=lo["paths"][k].decode(), end_q=int(lo["end_q"][k]),
                       end_t=int(lo["end_t"][k]), start_q=int(lo["start_q"][k]), start_t=int(lo["start_t"][k]))
            assert got == ref
    m = np.array(gold["hclust_matrix_lower"], np.float32)
    m = m + m.T
    links = dict(single=cl.single_link, complete=cl.complete_link, average=cl.average_link, median=cl.median_link,
                 centroid=cl.centroid_link, ward=cl.wards_method)
    for h in gold["hclust"]:
        mi, mj, md = cl.hclust_merge_log(40, m, links[h["rule"]], ctx)
        assert mi.tolist() == h["mat_i"] and mj.tolist() == h["mat_j"]
        assert md.view(np.uint32).tolist() == h["dist_bits"]


@pytest.mark.gpu
def test_c_entry_points_all_vs_all_and_one_vs_many(gold, ctx):
    """bsa_all_vs_all / bsa_one_vs_many called directly (the Python mirror goes through
    bsa_align_all_pairs)."""
    seqs = [s.encode() for s in gold["sequences"]]
    res, off = bs.pack(seqs)
    g = gold["global"][0]
    ctx.set_scoring(g["matrix"], g["gap_open"], g["gap_extend"])
    ctx.load_sequences(0, res, off)
    ctx.load_sequences(1, res, off)
    n = len(seqs)
    sc = np.zeros(n * (n - 1) // 2, np.int32)
    ni = np.zeros(n * (n - 1) // 2, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx._ck(ctx._L.bsa_all_vs_all(ctx._h, 0, 3, p(sc), p(ni)))
    assert sc.tolist() == g["scores"] and ni.tolist() == g["n_identical"]
    full = np.zeros(n * n, np.int32)
    ctx._ck(ctx._L.bsa_one_vs_many(ctx._h, 0, 1, 1, p(full), None))
    full = full.reshape(n, n)                      # [t][q]
    assert [int(full[t, q]) for t in range(n) for q in range(t)] == g["scores"]
    assert np.array_equal(full, full.T)
