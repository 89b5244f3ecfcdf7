"""Size-independent properties at (or near) BASELINE.json's sizes, plus invariance checks:
results do not depend on how the work is tiled or sharded."""
import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import sharding, synth
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def test_cfg1_full_parity(ctx, oracle_matrices):
    """BASELINE configs[0]: 1,000 proteins (len 50-500), BLOSUM62, -10/-1: all 499,500 pairs."""
    res, off = synth.config("cfg1")
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    scores, nid = ctx.all_vs_all(0)
    S = c_oracle.SeqSet.from_packed(res, off)
    M = oracle_matrices["BLOSUM62"]
    ref = c_oracle.align_all_pairs(S, S, M[0], M[1], -10, -1, True, n_threads=16)
    assert ref["n_pairs"] == 499500
    assert np.array_equal(scores, ref["score"]) and np.array_equal(nid, ref["n_identical"])
    st = ctx.stats()
    assert st["cells"] == int(ref["cells"]) and st["fallback_pairs"] == 0


def test_cfg2_sample_and_properties(ctx, oracle_matrices):
    """BASELINE configs[1] (10,000 UniRef50-like proteins) at full size: a random sample of
    pairs against the oracle, plus properties that need no oracle."""
    res, off = synth.config("cfg2")
    n = len(off) - 1
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    scores, nid = ctx.all_vs_all(0)
    assert len(scores) == n * (n - 1) // 2
    lens = np.diff(off.astype(np.int64))
    rng = np.random.default_rng(12)
    t = rng.integers(1, n, 3000)
    q = (rng.random(3000) * t).astype(np.int64)
    S = c_oracle.SeqSet.from_packed(res, off)
    M = oracle_matrices["BLOSUM62"]
    ref = c_oracle.align_pair_list(S, S, M[0], M[1], -10, -1, q, t, int(lens.max()), n_threads=16)
    k = t * (t - 1) // 2 + q
    assert np.array_equal(scores[k], ref["score"]) and np.array_equal(nid[k], ref["n_identical"])
    # properties over ALL 5e7 pairs
    tt = np.repeat(np.arange(n), np.arange(n))
    qq = np.arange(len(tt)) - (tt * (tt - 1) // 2)
    mn = np.minimum(lens[qq], lens[tt])
    assert np.all(nid <= mn)
    assert np.all(scores <= 11 * mn)                                        # max BLOSUM62 entry
    assert np.all(scores >= 2 * -10 + (lens[qq] + lens[tt]) * -1 - 4 * mn)  # all-gap / all-mismatch floor
    # a checksum of checksums, stable across tilings (recomputed below with another tiling)
    chk = (int(scores.astype(np.int64).sum()), int(nid.astype(np.int64).sum()))
    b = ctx.plan_shards(0, 0, np.arange(n, dtype=np.uint32), 3)
    parts = [ctx.align_all_pairs(0, 0, np.arange(n, dtype=np.uint32), int(b[i]), int(b[i + 1])) for i in range(3)]
    s2 = np.concatenate([p[0] for p in parts])
    n2 = np.concatenate([p[1] for p in parts])
    assert (int(s2.astype(np.int64).sum()), int(n2.astype(np.int64).sum())) == chk
    assert np.array_equal(s2, scores) and np.array_equal(n2, nid)


def test_self_alignment_and_symmetry_properties(ctx):
    """A sequence against itself scores the sum of its diagonal and is 100 % identical;
    scores are symmetric under swapping the two sets (SURVEY.md 8a note 5)."""
    res, off = synth.generate(300, seed=4, dist=1)
    m = bs.SubstitutionMatrix.load("BLOSUM62")
    ctx.set_scoring(m, -10, -1)
    ctx.load_sequences(0, res, off)
    ctx.load_sequences(1, res, off)
    s, nid = ctx.one_vs_many(0, 1, want_identical=True)
    s = s.reshape(300, 300)          # [t][q]
    nid = nid.reshape(300, 300)
    lens = np.diff(off.astype(np.int64))
    idx = m.aa_indexes[res].astype(np.int64)
    diag = m.score[idx * 21 + idx].astype(np.int64)
    c = np.concatenate([[0], np.cumsum(diag)])
    self_score = c[off[1:].astype(np.int64)] - c[off[:-1].astype(np.int64)]
    assert np.array_equal(np.diag(s), self_score)
    assert np.array_equal(np.diag(nid), lens)
    assert np.array_equal(s, s.T)


def test_sharded_ranges_concatenate_to_the_whole(ctx):
    res, off = synth.generate(700, seed=21, dist=1)
    n = 700
    ctx.set_scoring("BLOSUM62", -11, -1)
    ctx.load_sequences(0, res, off)
    counts = np.arange(n, dtype=np.uint32)
    whole = ctx.align_all_pairs(0, 0, counts)
    lens = np.diff(off.astype(np.int64))
    for shards in (2, 4, 8):
        b = ctx.plan_shards(0, 0, counts, shards)
        assert np.array_equal(b.astype(np.int64), sharding.plan_shards(lens, lens, counts, shards))
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b.astype(np.int64)) >= 0)
        parts = [ctx.align_all_pairs(0, 0, counts, int(b[i]), int(b[i + 1])) for i in range(shards)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), whole[0])
        assert np.array_equal(np.concatenate([p[1] for p in parts]), whole[1])


def test_page_locked_result_buffers_take_direct_stores(ctx):
    """Results stored by the kernels straight into page-locked caller buffers (no staging, no copy) are the
    bytes the staged path delivers into pageable buffers -- 32-bit score + identity and 16-bit score only."""
    import torch
    res, off = synth.generate(600, seed=21)
    n = len(off) - 1
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    s_ref, i_ref = ctx.all_vs_all(0)                       # pageable numpy buffers: staged + copied
    k = len(s_ref)
    hs = torch.empty(k, dtype=torch.int32).pin_memory()
    hi = torch.empty(k, dtype=torch.int32).pin_memory()
    hs.fill_(-1), hi.fill_(-1)
    ctx.align_all_pairs(0, 0, np.arange(n, dtype=np.uint32), scores=hs.numpy(), n_identical=hi.numpy().view(np.uint32))
    assert ctx.stats()["d2h_bytes"] == 8 * k
    assert np.array_equal(hs.numpy(), s_ref) and np.array_equal(hi.numpy().view(np.uint32), i_ref)
    hs.fill_(-1)
    ctx.align_all_pairs(0, 0, np.arange(n, dtype=np.uint32), want_identical=False, scores=hs.numpy())
    assert np.array_equal(hs.numpy(), s_ref)


def test_two_contexts_with_different_alphabets_take_turns():
    """The dynamic shared-memory limit of a kernel belongs to the (device, function), not to a context: a context
    with a small alphabet must not lower it under a context with a large one (the library only ever raises it).
    Streaming kernels and the wavefront kernel, calls alternating between the two contexts."""
    rng = np.random.default_rng(3)
    res, off = synth.generate(120, seed=31, dist=0, lo=40, hi=700)
    wide = res.copy()
    idx = rng.integers(0, len(wide), 400)
    wide[idx] = rng.integers(130, 200, 400).astype(np.uint8)        # 70 more byte values: a much larger profile
    lres, loff = synth.pair_set(1, seed=9, lo=4300, hi=4800)
    lwide = lres.copy()
    lwide[rng.integers(0, len(lwide), 50)] = rng.integers(130, 200, 50).astype(np.uint8)
    with bs.Context(0) as a, bs.Context(0) as b:
        for c, r, lr in ((a, wide, lwide), (b, res, lres)):
            c.set_scoring("BLOSUM62", -10, -1)
            c.load_sequences(0, r, off)
            c.load_sequences(1, lr, loff)
        first = {}
        for rnd in range(2):
            for name, c in (("a", a), ("b", b)):
                s, n = c.all_vs_all(0)
                s2, n2, p2 = c.align_pairs_paths(1, 1, [0], [1])
                got = (s.tobytes(), n.tobytes(), int(s2[0]), int(n2[0]), bytes(p2[0]))
                if rnd == 0:
                    first[name] = got
                else:
                    assert got == first[name], name
