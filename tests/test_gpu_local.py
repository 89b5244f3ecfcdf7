"""GPU local alignment (bsa_local_align_pairs) vs the oracle: the reference's own test cases and
random / homologous pairs, bit-exact in score, end point, start point and path."""
import random

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import synth
from oracle import c_oracle

pytestmark = pytest.mark.gpu
AA = b"ARNDCQEGHILKMFPSTWYV"
DIRTY = b"ARNDCQEGHILKMFPSTWYVXBZ-_ax"


def test_reference_local_kats_on_gpu(kats, ctx):
    """bioshell-seq/tests/test_aligners.rs:73-149, both orientations, through the reference-named wrapper."""
    k = kats["local_alignment"]
    al = bs.LocalAlignment(40, ctx)
    for c in k["cases"]:
        score = al.align(c["query"], c["template"], k["matrix"], k["gap_open"], k["gap_extend"])
        path, qs, ts = al.backtrace()
        aq, at = bs.aligned_strings(path, c["query"][qs:], c["template"][ts:], "-")
        assert score == c["score"] == al.recent_score()
        assert (qs, ts) == (c["query_start"], c["template_start"])
        assert (aq, at) == (c["aligned_query"], c["aligned_template"])
        assert path == c["alignment"]
        score = al.align(c["template"], c["query"], k["matrix"], k["gap_open"], k["gap_extend"])
        path, qs, ts = al.backtrace()
        aq, at = bs.aligned_strings(path, c["template"][qs:], c["query"][ts:], "-")
        assert score == c["score"]
        assert (aq, at) == (c["aligned_template"], c["aligned_query"])
    # local.rs:75-81 doc-test
    assert al.align("TSAILDSLGAEEIRAYLP", "MQRPILDSLGNPTAEEVKAFHW", "BLOSUM62", -10, -2) == 40


@pytest.mark.parametrize("go,ge", [(-10, -2), (-10, -1), (-4, -4), (-2, -1), (-3, 0), (-12, -3)])
def test_local_random_pairs_vs_oracle(ctx, oracle_matrices, go, ge):
    rng = random.Random(go * 31 + ge)
    seqs = []
    for i in range(60):
        n = rng.choice([1, 2, 5, 17, 31, 32, 33, 64, 65, 100, 257, 300, 700, 1100])
        s = bytes(rng.choice(DIRTY if i % 4 == 0 else AA) for _ in range(n))
        seqs.append(s)
        if i % 2:
            m = bytearray(seqs[-2])
            for x in range(0, len(m), 4):
                m[x] = rng.choice(AA)
            del m[len(m) // 2: len(m) // 2 + rng.randint(0, 3)]
            seqs[-1] = (bytes(rng.choice(AA) for _ in range(rng.randint(0, 20))) + bytes(m)) or b"A"
    seqs += [b"", b"WWWW", b"W"]
    res, off = bs.pack(seqs)
    ctx.set_scoring("BLOSUM62", go, ge)
    ctx.load_sequences(0, res, off)
    qi = np.array([rng.randrange(len(seqs)) for _ in range(500)])
    ti = np.array([rng.randrange(len(seqs)) for _ in range(500)])
    out = ctx.local_align_pairs(0, 0, qi, ti)
    M = oracle_matrices["BLOSUM62"]
    for k in range(len(qi)):
        ref = c_oracle.local_align(seqs[qi[k]], seqs[ti[k]], M[0], M[1], go, ge)
        got = dict(score=int(out["score"][k]), path=out["paths"][k].decode(), end_q=int(out["end_q"][k]),
                   end_t=int(out["end_t"][k]), start_q=int(out["start_q"][k]), start_t=int(out["start_t"][k]))
        assert got == ref, (k, seqs[qi[k]], seqs[ti[k]])


def test_local_long_pairs_multipass(ctx, oracle_matrices):
    res, off = synth.generate(6, seed=12, dist=0, lo=1200, hi=3000, homolog_fraction=0.8)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    qi, ti = np.array([0, 1, 2, 3, 5, 4]), np.array([1, 0, 3, 5, 2, 4])
    out = ctx.local_align_pairs(0, 0, qi, ti)
    raw = res.tobytes()
    M = oracle_matrices["BLOSUM62"]
    for k in range(6):
        q = raw[int(off[qi[k]]):int(off[qi[k] + 1])]
        t = raw[int(off[ti[k]]):int(off[ti[k] + 1])]
        ref = c_oracle.local_align(q, t, M[0], M[1], -10, -1)
        assert ref["score"] == out["score"][k] and ref["path"] == out["paths"][k].decode()
        assert (ref["end_q"], ref["end_t"], ref["start_q"], ref["start_t"]) == \
            (out["end_q"][k], out["end_t"][k], out["start_q"][k], out["start_t"][k])
