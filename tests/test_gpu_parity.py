"""(iii) GPU-vs-oracle parity, bit-exact, through the C ABI.  Every test compares the CUDA
path with the CPU oracle on the same inputs; full-size configurations are covered by
size-independent properties in test_gpu_properties.py."""
import random

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import synth
from oracle import c_oracle

pytestmark = pytest.mark.gpu

AA = b"ARNDCQEGHILKMFPSTWYV"
DIRTY = b"ARNDCQEGHILKMFPSTWYVXBZJUO*-_arndx"


def oracle_triangle(res, off, M, go, ge, threads=8):
    S = c_oracle.SeqSet.from_packed(res, off)
    return c_oracle.align_all_pairs(S, S, M[0], M[1], go, ge, True, n_threads=threads)


def test_reference_kats_on_gpu(kats, ctx):
    """bioshell-seq/tests/test_aligners.rs:13-58 through the GPU kernels, both orientations."""
    g = kats["global_aligner"]
    ctx.set_scoring(g["matrix"], g["gap_open"], g["gap_extend"])
    for c in g["cases"]:
        for q, t, eq, et in ((c["query"], c["template"], c["aligned_query"], c["aligned_template"]),
                             (c["template"], c["query"], c["aligned_template"], c["aligned_query"])):
            res, off = bs.pack([q, t])
            ctx.load_sequences(0, res, off)
            s, nid, paths = ctx.align_pairs_paths(0, 0, [0], [1])
            assert int(s[0]) == c["score"]
            aq, at = bs.aligned_strings(paths[0].decode(), q, t, "-")
            assert (aq, at) == (eq, et)
            if q == c["query"]:
                assert paths[0].decode() == c["alignment"]
            # the streaming (score + identity) kernel gives the same score
            s2, n2 = ctx.align_all_pairs(0, 0, np.array([0, 1], np.uint32))
            assert int(s2[0]) == c["score"] and int(n2[0]) == bs.count_identical(aq, at)
    # the reference-named single-pair wrapper
    al = bs.GlobalAligner(10, ctx)
    assert al.align("MAVRLLKTHL", "MKNITCYL", "BLOSUM62", -10, -2) == -2
    assert al.backtrace() == "*||*******" and al.recent_score() == -2


@pytest.mark.parametrize("matrix,go,ge", [("BLOSUM62", -10, -1), ("BLOSUM62", -10, -2), ("PAM30", -11, -1),
                                          ("BLOSUM80", -5, -5), ("PAM250", -1, -1), ("BLOSUM45", -12, -3),
                                          ("BLOSUM62", -3, 0)])
def test_all_vs_all_small_sets(ctx, oracle_matrices, matrix, go, ge):
    res, off = synth.generate(120, seed=hash((matrix, go, ge)) % 1000 + 1, dist=0, lo=1, hi=150)
    ctx.set_scoring(matrix, go, ge)
    ctx.load_sequences(0, res, off)
    scores, nid = ctx.all_vs_all(0)
    ref = oracle_triangle(res, off, oracle_matrices[matrix], go, ge)
    assert len(scores) == 120 * 119 // 2 == ref["n_pairs"]
    assert np.array_equal(scores, ref["score"])
    assert np.array_equal(nid, ref["n_identical"])
    assert ctx.stats()["fallback_pairs"] == 0


def test_every_columns_per_lane_variant(ctx, oracle_matrices):
    """Template lengths 1..1100 exercise every K = 1..32 kernel and the two-pass path."""
    rng = random.Random(3)
    lens = list(range(1, 40)) + [rng.randint(1, 1100) for _ in range(60)] + [32 * k for k in range(1, 33)] + \
           [32 * k + 1 for k in range(1, 33)] + [1025, 1100, 2049]
    seqs = [bytes(rng.choice(AA) for _ in range(n)) for n in lens]
    # make some of them related so gaps/ties occur
    for i in range(5, len(seqs), 7):
        base = bytearray(seqs[i - 1])
        for k in range(0, len(base), 9):
            base[k] = rng.choice(AA)
        seqs[i] = bytes(base[: max(1, len(base) - rng.randint(0, 5))])
    res, off = bs.pack(seqs)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    scores, nid = ctx.all_vs_all(0)
    ref = oracle_triangle(res, off, oracle_matrices["BLOSUM62"], -10, -1)
    bad = np.nonzero((scores != ref["score"]) | (nid != ref["n_identical"]))[0]
    assert len(bad) == 0, [(int(ref["q"][k]), int(ref["t"][k]), len(seqs[ref["q"][k]]), len(seqs[ref["t"][k]]),
                            int(scores[k]), int(ref["score"][k]), int(nid[k]), int(ref["n_identical"][k]))
                           for k in bad[:10]]


def test_dirty_bytes_and_identity_on_raw_bytes(ctx, oracle_matrices):
    """Unknown bytes score as 'A' but count as identical only when the RAW bytes match;
    '-' and '_' never count (similarity_score.rs:125-134, msa.rs:264)."""
    rng = random.Random(9)
    seqs = [bytes(rng.choice(DIRTY) for _ in range(rng.randint(1, 90))) for _ in range(80)]
    seqs += [b"BBBB", b"AAAA", b"bbbb", b"----", b"_-_-", b"XXXX", b"B", b"-"]
    res, off = bs.pack(seqs)
    for go, ge in ((-10, -1), (-2, -1)):
        ctx.set_scoring("BLOSUM62", go, ge)
        ctx.load_sequences(0, res, off)
        scores, nid = ctx.all_vs_all(0)
        ref = oracle_triangle(res, off, oracle_matrices["BLOSUM62"], go, ge)
        assert np.array_equal(scores, ref["score"]) and np.array_equal(nid, ref["n_identical"])
    # identity percentages incl. the ungapped-length denominator (raw '-' shortens it)
    r = bs.PairResults(scores, nid, np.arange(len(seqs)), bs.ungapped_lengths(res, off), bs.ungapped_lengths(res, off))
    got, exp = r.percent_identity(), ref["identity"]
    both_nan = np.isnan(got) & np.isnan(exp)
    assert np.array_equal(got[~both_nan], exp[~both_nan]) and np.array_equal(np.isnan(got), np.isnan(exp))


def test_rectangle_with_empty_and_length_one_sequences(ctx, oracle_matrices):
    rng = random.Random(2)
    Q = [b"", b"A", b"W", b"ARNDARND", b"", bytes(rng.choice(AA) for _ in range(200))]
    T = [b"K", b"", bytes(rng.choice(AA) for _ in range(70)), b"ARND", b"A"]
    ctx.set_scoring("BLOSUM62", -10, -2)
    qr, qo = bs.pack(Q)
    tr, to = bs.pack(T)
    ctx.load_sequences(0, qr, qo)
    ctx.load_sequences(1, tr, to)
    scores, nid = ctx.one_vs_many(0, 1, want_identical=True)
    M = oracle_matrices["BLOSUM62"]
    ref = c_oracle.align_all_pairs(c_oracle.SeqSet(Q), c_oracle.SeqSet(T), M[0], M[1], -10, -2, False)
    assert ref["n_pairs"] == 30
    assert np.array_equal(scores, ref["score"]) and np.array_equal(nid, ref["n_identical"])


def test_paths_strings_and_identity_vs_oracle(ctx, oracle_matrices):
    """K2: direction store + device traceback == GlobalAligner::backtrace, glyph for glyph."""
    rng = random.Random(17)
    seqs = []
    for i in range(70):
        n = rng.choice([1, 2, 3, 5, 17, 31, 32, 33, 64, 65, 100, 257, 300])
        s = bytes(rng.choice(DIRTY if i % 3 == 0 else AA) for _ in range(n))
        seqs.append(s)
        if i % 2:
            m = bytearray(seqs[-2])
            for k in range(0, len(m), 4):
                m[k] = rng.choice(AA)
            del m[len(m) // 2: len(m) // 2 + rng.randint(0, 3)]
            seqs[-1] = bytes(m) or b"A"
    res, off = bs.pack(seqs)
    M = oracle_matrices["BLOSUM62"]
    for go, ge in ((-10, -1), (-10, -2), (-4, -4), (-2, 0)):
        ctx.set_scoring("BLOSUM62", go, ge)
        ctx.load_sequences(0, res, off)
        qi = np.array([rng.randrange(len(seqs)) for _ in range(400)])
        ti = np.array([rng.randrange(len(seqs)) for _ in range(400)])
        s, nid, paths = ctx.align_pairs_paths(0, 0, qi, ti)
        for k in range(len(qi)):
            q, t = seqs[qi[k]], seqs[ti[k]]
            if ge == 0 and max(len(q), len(t)) < 2:
                continue
            one = c_oracle.align_pair(q, t, M[0], M[1], go, ge, lmax=300)
            assert one["score"] == s[k] and one["path"] == paths[k].decode() and one["n_identical"] == nid[k], \
                (k, q, t, go, ge)
            aq, at = bs.aligned_symbols(paths[k], q, t)
            assert aq == one["aligned_q"] and at == one["aligned_t"]


def test_long_templates_multipass_and_traceback(ctx, oracle_matrices):
    """Templates beyond 1024 columns take the multi-pass kernels; 2.5k-5k residues."""
    res, off = synth.generate(10, seed=77, dist=0, lo=1500, hi=5000)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    scores, nid = ctx.all_vs_all(0)
    ref = oracle_triangle(res, off, oracle_matrices["BLOSUM62"], -10, -1)
    assert np.array_equal(scores, ref["score"]) and np.array_equal(nid, ref["n_identical"])
    qi, ti = np.array([0, 2, 9, 4]), np.array([5, 3, 1, 4])
    s, n2, paths = ctx.align_pairs_paths(0, 0, qi, ti)
    raw = res.tobytes()
    M = oracle_matrices["BLOSUM62"]
    for k in range(4):
        q = raw[int(off[qi[k]]):int(off[qi[k] + 1])]
        t = raw[int(off[ti[k]]):int(off[ti[k] + 1])]
        one = c_oracle.align_pair(q, t, M[0], M[1], -10, -1)
        assert one["score"] == s[k] and one["n_identical"] == n2[k] and one["path"] == paths[k].decode()


def test_titin_scale_pair_wavefront(ctx, oracle_matrices):
    """BASELINE configs[4]: a 21k x 26k pair (and its transpose) through the K3 wavefront with
    full traceback; the oracle needs 3 x 26001^2 bytes, so only two pairs."""
    res, off = synth.generate(3, seed=1006, dist=0, lo=21000, hi=26000)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    qi, ti = np.array([0, 2, 1]), np.array([1, 0, 1])
    s, nid, paths = ctx.align_pairs_paths(0, 0, qi, ti)
    raw = res.tobytes()
    M = oracle_matrices["BLOSUM62"]
    for k in range(3):
        q = raw[int(off[qi[k]]):int(off[qi[k] + 1])]
        t = raw[int(off[ti[k]]):int(off[ti[k] + 1])]
        one = c_oracle.align_pair(q, t, M[0], M[1], -10, -1)
        assert one["score"] == s[k] and one["n_identical"] == nid[k]
        assert one["path"] == paths[k].decode()
    assert nid[2] == len(raw[int(off[1]):int(off[2])])      # self alignment: 100 % identical


def test_range_fallback_uses_direction_path(ctx, oracle_matrices):
    """Sequences whose score range does not fit the packed lanes go through the
    direction-store kernels and still match (7k residues with BLOSUM62 exceeds 2^16)."""
    res, off = synth.generate(4, seed=5, dist=0, lo=6500, hi=7200)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    scores, nid = ctx.all_vs_all(0)
    assert ctx.stats()["fallback_pairs"] == 6
    ref = oracle_triangle(res, off, oracle_matrices["BLOSUM62"], -10, -1)
    assert np.array_equal(scores, ref["score"]) and np.array_equal(nid, ref["n_identical"])


def test_protocol_replay_matches_reference_order(ctx, oracle_matrices):
    """align_all_pairs with a reporter: t-major order, aligned Sequences with inherited
    descriptions, and the SequenceIdentityMatrix of bin/cluster_sequences.rs."""
    res, off = synth.generate(14, seed=31, dist=0, lo=5, hi=60)
    raw = res.tobytes()
    seqs = [bs.Sequence("syn|%07d" % i, raw[int(off[i]):int(off[i + 1])]) for i in range(14)]
    seqs.append(bs.Sequence(seqs[3].description(), seqs[3].as_u8()))     # duplicate record -> break at q=3
    col = bs.CollectReporter()
    n = bs.align_all_pairs(seqs, seqs, "BLOSUM62", -10, -1, True, col, ctx=ctx)
    S = c_oracle.SeqSet([s.as_u8() for s in seqs], [s.description() for s in seqs])
    M = oracle_matrices["BLOSUM62"]
    ref = c_oracle.align_all_pairs(S, S, M[0], M[1], -10, -1, True)
    assert n == ref["n_pairs"] == 14 * 13 // 2 + 3
    for k, (aq, at) in enumerate(col.pairs):
        q, t = int(ref["q"][k]), int(ref["t"][k])
        assert aq.description() == seqs[q].description() and at.description() == seqs[t].description()
        one = c_oracle.align_pair(seqs[q].as_u8(), seqs[t].as_u8(), M[0], M[1], -10, -1, lmax=60)
        assert aq.as_u8() == one["aligned_q"] and at.as_u8() == one["aligned_t"]
    uniq = seqs[:14]
    mat = bs.SequenceIdentityMatrix(uniq)
    bs.align_all_pairs(uniq, uniq, "BLOSUM62", -10, -1, True, mat, ctx=ctx)
    batched = bs.SequenceIdentityMatrix(uniq)
    batched.fill_from(bs.align_all_vs_all(uniq, "BLOSUM62", -10, -1, ctx=ctx))
    assert np.array_equal(mat.similarity_matrix, batched.similarity_matrix)
    assert np.all(np.tril(mat.similarity_matrix) == 0)        # the reference never writes [t][q]


def test_unsupported_gap_settings_are_refused(ctx):
    for go, ge in ((-1, -2), (0, 0), (-5, 1), (2, -1)):
        with pytest.raises(bs.BsaError) as e:
            ctx.set_scoring("BLOSUM62", go, ge)
        assert e.value.rc == -2
    ctx.set_scoring("BLOSUM62", -10, -1)
    with pytest.raises(bs.BsaError):
        ctx.load_sequences(0, np.array([65, 255, 66], np.uint8), np.array([0, 3], np.uint64))


def test_alphabet_limit_and_odd_even_lengths(oracle_matrices):
    """Residue code 0 is reserved for the PAD rows of the even-aligned stream: 127 distinct byte values
    are accepted (and align like the oracle, unknown bytes scoring as 'A'), the 128th is refused.
    Lengths of both parities side by side: the aligned copy pads only the odd ones."""
    M = oracle_matrices["BLOSUM62"]
    rng = random.Random(77)
    vals = [b for b in range(1, 200) if b not in (ord("-"), ord("_"))][:127]
    seqs = [bytes(vals[i:i + n]) for i, n in ((0, 40), (40, 41), (81, 46), (3, 1), (90, 2), (60, 67))]
    seqs += [bytes(rng.choice(AA) for _ in range(n)) for n in (1, 2, 3, 4, 5, 63, 64, 65, 127, 128, 129, 300, 301)]
    res, off = bs.pack(seqs)
    with bs.Context(0) as c:
        c.set_scoring("BLOSUM62", -10, -1)
        c.load_sequences(0, res, off)
        scores, nid = c.all_vs_all(0)
        ref = oracle_triangle(res, off, M, -10, -1)
        assert np.array_equal(scores, ref["score"]) and np.array_equal(nid, ref["n_identical"])
        s16, _ = c.all_vs_all(0, want_identical=False)
        assert np.array_equal(s16, ref["score"])
        with pytest.raises(bs.BsaError) as e:
            c.load_sequences(1, np.array([vals[0], 250], np.uint8), np.array([0, 2], np.uint64))
        assert e.value.rc == -7


def test_score_only_16bit_lanes(ctx, oracle_matrices):
    """Score-only requests pair two templates per warp in s16x2 lanes (gotoh_score16_kernel);
    scores must equal the oracle's (and the 32-bit kernel's)."""
    rng = random.Random(41)
    Q = [bytes(rng.choice(AA) for _ in range(rng.randint(1, 400))) for _ in range(90)]
    # long queries too: against the long templates they stay out of the 16-bit owner swap (long x long)
    Q += [bytes(rng.choice(AA) for _ in range(n)) for n in (641, 700, 1501)]
    T = [bytes(rng.choice(AA) for _ in range(n)) for n in
         [1, 2, 31, 32, 33, 64, 100, 101, 233, 240, 250, 256, 300, 600, 640, 641, 700, 1290, 1300, 2000, 2100]]
    T += [Q[3], Q[10][:50], Q[20] + Q[21], Q[91][:650] + Q[5]]
    M = oracle_matrices["BLOSUM62"]
    qr, qo = bs.pack(Q)
    tr, to = bs.pack(T)
    # a matrix that is NOT symmetric: the owner swap (long templates streamed as rows through the queries'
    # columns) must look the substitution score up transposed
    asym = bs.SubstitutionMatrix(M[0].copy(), M[1])
    asym.score[0 * 21 + 1] += 2
    asym.score[5 * 21 + 7] -= 3
    asym.score[10 * 21 + 9] += 1
    ctx.set_scoring(asym, -10, -1)
    ctx.load_sequences(0, qr, qo)
    ctx.load_sequences(1, tr, to)
    s16, _ = ctx.one_vs_many(0, 1, want_identical=False)
    ref = c_oracle.align_all_pairs(c_oracle.SeqSet(Q), c_oracle.SeqSet(T), asym.score, M[1], -10, -1, False,
                                   n_threads=8, with_backtrace=False)
    assert np.array_equal(s16, ref["score"]), "asymmetric matrix"
    s32, _ = ctx.one_vs_many(0, 1, want_identical=True)
    assert np.array_equal(s16, s32), "asymmetric matrix, 32-bit lanes"
    for go, ge in ((-10, -1), (-11, -2), (-4, -4)):
        ctx.set_scoring("BLOSUM62", go, ge)
        ctx.load_sequences(0, qr, qo)
        ctx.load_sequences(1, tr, to)
        s16, _ = ctx.one_vs_many(0, 1, want_identical=False)
        ref = c_oracle.align_all_pairs(c_oracle.SeqSet(Q), c_oracle.SeqSet(T), M[0], M[1], go, ge, False,
                                       n_threads=8, with_backtrace=False)
        assert np.array_equal(s16, ref["score"]), (go, ge)
        s32, _ = ctx.one_vs_many(0, 1, want_identical=True)
        assert np.array_equal(s16, s32)
    # the triangle (different query counts per template) and a set with an empty query
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, qr, qo)
    s_tri, _ = ctx.all_vs_all(0, want_identical=False)
    ref = c_oracle.align_all_pairs(c_oracle.SeqSet(Q), c_oracle.SeqSet(Q), M[0], M[1], -10, -1, True, n_threads=8,
                                   with_backtrace=False)
    assert np.array_equal(s_tri, ref["score"])
    Q2 = Q[:10] + [b""] + Q[10:20]
    q2r, q2o = bs.pack(Q2)
    ctx.load_sequences(0, q2r, q2o)
    s_e, _ = ctx.one_vs_many(0, 1, want_identical=False)
    ref = c_oracle.align_all_pairs(c_oracle.SeqSet(Q2), c_oracle.SeqSet(T), M[0], M[1], -10, -1, False, n_threads=8,
                                   with_backtrace=False)
    assert np.array_equal(s_e, ref["score"])


def test_score_only_long_sequences_leave_16bit_lanes(ctx, oracle_matrices):
    """min(len) * 11 > 32k cannot be held in 16 bits: those templates take the 32-bit kernel."""
    res, off = synth.generate(6, seed=8, dist=0, lo=2500, hi=3900)
    ctx.set_scoring("BLOSUM62", -10, -1)
    ctx.load_sequences(0, res, off)
    ctx.load_sequences(1, res, off)
    s, _ = ctx.one_vs_many(0, 1, want_identical=False)
    S = c_oracle.SeqSet.from_packed(res, off)
    M = oracle_matrices["BLOSUM62"]
    ref = c_oracle.align_all_pairs(S, S, M[0], M[1], -10, -1, False, n_threads=8, with_backtrace=False)
    assert np.array_equal(s, ref["score"])
    assert ctx.stats()["items"] > 0


def _long_gap_set(seed):
    """Pairs whose optimal paths hold gaps far longer than the TAG cell's 5-bit streak field and
    the 16-row clearing period: very short against long sequences (both orientations), long
    sequences with a long insertion, runs of one residue (many equal-score ties), plus filler."""
    rng = random.Random(seed)
    core = bytes(rng.choice(AA) for _ in range(40))
    seqs = [bytes(rng.choice(AA) for _ in range(n)) for n in (3, 7, 12, 25, 33, 64, 65)]
    seqs += [bytes(rng.choice(AA) for _ in range(n)) for n in (300, 317, 500, 640, 641, 900)]
    for ins in (35, 70, 130, 260):           # same ends, a long insertion in the middle
        seqs.append(core[:20] + bytes(rng.choice(AA) for _ in range(ins)) + core[20:])
    seqs.append(core)
    seqs += [b"A" * 50, b"A" * 333, b"AW" * 100, b"W" * 45 + b"A" * 200 + b"W" * 45, b"W" * 90]
    seqs += [bytes(rng.choice(AA) for _ in range(rng.randint(1, 700))) for _ in range(20)]
    rng.shuffle(seqs)
    return seqs


@pytest.mark.parametrize("matrix,go,ge", [("BLOSUM62", -10, -1), ("BLOSUM62", -11, -1), ("BLOSUM62", -4, -4),
                                          ("PAM250", -2, 0)])
def test_long_gaps_through_every_cell_variant(ctx, oracle_matrices, monkeypatch, matrix, go, ge):
    """The TAG cell (streak field cleared at lane boundaries and every 16 rows), its two-row step,
    the paired short-template kernel and the classic cell must all give the oracle's scores and
    identities when gaps run over hundreds of cells; BSA_NO_TAG / BSA_NO_PAIR select the classic
    cell and the one-template kernels for the same inputs."""
    seqs = _long_gap_set(17)
    res, off = bs.pack(seqs)
    ref = oracle_triangle(res, off, oracle_matrices[matrix], go, ge)
    ctx.set_scoring(matrix, go, ge)
    ctx.load_sequences(0, res, off)
    for env in ({}, {"BSA_NO_TAG": "1"}, {"BSA_NO_PAIR": "1"}, {"BSA_NO_TAG": "1", "BSA_NO_PAIR": "1"}):
        with monkeypatch.context() as mp:
            for k, v in env.items():
                mp.setenv(k, v)
            scores, nid = ctx.all_vs_all(0)
        bad = np.nonzero((scores != ref["score"]) | (nid != ref["n_identical"]))[0]
        assert len(bad) == 0, (env, bad[:5], scores[bad[:5]], ref["score"][bad[:5]])
        assert ctx.stats()["fallback_pairs"] == 0


def test_score_only_lanes_long_gaps_and_large_gap_open(ctx, oracle_matrices):
    """16-bit biased lanes (two-row step, `h + go` as one 32-bit IMAD): long gaps, and a gap-open
    large enough that a borrow between the halves would show if a biased value fell below |go|."""
    seqs = _long_gap_set(23)
    res, off = bs.pack(seqs)
    for matrix, go, ge in (("BLOSUM62", -10, -1), ("BLOSUM62", -700, -1), ("PAM30", -30, -5)):
        ref = oracle_triangle(res, off, oracle_matrices[matrix], go, ge)
        ctx.set_scoring(matrix, go, ge)
        ctx.load_sequences(0, res, off)
        scores, _ = ctx.align_all_pairs(0, 0, np.arange(len(seqs), dtype=np.uint32), want_identical=False)
        bad = np.nonzero(scores != ref["score"])[0]
        assert len(bad) == 0, (matrix, go, ge, bad[:5], scores[bad[:5]], ref["score"][bad[:5]])
