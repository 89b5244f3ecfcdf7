"""(ii) oracle-vs-oracle: the C restatement against the independent Python one, and the
kernels' packed-lane arithmetic model against both, on seeded random / homologous pairs
including dirty bytes, length-1 sequences and several gap settings."""
import random

import pytest

from bioshell_b200.scoring import ncbi_text
from oracle import c_oracle, pyoracle
from packed_model import frame_align, packed_align, tagged_align, walk_dirs

AA = b"ARNDCQEGHILKMFPSTWYV"
DIRTY = b"ARNDCQEGHILKMFPSTWYVXBZJUO*-_arndx"
GAPS = [(-10, -1), (-10, -2), (-11, -1), (-5, -5), (-1, -1), (-12, -3), (-3, 0), (-1, 0)]


def rand_seq(rng, n, alphabet):
    return bytes(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, alphabet):
    out = bytearray()
    for ch in s:
        u = rng.random()
        if u < 0.05:
            continue
        if u < 0.10:
            out.extend(rand_seq(rng, rng.randint(1, 3), alphabet))
        out.append(rng.choice(alphabet) if rng.random() < 0.3 else ch)
    return bytes(out) or b"A"


def pairs(seed, count, maxlen, alphabet):
    rng = random.Random(seed)
    for k in range(count):
        n = rng.randint(1, maxlen)
        q = rand_seq(rng, n, alphabet)
        t = mutate(rng, q, alphabet) if k % 2 else rand_seq(rng, rng.randint(1, maxlen), alphabet)
        yield q, t


@pytest.mark.parametrize("matrix", ["BLOSUM62", "PAM30"])
def test_c_vs_python_restatement(matrix):
    text = ncbi_text(matrix)
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    k = 0
    for q, t in pairs(11, 400, 40, DIRTY):
        go, ge = GAPS[k % len(GAPS)]
        k += 1
        if ge == 0 and max(len(q), len(t)) < 2:
            continue
        lmax = max(len(q), len(t)) + (k % 7) * 50
        a = c_oracle.align_pair(q, t, sc, ai, go, ge, lmax)
        b = pyoracle.align_pair(q, t, psc, pai, go, ge, lmax)
        for key in ("score", "path", "aligned_q", "aligned_t", "n_identical", "len_q", "len_t"):
            assert a[key] == b[key], (key, q, t, go, ge)


@pytest.mark.parametrize("alphabet", [AA, DIRTY])
def test_packed_lane_model_matches_oracle(alphabet):
    """The 8-instruction packed cell (score|prio|count in one int, no sentinel, eager E/F)
    reproduces score, n_identical AND the full traceback of the literal restatement."""
    text = ncbi_text("BLOSUM62")
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    k = 0
    for q, t in pairs(23, 500, 48, alphabet):
        go, ge = GAPS[k % len(GAPS)]
        k += 1
        if ge == 0 and max(len(q), len(t)) < 2:
            continue
        ref = c_oracle.align_pair(q, t, sc, ai, go, ge, max(len(q), len(t)) + 100 * (k % 3))
        cs = max(len(q), len(t)).bit_length()
        s, nid, _ = packed_align(q, t, psc, pai, go, ge, cs)
        assert (s, nid) == (ref["score"], ref["n_identical"]), (q, t, go, ge)
        s0, _, dirs = packed_align(q, t, psc, pai, go, ge, 0, want_dirs=True)
        assert s0 == ref["score"]
        assert walk_dirs(len(q), len(t), dirs) == ref["path"], (q, t, go, ge)


@pytest.mark.parametrize("alphabet", [AA, DIRTY])
def test_tagged_lane_model_matches_oracle(alphabet):
    """The TAG cell (4 ALU + 3 IMAD: priorities born in place, extend-beats-open through a streak
    field that is cleared at lane boundaries / every R rows) gives the reference's score and
    n_identical for every lane width and clearing period, including the degenerate ones."""
    text = ncbi_text("BLOSUM62")
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    k = 0
    for q, t in pairs(29, 400, 48, alphabet):
        go, ge = GAPS[k % len(GAPS)]
        k += 1
        if ge == 0 and max(len(q), len(t)) < 2:
            continue
        ref = c_oracle.align_pair(q, t, sc, ai, go, ge, max(len(q), len(t)) + 100 * (k % 3))
        cs = max(len(q), len(t)).bit_length()
        for K, R, xb in ((4, 16, 5), (1, 1, 1), (20, 16, 5), (3, 5, 3)):
            s, nid = tagged_align(q, t, psc, pai, go, ge, cs, xb=xb, K=K, R=R)
            assert (s, nid) == (ref["score"], ref["n_identical"]), (q, t, go, ge, K, R)


@pytest.mark.parametrize("alphabet", [AA, DIRTY])
def test_frame_lane_model_matches_oracle(alphabet):
    """The FRAME cell (TAG cell in the moving frame score - (i+j) ge: 4 ALU + 2 IMAD, both extensions free,
    F tie rule through decreasing candidate tags) gives the reference's score and n_identical for every
    lane width and clearing period."""
    text = ncbi_text("BLOSUM62")
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    k = 0
    for q, t in pairs(31, 500, 56, alphabet):
        go, ge = GAPS[k % len(GAPS)]
        k += 1
        if ge == 0 and max(len(q), len(t)) < 2:
            continue
        ref = c_oracle.align_pair(q, t, sc, ai, go, ge, max(len(q), len(t)) + 100 * (k % 3))
        cs = max(len(q), len(t)).bit_length()
        for K, R, xb in ((4, 16, 5), (1, 1, 1), (20, 16, 5), (3, 4, 3), (7, 2, 5)):
            s, nid = frame_align(q, t, psc, pai, go, ge, cs, xb=xb, K=K, R=R)
            assert (s, nid) == (ref["score"], ref["n_identical"]), (q, t, go, ge, K, R)
            # the even-aligned stream of the two-row kernels: a PAD row ahead of odd-length queries,
            # the query starting anywhere inside a block of R rows
            s, nid = frame_align(q, t, psc, pai, go, ge, cs, xb=xb, K=K, R=R, pad=True, phase=(k * 5 + K) % 32)
            assert (s, nid) == (ref["score"], ref["n_identical"]), (q, t, go, ge, K, R, "pad")
            # ... with the E openings tagged by column instead of the streak count (BSA_ETAG), and the count field
            # as wide as the columns-per-lane class needs instead of the template
            for cs2 in (cs, cs + 1):
                s, nid = frame_align(q, t, psc, pai, go, ge, cs2, xb=max(xb, 2), K=K, R=R, pad=True,
                                     phase=(k * 3 + K) % 32, etag=True)
                assert (s, nid) == (ref["score"], ref["n_identical"]), (q, t, go, ge, K, R, "etag")


@pytest.mark.parametrize("alphabet", [AA, DIRTY])
def test_wave_frame_cell_model_matches_oracle(alphabet):
    """The K3 direction-frame cell ((score - (i+j) ge) << 3 | prio << 1 | m, rows padded at the top with a
    zero-score residue, 3 + 1 direction bits per cell) and the run-based walk over its words give the
    reference's score and path glyph for glyph -- every query length mod 4, with and without the rows a
    lane spends ahead of its first row."""
    from packed_model import wave_frame_align
    text = ncbi_text("BLOSUM62")
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    k = 0
    for q, t in pairs(77, 60, 90, alphabet):
        go, ge = GAPS[k % len(GAPS)]
        k += 1
        ref = c_oracle.align_pair(q, t, sc, ai, go, ge, max(len(q), len(t)) + 100 * (k % 3))
        for pre in (0, 4):
            s, path, _ = wave_frame_align(q, t, psc, pai, go, ge, pre=pre)
            assert s == ref["score"], (q, t, go, ge, pre)
            assert path == ref["path"], (q, t, go, ge, pre)


def test_orientation_matters_for_identity_not_score():
    """SURVEY.md 8a note 5: swapping query and template keeps the score but can change the
    path/identity, so the kernels must keep the reference's orientation."""
    sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
    differ = 0
    for q, t in pairs(5, 600, 60, AA):
        a = c_oracle.align_pair(q, t, sc, ai, -10, -1)
        b = c_oracle.align_pair(t, q, sc, ai, -10, -1)
        assert a["score"] == b["score"]
        differ += a["n_identical"] != b["n_identical"]
    assert differ > 0


def test_wave_ring_hand_off_schedule_model():
    """Index logic of the K3 boundary batches (BSA_WAVE_RING; batch 32 is the shipped build, batch 16
    and later fetch points are prepared switches): lane 0 takes entry S at step S for every row, every
    batch is fetched before its first use and only needs entries of that batch to be published."""
    from packed_model import wave_ring_schedule
    for WB, PFs in ((32, (0, 8, 24, 30)), (16, (0, 8, 14))):
        for PF in PFs:
            for X in (1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 100, 1000, 1001):
                used, fetches = wave_ring_schedule(X, WB, PF)
                assert used == {S: S for S in range(X)}, (WB, PF, X)
                for step, first, last, need in fetches:
                    assert need == last <= X and first % WB == 0
                    assert step < first                      # fetched before the first step that uses it
                covered = sorted((f, l) for _, f, l, _ in fetches)
                assert covered[0][0] == 0 and covered[-1][1] == X
                assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


@pytest.mark.parametrize("alphabet", [AA, DIRTY])
def test_frame16_cell_model_matches_oracle(alphabet):
    """The 16-bit score-only cell in the moving frame over the even-aligned stream (4 DPX instructions per two
    cells, constant borders, a PAD row ahead of odd-length queries, padded columns) gives the reference's score."""
    from packed_model import frame16_align
    text = ncbi_text("BLOSUM62")
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    k = 0
    for q, t in pairs(53, 400, 70, alphabet):
        go, ge = GAPS[k % len(GAPS)]
        k += 1
        if ge == 0 and max(len(q), len(t)) < 2:
            continue
        ref = c_oracle.align_pair(q, t, sc, ai, go, ge, max(len(q), len(t)) + 100 * (k % 3))
        for pad, mpad in ((True, 0), (False, 0), (True, 9)):
            s = frame16_align(q, t, psc, pai, go, ge, pad=pad, mpad=mpad)
            assert s == ref["score"], (q, t, go, ge, pad, mpad)
