"""Pins BOTH restatements of the reference (oracle/bioshell_oracle.c and
oracle/pyoracle.py) to the reference's own known-answer tests (tests/golden/ref_kats.json,
each entry citing the reference file:line)."""
import numpy as np
import pytest

from bioshell_b200.scoring import ncbi_text
from oracle import c_oracle, pyoracle


def _impls(name):
    text = ncbi_text(name)
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    return [("c", lambda q, t, go, ge, lmax: c_oracle.align_pair(q, t, sc, ai, go, ge, lmax)),
            ("py", lambda q, t, go, ge, lmax: pyoracle.align_pair(q, t, psc, pai, go, ge, lmax))]


def test_global_aligner_kats(kats):
    g = kats["global_aligner"]
    for label, align in _impls(g["matrix"]):
        for c in g["cases"]:
            r = align(c["query"].encode(), c["template"].encode(), g["gap_open"], g["gap_extend"],
                      g["aligner_capacity"])
            assert r["score"] == c["score"], label
            assert r["path"] == c["alignment"], label
            assert r["aligned_q"].decode() == c["aligned_query"], label
            assert r["aligned_t"].decode() == c["aligned_template"], label
            # swapped orientation (test_aligners.rs:48-56)
            r2 = align(c["template"].encode(), c["query"].encode(), g["gap_open"], g["gap_extend"],
                       g["aligner_capacity"])
            assert r2["score"] == c["score"], label
            assert r2["aligned_t"].decode() == c["aligned_query"], label
            assert r2["aligned_q"].decode() == c["aligned_template"], label


def test_matrix_values(kats):
    for entry in kats["matrix_values"]:
        text = ncbi_text(entry["matrix"])
        for sc, ai in (c_oracle.parse_ncbi(text), pyoracle.parse_ncbi(text)):
            for a, b, v in entry["pairs"]:
                assert sc[int(ai[ord(a)]) * 21 + int(ai[ord(b)])] == v, (entry["matrix"], a, b)


def test_both_parsers_agree_on_all_matrices(oracle_matrices):
    for name, (sc, ai) in oracle_matrices.items():
        psc, pai = pyoracle.parse_ncbi(ncbi_text(name))
        assert list(sc) == psc and list(ai) == pai, name
        # unknown bytes -> index 0; X -> 20 (substitution_matrix.rs:38,130)
        assert ai[ord("B")] == 0 and ai[ord("-")] == 0 and ai[ord("a")] == 0 and ai[ord("X")] == 20
        assert sc[20 * 21 + 20] == -1
        assert np.array_equal(sc.reshape(21, 21), sc.reshape(21, 21).T)


def test_similarity_score_kats(kats, oracle_matrices):
    s = kats["similarity_score"]
    sc, ai = oracle_matrices[s["matrix"]]
    for c in s["position_scores"]:
        q, t = c["query"].encode(), c["template"].encode()
        assert sc[int(ai[q[c["i"]]]) * 21 + int(ai[t[c["j"]]])] == c["score"]
    seqs = [x.encode() for x in s["diagonal_sums"]["sequences"]]
    got = []
    for i in range(len(seqs)):
        for j in range(i + 1, len(seqs)):
            got.append(int(sum(sc[int(ai[seqs[i][k]]) * 21 + int(ai[seqs[j][k]])] for k in range(5))))
    assert got == s["diagonal_sums"]["expected"]


def test_path_expansion_and_statistics(kats):
    p = kats["path_expansion"]
    aq, at = pyoracle.expand(p["path"], p["query"].encode(), p["template"].encode())
    assert aq.decode() == p["aligned_query"] and at.decode() == p["aligned_template"]
    import ctypes as C
    L = c_oracle.lib()
    path = np.frombuffer(p["path"].encode(), np.uint8).copy()
    q = np.frombuffer(p["query"].encode(), np.uint8).copy()
    t = np.frombuffer(p["template"].encode(), np.uint8).copy()
    oq, ot = np.zeros(len(path), np.uint8), np.zeros(len(path), np.uint8)
    nid, lq, lt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert L.orc_expand_and_count(path.ctypes.data, len(path), q.ctypes.data, len(q), t.ctypes.data, len(t),
                                  ord("-"), oq.ctypes.data, ot.ctypes.data, C.byref(nid), C.byref(lq),
                                  C.byref(lt)) == 0
    assert oq.tobytes().decode() == p["aligned_query"] and ot.tobytes().decode() == p["aligned_template"]
    assert (nid.value, lq.value, lt.value) == (4, 4, 5)

    s = kats["alignment_statistics"]
    a, b = s["aligned_query"].encode(), s["aligned_template"].encode()
    assert pyoracle.count_identical(a, b) == s["n_identical"]
    assert pyoracle.len_ungapped(a) == s["query_length"] and pyoracle.len_ungapped(b) == s["template_length"]
    pct = L.orc_percent_identity(s["n_identical"], s["query_length"], s["template_length"])
    assert "%6.2f %%" % pct == s["percent"]

    h = kats["identity_helpers"]
    for x, y, v in h["count_identical"]:
        assert pyoracle.count_identical(x.encode(), y.encode()) == v
    for x, v in h["len_ungapped"]:
        assert pyoracle.len_ungapped(x.encode()) == v


def test_all_pairs_order_and_triangle_rule(oracle_matrices):
    """alignment_protocols.rs:94-102: t-major, break at the first query equal (description AND
    bytes) to the template."""
    sc, ai = oracle_matrices["BLOSUM62"]
    seqs = [b"ARND", b"ARNDC", b"WWYV", b"ARND", b"K"]
    descs = ["a", "b", "c", "d", "e"]
    S = c_oracle.SeqSet(seqs, descs)
    r = c_oracle.align_all_pairs(S, S, sc, ai, -10, -1, True)
    expect = pyoracle.all_pairs_order(list(zip(descs, seqs)), list(zip(descs, seqs)), True)
    assert list(zip(r["q"].tolist(), r["t"].tolist())) == expect
    assert expect == [(q, t) for t in range(5) for q in range(t)]
    # same bytes but a different description is NOT equal -> no early break at (0,3)
    assert (0, 3) in expect and (3, 3) not in expect
    # duplicate record (same description and bytes): the loop for t=3 breaks at q=0
    descs2 = ["a", "b", "c", "a", "e"]
    S2 = c_oracle.SeqSet(seqs, descs2)
    r2 = c_oracle.align_all_pairs(S2, S2, sc, ai, -10, -1, True)
    expect2 = pyoracle.all_pairs_order(list(zip(descs2, seqs)), list(zip(descs2, seqs)), True)
    assert list(zip(r2["q"].tolist(), r2["t"].tolist())) == expect2
    assert not any(t == 3 for _, t in expect2)
    # rectangle: every pair including self pairs
    r3 = c_oracle.align_all_pairs(S, S, sc, ai, -10, -1, False)
    assert r3["n_pairs"] == 25
    # per-pair values equal the single-pair entry point
    for k in range(r["n_pairs"]):
        one = c_oracle.align_pair(seqs[r["q"][k]], seqs[r["t"][k]], sc, ai, -10, -1, lmax=5)
        assert one["score"] == r["score"][k] and one["n_identical"] == r["n_identical"][k]
    # threads do not change anything
    r4 = c_oracle.align_all_pairs(S, S, sc, ai, -10, -1, True, n_threads=3, clear_mode=0)
    assert np.array_equal(r4["score"], r["score"]) and np.array_equal(r4["identity"], r["identity"])


def test_empty_input_panics_like_reference(oracle_matrices):
    sc, ai = oracle_matrices["BLOSUM62"]
    S = c_oracle.SeqSet([b"AR"], ["a"])
    E = c_oracle.SeqSet([], [])
    with pytest.raises(c_oracle.OracleError):
        c_oracle.align_all_pairs(E, S, sc, ai, -10, -1, False)
