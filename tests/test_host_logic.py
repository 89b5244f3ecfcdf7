"""CPU-side checks: the C ABI loads and exports every declared symbol, the product's NCBI
parser agrees with the oracle's, the host mirrors of the reference helpers behave like the
reference's doc-tests, and the synthetic generator is deterministic.  No GPU compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import _lib, synth
from bioshell_b200.scoring import SubstitutionMatrixList, ncbi_text
from oracle import c_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "bioshell_align.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(bsa_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    out = os.popen("nm -D --defined-only %s" % _lib.LIB_PATH).read()
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name


def test_no_cpu_fallback_without_a_device():
    L = _lib.lib()
    if L.bsa_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(bs.BsaError) as e:
        bs.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_parser_matches_oracle_parser(oracle_matrices):
    for name in SubstitutionMatrixList.ALL:
        m = bs.SubstitutionMatrix.load(name)
        sc, ai = oracle_matrices[name]
        assert np.array_equal(m.score, sc) and np.array_equal(m.aa_indexes, ai), name


def test_parser_format_rules():
    """substitution_matrix.rs:105-108,113-116: '#'/' ' lines skipped, < 23 tokens is an error,
    non-integer token is an error, CRLF tolerated, rows after the 20th ignored."""
    good = ncbi_text("BLOSUM62")
    ref = bs.SubstitutionMatrix.load("BLOSUM62")
    crlf = bs.SubstitutionMatrix.ncbi_matrix_from_buffer(good.replace("\n", "\r\n"))
    assert np.array_equal(crlf.score, ref.score)
    lines = good.split("\n")
    short = "\n".join(lines[:5] + ["A 1 2 3"] + lines[5:])
    with pytest.raises(bs.BsaError):
        bs.SubstitutionMatrix.ncbi_matrix_from_buffer(short)
    with pytest.raises(c_oracle.OracleError):
        c_oracle.parse_ncbi(short)
    bad = good.replace(" 11 ", " 1x ", 1)
    with pytest.raises(bs.BsaError):
        bs.SubstitutionMatrix.ncbi_matrix_from_buffer(bad)
    with pytest.raises(c_oracle.OracleError):
        c_oracle.parse_ncbi(bad)
    # an asymmetric file: the later row wins both mirrored cells (substitution_matrix.rs:117-118)
    rows = [l for l in lines if l and l[0] not in "# "]
    toks = rows[1].split()          # row R, column A
    toks[1] = "7"
    asym = good.replace(rows[1], " ".join(toks))
    a = bs.SubstitutionMatrix.ncbi_matrix_from_buffer(asym)
    sc, _ = c_oracle.parse_ncbi(asym)
    assert np.array_equal(a.score, sc)
    assert a.score_by_aa("A", "R") == 7 and a.score_by_aa("R", "A") == 7


def test_matrix_kats_through_product_api(kats):
    for entry in kats["matrix_values"]:
        m = bs.SubstitutionMatrix.load(entry["matrix"])
        for a, b, v in entry["pairs"]:
            assert m.score_by_aa(a, b) == v
    m = bs.SubstitutionMatrix.load("BLOSUM62")
    assert m.aa_index("A") == 0 and m.aa_index("X") == 20 and m.aa_index("B") == 0 and m.aa_index("-") == 0
    with pytest.raises(IndexError):
        m.aa_index(255)


def test_path_expansion_and_statistics_mirrors(kats):
    p = kats["path_expansion"]
    assert bs.aligned_strings(p["path"], p["query"], p["template"], "-") == (p["aligned_query"], p["aligned_template"])
    q, t = bs.Sequence("query", p["query"]), bs.Sequence("template", p["template"])
    aq, at = bs.aligned_sequences(p["path"], q, t, "-")
    assert aq.description() == "query" and at.description() == "template"
    assert aq.to_string(0) == p["aligned_query"] and at.to_string(0) == p["aligned_template"]
    s = kats["alignment_statistics"]
    st = bs.AlignmentStatistics.from_strings("query", s["aligned_query"], "templ", s["aligned_template"])
    assert (st.query_length, st.template_length, st.n_identical) == (s["query_length"], s["template_length"], s["n_identical"])
    assert str(st) == "query templ  48.00 %  12   25   29"
    h = kats["identity_helpers"]
    for x, y, v in h["count_identical"]:
        assert bs.count_identical(x, y) == v
    for x, v in h["len_ungapped"]:
        assert bs.len_ungapped(x) == v
    with pytest.raises(ValueError):
        bs.count_identical("AB", "A")


def test_triangle_counts_follow_the_break_rule():
    seqs = [bs.Sequence("a", "ARND"), bs.Sequence("b", "ARNDC"), bs.Sequence("c", "WW"),
            bs.Sequence("d", "ARND"), bs.Sequence("a", "ARND")]
    assert bs.triangle_counts(seqs, seqs, True).tolist() == [0, 1, 2, 3, 0]   # last equals the first
    assert bs.triangle_counts(seqs, seqs, False).tolist() == [5] * 5
    other = [bs.Sequence("z", "KK"), bs.Sequence("c", "WW")]
    assert bs.triangle_counts(seqs, other, True).tolist() == [5, 2]


def test_pair_results_identity_is_f64_then_f32():
    r = bs.PairResults(np.array([1, 2, 3], np.int32), np.array([12, 1, 0], np.uint32), [0, 1, 2],
                       np.array([25, 3, 7]), np.array([29, 25, 3]))
    q, t = r.pair_indices()
    assert q.tolist() == [0, 0, 1] and t.tolist() == [1, 2, 2]
    got = r.percent_identity()
    assert got.dtype == np.float32
    exp = [np.float32(12 / 25 * 100.0), np.float32(1 / 3 * 100.0), np.float32(0.0)]
    assert got.tolist() == [float(x) for x in exp]
    assert r.index(1, 2) == 2
    with pytest.raises(KeyError):
        r.index(1, 1)


def test_ungapped_lengths_vectorised():
    res, off = bs.pack([b"P-RF", b"__PERF_", b"", b"ACD"])
    assert bs.ungapped_lengths(res, off).tolist() == [3, 4, 0, 3]


def test_synth_is_deterministic_and_shaped():
    r1, o1 = synth.generate(300, seed=1002)
    r2, o2 = synth.generate(300, seed=1002)
    assert np.array_equal(r1, r2) and np.array_equal(o1, o2)
    r3, _ = synth.generate(300, seed=1003)
    assert not np.array_equal(r1[:1000], r3[:1000])
    lens = np.diff(o1.astype(np.int64))
    assert lens.min() >= 30 and lens.max() <= 4000
    assert set(np.unique(r1).tolist()) <= set(b"ARNDCQEGHILKMFPSTWYV")
    r, o = synth.config("cfg1")
    lens = np.diff(o.astype(np.int64))
    assert len(lens) == 1000 and lens.min() >= 50 and lens.max() <= 500
    # golden checksum so the C generator cannot drift silently
    assert int(o[-1]) == 281827 and int(r.astype(np.uint64).sum()) == int(np.frombuffer(r.tobytes(), np.uint8).astype(np.uint64).sum())


def test_alignment_path_and_step_types():
    # alignment_path.rs:96-100 (doc-test), :50-62, :81-84
    path = bs.AlignmentPath.try_from("**-**")
    assert path.to_string() == "**-**" and path == "**-**"
    assert list(path.iter()) == [bs.AlignmentStep.Match, bs.AlignmentStep.Match, bs.AlignmentStep.Horizontal,
                                 bs.AlignmentStep.Match, bs.AlignmentStep.Match]
    assert bs.aligned_strings(path, "ALIV", "ALRIV", "-") == ("AL-IV", "ALRIV")     # alignment_path.rs:153-159
    assert [str(s) for s in (bs.AlignmentStep.Horizontal, bs.AlignmentStep.Vertical, bs.AlignmentStep.Match)] == ["-", "|", "*"]
    assert bs.AlignmentStep.try_from(ord("|")) is bs.AlignmentStep.Vertical
    with pytest.raises(ValueError, match="Invalid value for AlignmentStep"):
        bs.AlignmentPath.try_from("**x*")
    assert bs.AlignmentPath.from_attrs([bs.AlignmentStep.Match, bs.AlignmentStep.Vertical]) == "*|"


def test_every_entry_point_rejects_a_null_context_without_touching_a_device():
    """Argument validation only (no compute, no GPU needed): a NULL context is BSA_ERR_BAD_ARG
    everywhere, never a crash; bsa_destroy(NULL) and bsa_host_free_pinned(NULL) are no-ops."""
    import ctypes as C
    L = _lib.lib()
    null = C.c_void_p(None)
    bad = L.bsa_set_scoring(null, None, None, -10, -1)
    assert bad < 0
    calls = [
        lambda: L.bsa_load_sequences(null, 0, None, None, 0),
        lambda: L.bsa_plan_shards(null, 0, 0, None, 1, None),
        lambda: L.bsa_align_all_pairs(null, 0, 0, None, 0, 0, 3, None, None, None),
        lambda: L.bsa_all_vs_all(null, 0, 3, None, None),
        lambda: L.bsa_one_vs_many(null, 0, 1, 1, None, None),
        lambda: L.bsa_align_pairs_paths(null, 0, 0, None, None, 0, None, None, None, None),
        lambda: L.bsa_local_align_pairs(null, 0, 0, None, None, 0, None, None, None, None, None, None, None),
        lambda: L.bsa_hclust(null, 0, None, 0, 0, None, None, None),
        lambda: L.bsa_get_stats(null, None),
        lambda: L.bsa_measure_int_peak(null, 0, None, None),
    ]
    for call in calls:
        assert call() == bad
    L.bsa_destroy(null)
    L.bsa_host_free_pinned(null)
