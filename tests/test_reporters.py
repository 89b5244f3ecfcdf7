"""The reference's printing reporters (alignment_reporter.rs:31-268) and the needleman_wunsh command
line (examples/needleman_wunsh.rs).  Expected text is derived from the reference's format strings;
the known answers the reference states itself (doc-tests) are marked."""
import io

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import needleman_wunsh as nw
from bioshell_b200.alignment import triangle_counts
from bioshell_b200.reporters import (IdentityMatrixReporter, PrintAsFasta, PrintAsPairwise,
                                     ReportWithSequenceIdentity, SimilarityReport, len_ungapped_str)

Q = bs.Sequence.from_str("query", "AL-IV")
T = bs.Sequence.from_str("template", "ALRIV")


def test_len_ungapped_str_doc_tests():
    # sequence.rs:541-542
    assert len_ungapped_str("P-RF") == 3
    assert len_ungapped_str("__PERF_") == 4


def test_identity_matrix_reporter_doc_test():
    # alignment_reporter.rs:176-184
    out = io.StringIO()
    rep = IdentityMatrixReporter(5, False, "stdout", out=out)
    rep.report(Q, bs.Sequence.from_str("tmplt", "ALRIV"))
    assert rep.sequence_index("query") == 0
    assert rep.sequence_index("tmplt") == 1
    assert rep.sequence_index("nobody") is None
    assert rep.n_identical_residues(0, 1) == 4 and rep.n_identical_residues(1, 0) == 4
    assert rep.num_sequences() == 2
    rep.finish()
    rep.finish()                      # written once, like a Drop
    # labels, blank line, then `{:<5}` + ` {:3}` per stored entry (alignment_reporter.rs:235-251)
    assert out.getvalue() == "query\ntmplt\n\nquery   4\ntmplt   4   5\n"


def test_identity_matrix_reporter_orders_by_first_appearance_and_infers_ids(tmp_path):
    path = tmp_path / "matrix.txt"
    a = bs.Sequence.from_str("sp|P12345|AAA_HUMAN first", "MKV-L")
    b = bs.Sequence.from_str("sp|Q67890|BBB_HUMAN second", "MKVAL")
    c = bs.Sequence.from_str("third one", "M-VAL")
    with IdentityMatrixReporter(12, True, str(path)) as rep:
        rep.report(b, a)              # query first: b gets index 0
        rep.report(c, a)
        rep.report(c, b)
        assert [rep.sequence_index(k) for k in ("sp|Q67890|BB", "sp|P12345|AA", "third")] == [0, 1, 2]
        assert rep.n_identical_residues(0, 1) == 4 and rep.n_identical_residues(2, 1) == 3
    lines = path.read_text().split("\n")
    assert lines[:4] == ["sp|Q67890|BB", "sp|P12345|AA", "third", ""]
    assert lines[4:7] == ["sp|Q67890|BB   5", "sp|P12345|AA   4   4", "third          4   3   4"]


def test_print_as_pairwise_layout():
    out = io.StringIO()
    PrintAsPairwise(5, 10, out=out).report(Q, T)          # the reference's doc example (:79-86)
    assert out.getvalue() == ("query     1 AL-IV     5\n"
                              "            || ||\n"
                              "templ     1 ALRIV     6\n"
                              "\n\n")
    # several blocks: the next block starts at from + ungapped - 1 (alignment_reporter.rs:120-121)
    out = io.StringIO()
    q = bs.Sequence.from_str("q", "ABCDEFGH--KL")
    t = bs.Sequence.from_str("a long template name", "ABC-EFGHIJKL")
    PrintAsPairwise(8, 5, out=out).report(q, t)
    assert out.getvalue() == ("q            1 ABCDE     6\n"
                              "               ||| |\n"
                              "a long t     1 ABC-E     5\n"
                              "q            5 FGH--     8\n"
                              "               |||  \n"
                              "a long t     4 FGHIJ     9\n"
                              "q            7 KL     9\n"
                              "               ||\n"
                              "a long t     8 KL    10\n"
                              "\n\n")


def test_print_as_fasta_and_similarity_report():
    out = io.StringIO()
    PrintAsFasta(out=out).report(Q, T)
    assert out.getvalue() == "> query\nAL-IV\n\n> template\nALRIV\n\n"
    out = io.StringIO()
    SimilarityReport(5, False, out=out).report(Q, T)
    assert out.getvalue() == "query templ 100.00 %   4    4    5\n"
    out = io.StringIO()
    SimilarityReport(out=out).report(Q, T)                   # Default: 32 characters, descriptions
    assert out.getvalue() == "query template 100.00 %   4    4    5\n"
    out = io.StringIO()
    a = bs.Sequence.from_str("gi|5524211|gb|AAD44166.1| cytochrome b [Elephas maximus maximus]", "MK-L")
    b = bs.Sequence.from_str("plain name", "MKVI")
    SimilarityReport(14, True, out=out).report(a, b)
    assert out.getvalue() == "gb|AAD44166.1| plain  66.67 %   2    3    4\n"


def test_report_with_sequence_identity_window():
    got = bs.CollectReporter()
    rep = ReportWithSequenceIdentity(50.0, 80.0, got)
    rep.report(bs.Sequence.from_str("a", "MKVL"), bs.Sequence.from_str("b", "MKAA"))      # 50 %: kept (>=)
    rep.report(bs.Sequence.from_str("a", "MKVL"), bs.Sequence.from_str("c", "MKVL"))      # 100 %
    rep.report(bs.Sequence.from_str("a", "MKVLA"), bs.Sequence.from_str("d", "MKVLG"))    # 80 %: kept (<=)
    rep.report(bs.Sequence.from_str("a", "----"), bs.Sequence.from_str("e", "MKVL"))      # 0/0 = NaN: dropped
    assert [t.description() for _, t in got.pairs] == ["b", "d"]
    assert (ReportWithSequenceIdentity.higher_than(30.0, got).min_seq_id,
            ReportWithSequenceIdentity.higher_than(30.0, got).max_seq_id) == (30.0, 100.0)
    assert (ReportWithSequenceIdentity.lower_than(30.0, got).min_seq_id,
            ReportWithSequenceIdentity.lower_than(30.0, got).max_seq_id) == (0.0, 30.0)


# ----------------------------------------------------------------------------- needleman_wunsh
def test_needleman_wunsh_arguments_and_reporter_wiring():
    a = nw.build_parser().parse_args(["-q", "MKV"])
    assert (a.open, a.extend, a.name_width, a.template) == (-10, -2, 20, None)
    rep, matrix = nw.build_reporters(a)
    assert matrix is None and isinstance(rep.reporters[0], SimilarityReport)        # the default report
    a = nw.build_parser().parse_args(["-q", "q.fasta", "-t", "t.fasta", "-o", "-11", "-e", "-1", "--pairwise",
                                      "--identity", "--report-more-similar", "90", "--report-less-similar", "30",
                                      "-w", "12", "--infer-seq-id"])
    rep, matrix = nw.build_reporters(a)
    filt = rep.reporters[0]
    assert isinstance(filt, ReportWithSequenceIdentity) and (filt.min_seq_id, filt.max_seq_id) == (30.0, 90.0)
    assert [type(r) for r in filt.reporter.reporters] == [PrintAsPairwise, SimilarityReport]
    a = nw.build_parser().parse_args(["-q", "q.fasta", "--identity-matrix", "--report-more-similar", "90"])
    rep, matrix = nw.build_reporters(a)
    assert rep.reporters == [matrix]                      # thresholds are ignored with --identity-matrix


def _oracle_replay(oracle_matrices):
    """An `align_all_pairs` that replays the CPU oracle's alignments in the reference's order."""
    from oracle import c_oracle
    sc, ai = oracle_matrices["BLOSUM62"]

    def replay(queries, templates, matrix, go, ge, tri, reporter, ctx=None):
        counts = triangle_counts(queries, templates, tri)
        for t, cnt in enumerate(counts):
            for q in range(int(cnt)):
                one = c_oracle.align_pair(queries[q].as_u8(), templates[t].as_u8(), sc, ai, go, ge)
                reporter.report(bs.Sequence(queries[q].description(), one["aligned_q"]),
                                bs.Sequence(templates[t].description(), one["aligned_t"]))
    return replay


FASTA = (">sp|P10001|AAA_HUMAN one\nMAVRLLKTHLGGSW\n>sp|P10002|BBB_HUMAN two\nMKNITCYLGGW\n"
         ">three\nMAVKNLKTYLGSW\n>four\nMAVRLLKTHLGGSW\n")
NW_ARGS = ["--pairwise", "--identity", "--identity-matrix", "-w", "10", "--infer-seq-id"]


def _run_cli(tmp_path, extra, out):
    p = tmp_path / "q.fasta"
    p.write_text(FASTA)
    assert nw.main(["-q", str(p)] + extra, out=out) == 0


def test_needleman_wunsh_cli_text_against_oracle_replay(tmp_path, oracle_matrices, monkeypatch):
    """The command line's wiring on the CPU: the GPU call is replaced by an oracle replay, the
    expected text comes from driving the reporters by hand with the same alignments."""
    class NoContext:
        def __init__(self, device):
            pass

        def __enter__(self):
            return None

        def __exit__(self, *exc):
            return False
    replay = _oracle_replay(oracle_matrices)
    monkeypatch.setattr(nw, "Context", NoContext)
    monkeypatch.setattr(nw, "align_all_pairs", replay)
    got = io.StringIO()
    _run_cli(tmp_path, NW_ARGS, got)
    assert got.getvalue() == _expected_text(oracle_matrices)
    assert got.getvalue().count(" % ") == 6               # 4 sequences: 6 pairs of the triangle


def _expected_text(oracle_matrices):
    seqs = list(bs.FastaIterator(FASTA))
    want = io.StringIO()
    m = bs.MultiReporter()
    m.add_reporter(PrintAsPairwise(10, 80, out=want))
    m.add_reporter(SimilarityReport(10, True, out=want))
    im = IdentityMatrixReporter(10, True, "stdout", out=want)
    m.add_reporter(im)
    _oracle_replay(oracle_matrices)(seqs, seqs, "BLOSUM62", -10, -2, True, m)
    im.finish()
    return want.getvalue()
