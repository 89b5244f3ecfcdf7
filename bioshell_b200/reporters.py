"""The reference's printing reporters, for callers that replay the batched results through an
`AlignmentReporter` (`align_all_pairs`): bioshell-seq/src/alignment/alignment_reporter.rs:31-268.
Host text only; every reporter takes an `out` stream (default: the process' stdout) where the
reference calls `println!`.

  ReportWithSequenceIdentity   alignment_reporter.rs:31-64
  PrintAsFasta                 alignment_reporter.rs:66-73
  PrintAsPairwise              alignment_reporter.rs:75-126
  SimilarityReport             alignment_reporter.rs:128-170
  IdentityMatrixReporter       alignment_reporter.rs:172-268 (its `Drop` is `finish()` here)
"""
import sys

from .alignment import AlignmentReporter, AlignmentStatistics
from .sequence import count_identical, len_ungapped
from .sequence_id import LabelStyle, sequence_label

_U64 = (1 << 64) - 1


def len_ungapped_str(s):
    """sequence.rs:544-546"""
    return sum(1 for c in s if c != "-" and c != "_")


def _display(seq, width=0):
    """`impl Display for Sequence` (display_sequence.rs:4-9)"""
    s = seq.to_string(0)
    if width:
        s = "\n".join(s[k:k + width] for k in range(0, len(s), width))
    return "> %s\n%s\n" % (seq.description(), s)


def _description_n(seq, n):
    """sequence.rs:99-104"""
    d = seq.description()
    return d if n == 0 else d[:min(len(d), n)]


def _label_style(header_width, infer_seq_id):
    """alignment_reporter.rs:147-152,199-203"""
    return LabelStyle.FullId(True, header_width) if infer_seq_id else LabelStyle.Description(header_width)


def out_writer(out_fname, if_append=False):
    """bioshell-core/src/io/utils.rs:35-58; returns (stream, close_it)"""
    if out_fname in ("", "stdout"):
        return sys.stdout, False
    if out_fname == "stderr":
        return sys.stderr, False
    return open(out_fname, "a" if if_append else "w"), True


class ReportWithSequenceIdentity(AlignmentReporter):
    """Passes on only the alignments whose identity lies in [min_seq_id, max_seq_id]."""

    def __init__(self, min_seq_id, max_seq_id, reporter):
        self.min_seq_id, self.max_seq_id, self.reporter = float(min_seq_id), float(max_seq_id), reporter

    @classmethod
    def higher_than(cls, min_seq_id, reporter):
        return cls(min_seq_id, 100.0, reporter)

    @classmethod
    def lower_than(cls, max_seq_id, reporter):
        return cls(0.0, max_seq_id, reporter)

    def report(self, aligned_query, aligned_template):
        n_identical = count_identical(aligned_query, aligned_template)
        mn = min(len_ungapped(aligned_query), len_ungapped(aligned_template))
        # f64 division: 0/0 is NaN and fails both comparisons, x/0 is +inf (alignment_reporter.rs:58-62)
        seq_id = (n_identical / mn * 100.0) if mn else (float("nan") if n_identical == 0 else float("inf"))
        if self.min_seq_id <= seq_id <= self.max_seq_id:
            self.reporter.report(aligned_query, aligned_template)


class PrintAsFasta(AlignmentReporter):
    def __init__(self, out=None):
        self.out = out

    def report(self, aligned_query, aligned_template):
        (self.out or sys.stdout).write("%s\n%s\n" % (_display(aligned_query), _display(aligned_template)))


class PrintAsPairwise(AlignmentReporter):
    """Blocks of `alignment_width` columns: query line, a line of '|' under equal symbols, template
    line, each with its running residue numbers exactly as the reference counts them (the next
    block starts at `from + ungapped - 1`, alignment_reporter.rs:120-121)."""

    def __init__(self, seq_name_width, alignment_width, out=None):
        self.seq_name_width, self.alignment_width, self.out = seq_name_width, alignment_width, out

    def report(self, aligned_query, aligned_template):
        out = self.out or sys.stdout
        w = self.seq_name_width
        q_name = _description_n(aligned_query, w).ljust(w)
        t_name = _description_n(aligned_template, w).ljust(w)
        qs, ts = aligned_query.to_string(0), aligned_template.to_string(0)
        aw = self.alignment_width
        q_chunks = [qs[k:k + aw] for k in range(0, len(qs), aw)]
        t_chunks = [ts[k:k + aw] for k in range(0, len(ts), aw)]
        q_from = t_from = 1
        num_spacer = " " * 5
        for q, t in zip(q_chunks, t_chunks):
            mid_name = " " * len(q_name)
            middle = "".join("|" if c1 == c2 else " " for c1, c2 in zip(q, t))
            q_add, t_add = len_ungapped_str(q), len_ungapped_str(t)
            out.write("%s %5d %s %5d\n%s %s %s\n%s %5d %s %5d\n" % (
                q_name, q_from, q, (q_from + q_add) & _U64, mid_name, num_spacer, middle,
                t_name, t_from, t, (t_from + t_add) & _U64))
            q_from = (q_from + q_add - 1) & _U64      # usize arithmetic of a release build
            t_from = (t_from + t_add - 1) & _U64
        out.write("\n\n")


class SimilarityReport(AlignmentReporter):
    """One `AlignmentStatistics` line per alignment."""

    def __init__(self, header_width=32, infer_seq_id=False, out=None):
        self.label_style = _label_style(header_width, infer_seq_id)
        self.out = out

    def report(self, aligned_query, aligned_template):
        stats = AlignmentStatistics.from_sequences(aligned_query, aligned_template, self.label_style)
        (self.out or sys.stdout).write("%s\n" % stats)


class IdentityMatrixReporter(AlignmentReporter):
    """Lower-triangular matrix of identical-residue counts (ungapped lengths on the diagonal), the
    sequences in the order they were first reported (query before template).  The reference writes
    the matrix when the reporter is dropped; call `finish()` (or use it as a context manager)."""

    def __init__(self, header_width, infer_seq_id, out_fname, out=None):
        self.out_fname = out_fname
        self.header_width = header_width
        self.label_style = _label_style(header_width, infer_seq_id)
        self.identity_matrix = []
        self.sequence_order = {}
        self._out = out
        self._finished = False

    def sequence_index(self, seq_name):
        return self.sequence_order.get(seq_name)

    def num_sequences(self):
        return len(self.sequence_order)

    def n_identical_residues(self, query_idx, tmplt_idx):
        if query_idx < tmplt_idx:
            return self.identity_matrix[tmplt_idx][query_idx]
        return self.identity_matrix[query_idx][tmplt_idx]

    def _index(self, name, aligned):
        if name not in self.sequence_order:
            self.sequence_order[name] = len(self.identity_matrix)
            self.identity_matrix.append([0] * len(self.sequence_order))
            idx = len(self.sequence_order) - 1
            self.identity_matrix[idx][idx] = len_ungapped(aligned)
        return self.sequence_order[name]

    def report(self, aligned_query, aligned_template):
        q_idx = self._index(sequence_label(aligned_query.description(), self.label_style), aligned_query)
        t_idx = self._index(sequence_label(aligned_template.description(), self.label_style), aligned_template)
        n_identical = count_identical(aligned_query, aligned_template)
        if t_idx > q_idx:
            self.identity_matrix[t_idx][q_idx] = n_identical
        else:
            self.identity_matrix[q_idx][t_idx] = n_identical

    def finish(self):
        """`impl Drop` (alignment_reporter.rs:226-252): the labels, a blank line, then one row per
        sequence: the label left-justified to `header_width` and the counts as ' %3d'."""
        if self._finished:
            return
        self._finished = True
        if self._out is not None:
            stream, close_it = self._out, False
        else:
            stream, close_it = out_writer(self.out_fname, False)
        keys = sorted(self.sequence_order.items(), key=lambda kv: kv[1])
        for key, _ in keys:
            stream.write("%s\n" % key)
        stream.write("\n")
        for key, idx in keys:
            stream.write(key.ljust(self.header_width))
            if idx < len(self.identity_matrix):
                for v in self.identity_matrix[idx]:
                    stream.write(" %3d" % v)
            stream.write("\n")
        if close_it:
            stream.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.finish()
