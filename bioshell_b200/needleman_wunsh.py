"""`needleman_wunsh` -- the command line of the reference's bioshell-seq/examples/needleman_wunsh.rs:11-117
on the batched GPU path:  python -m bioshell_b200.needleman_wunsh -q queries.fasta [-t templates.fasta] --pairwise

Same options and defaults; the alignments are computed by the GPU library and replayed into the
reference's reporters in its template-major order (`align_all_pairs`).  `--device` is an addition.
"""
import argparse
import logging
import sys

from .alignment import Context, MultiReporter, align_all_pairs
from .fasta import load_sequences
from .reporters import IdentityMatrixReporter, PrintAsPairwise, ReportWithSequenceIdentity, SimilarityReport

log = logging.getLogger("needleman_wunsh")


def build_parser():
    """needleman_wunsh.rs:11-57"""
    p = argparse.ArgumentParser(prog="needleman_wunsh",
                                description="Calculates global sequence alignment of amino acid sequences")
    p.add_argument("-q", "--query", required=True,
                   help="query sequence(s): either a FASTA string or a name of a file in FASTA format")
    p.add_argument("-t", "--template", default=None,
                   help="template sequence(s): either a FASTA string or a name of a file in FASTA format")
    p.add_argument("-o", "--open", type=int, default=-10, help="gap opening penalty")
    p.add_argument("-e", "--extend", type=int, default=-2, help="gap extension penalty")
    p.add_argument("--pairwise", action="store_true",
                   help="print pairwise alignments for every pair of aligned sequences")
    p.add_argument("--identity", action="store_true", help="print sequence identity report (default)")
    p.add_argument("--identity-matrix", action="store_true", help="print sequence identity as a triangular matrix")
    p.add_argument("-w", "--name-width", type=int, default=20, help="length of a sequence name to print")
    p.add_argument("--report-more-similar", type=float, default=None,
                   help="report only the alignments with sequence identity above the given threshold")
    p.add_argument("--report-less-similar", type=float, default=None,
                   help="report only the alignments with sequence identity below the given threshold")
    p.add_argument("--infer-seq-id", action="store_true",
                   help="print the sequence ID instead of the sequence description")
    p.add_argument("-v", "--verbose", action="store_true", help="be more verbose")
    p.add_argument("--device", type=int, default=0, help="(addition) CUDA device to run on")
    return p


def build_reporters(args, out=None):
    """needleman_wunsh.rs:75-98; returns (reporter, the IdentityMatrixReporter or None)"""
    multi = MultiReporter()
    matrix = None
    if args.pairwise:
        multi.add_reporter(PrintAsPairwise(args.name_width, 80, out=out))
    if args.identity:
        multi.add_reporter(SimilarityReport(args.name_width, args.infer_seq_id, out=out))
    if args.identity_matrix:
        matrix = IdentityMatrixReporter(args.name_width, args.infer_seq_id, "stdout", out=out)
        multi.add_reporter(matrix)
    if multi.count_reporters() == 0:
        multi.add_reporter(SimilarityReport(args.name_width, args.infer_seq_id, out=out))
    if not args.identity_matrix:
        lo = args.report_more_similar if args.report_more_similar is not None else -0.01
        hi = args.report_less_similar if args.report_less_similar is not None else 100.1
        if lo > hi:
            lo, hi = hi, lo
        if lo > -0.01 or hi < 100.1:
            m = MultiReporter()
            m.add_reporter(ReportWithSequenceIdentity(lo, hi, multi))
            multi = m
    return multi, matrix


def main(argv=None, out=None):
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO,
                        format="[%(levelname)s %(name)s] %(message)s")
    reporter, matrix = build_reporters(args, out)
    queries = load_sequences(args.query, "query")
    if not queries:
        log.warning("No sequences found in the query set. Exiting.")
        return 0
    with Context(args.device) as ctx:
        if args.template is not None:
            templates = load_sequences(args.template, "template")
            if not templates:
                log.warning("No sequences found in the templates set. Exiting.")
                return 0
            align_all_pairs(queries, templates, "BLOSUM62", args.open, args.extend, False, reporter, ctx=ctx)
        else:
            align_all_pairs(queries, queries, "BLOSUM62", args.open, args.extend, True, reporter, ctx=ctx)
    if matrix is not None:
        matrix.finish()          # the reference writes it when the reporter is dropped at the end of main
    return 0


if __name__ == "__main__":
    sys.exit(main())
