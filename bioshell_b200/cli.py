"""`cluster_sequences` -- the command line of the reference's bin/cluster_sequences.rs:18-74,133-261
on the batched GPU path:  python -m bioshell_b200.cli <in.fasta> --single-link -c 40

Same positional argument, options, defaults and output files as the reference binary.
Additions: `--device` (which GPU) and `--reference-compat` (the default) / `--symmetric` (mirror
the identity matrix before clustering; without it the matrix stays exactly as the reference's
reporter fills it, upper triangle only -- SURVEY.md 3.1 note).  There is no CPU fallback: without a
B200 the alignment call fails.
"""
import argparse
import logging
import os
import sys
import time

from . import clustering
from .alignment import Context
from .bucket_clustering import bucket_clustering_n
from .clustering import format_fasta
from .fasta import load_sequences

log = logging.getLogger("cluster_sequences")


def build_parser():
    """bin/cluster_sequences.rs:18-74 (clap derives --kebab-case names from the field names)"""
    p = argparse.ArgumentParser(prog="cluster_sequences",
                                description="Cluster amino acid sequences by sequence identity")
    p.add_argument("infile", help="input file in FASTA format")
    p.add_argument("-o", "--open", type=int, default=-10, help="gap opening penalty")
    p.add_argument("-e", "--extend", type=int, default=-2, help="gap extension penalty")
    p.add_argument("--detect-outliers", type=float, default=None,
                   help="don't cluster the sequences, detect outliers instead; an outlier is a sequence for which "
                        "no other sequence is within a given sequence identity fraction")
    p.add_argument("--single-link", action="store_true", help="use the single linkage clustering")
    p.add_argument("--complete-link", action="store_true", help="use the complete linkage clustering")
    p.add_argument("--average-link", action="store_true", help="use the average linkage clustering")
    p.add_argument("-c", "--identity-cutoff", type=float, default=None,
                   help="writes clusters created by stopping the clustering at a given sequence identity fraction")
    p.add_argument("-m", "--medoids", action="store_true",
                   help="print the representative sequence for each cluster (i.e. the medoid)")
    p.add_argument("-b", "--bucket-clustering", type=float, default=None,
                   help="clusters sequences into buckets at the given sequence identity fraction")
    p.add_argument("--n-threads", type=int, default=1, help="number of threads for the bucket clustering")
    p.add_argument("--prefix", default="", help="prefix to add to the output files")
    p.add_argument("--fasta", default=None,
                   help="writes the input sequences reordered according to the clustering tree")
    p.add_argument("--distance-matrix", default=None,
                   help="writes the distance matrix ordered by the clustering tree")
    p.add_argument("-w", "--name-width", type=int, default=20,
                   help="length of a sequence name to print; longer names will be trimmed that size")
    p.add_argument("--sequence-width", type=int, default=80,
                   help="length of a sequence itself to print; use 0 to print the whole sequence in a single line")
    p.add_argument("-v", "--verbose", action="store_true", help="be more verbose")
    p.add_argument("--device", type=int, default=0, help="(addition) CUDA device to run on")
    g = p.add_mutually_exclusive_group()
    g.add_argument("--reference-compat", action="store_true",
                   help="(addition, the default) keep the identity matrix exactly as the reference's reporter "
                        "fills it: only [q][t], q < t")
    g.add_argument("--symmetric", action="store_true",
                   help="(addition) mirror the identity matrix before clustering instead of keeping the "
                        "reference's upper-triangle-only matrix")
    return p


def _can_create_file(path):
    """bioshell-core io::can_create_file: try to create (and remove) the file"""
    try:
        with open(path, "w"):
            pass
        os.remove(path)
        return True
    except OSError:
        return False


def main(argv=None, out=None):
    out = out or sys.stdout
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO,
                        format="[%(levelname)s %(name)s] %(message)s")
    sequences = load_sequences(args.infile, "")                                            # :143
    with Context(args.device) as ctx:
        if args.bucket_clustering is not None:                                             # :145-173
            probe = "%s%s" % (args.prefix, "bsa_write_probe_%d" % os.getpid())
            if not _can_create_file(probe):
                log.error("Can't write with prefix %s", args.prefix)
                return 0
            t0 = time.perf_counter()
            log.info("Bucket clustering of %d sequences with cutoff %s", len(sequences), args.bucket_clustering)
            clusters = bucket_clustering_n(sequences, args.bucket_clustering, args.n_threads, ctx=ctx)
            log.info("%d sequences clustered in %.3fs", len(sequences), time.perf_counter() - t0)
            for i, cluster in enumerate(clusters):
                with open("%scluster_%d-%d.fasta" % (args.prefix, i, len(cluster)), "w") as fh:
                    for s in cluster:
                        fh.write(format_fasta(s, args.sequence_width) + "\n")
            return 0

        if args.detect_outliers is not None:                                               # :180-187
            res = clustering.cluster_sequences(sequences, None, args.open, args.extend,
                                               detect_outliers=args.detect_outliers, name_width=args.name_width,
                                               reference_compat=not args.symmetric, ctx=ctx, write_files=False)
            for i in res["outliers"]:
                out.write(format_fasta(sequences[i]) + "\n")
            return 0

        picked = [l for l, on in ((clustering.single_link, args.single_link),
                                  (clustering.complete_link, args.complete_link),
                                  (clustering.average_link, args.average_link)) if on]
        t0 = time.perf_counter()
        res = clustering.cluster_sequences(                                                # :173-256
            sequences, picked[0] if len(picked) == 1 else None, args.open, args.extend,
            identity_cutoff=args.identity_cutoff, medoids=args.medoids, prefix=args.prefix,
            sequence_width=args.sequence_width, distance_matrix=args.distance_matrix, fasta=args.fasta,
            name_width=args.name_width, reference_compat=not args.symmetric, ctx=ctx)
        log.info("%d sequences aligned and clustered in %.3fs", len(sequences), time.perf_counter() - t0)
        if res["clusters"] is not None:
            log.info("%d clusters retrieved for seq_id %s", len(res["clusters"]), args.identity_cutoff)
    return 0


if __name__ == "__main__":
    sys.exit(main())
