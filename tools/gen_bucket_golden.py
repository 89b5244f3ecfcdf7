#!/usr/bin/env python3
"""Golden clusterings for bucket clustering (tests/golden/bucket_kats.json), made in the build container
where /root/reference is readable.  Inputs are the reference's own test sequences:
  * FDX: the 9 ferredoxins of bioshell-seq/tests/test_bucket_clustering.rs:8-34 (the test that only runs)
  * 4Fe-4S: bioshell-seq/tests/test_files/4Fe-4S-example.fasta (6 sequences)
For each set the pairwise n_identical of GlobalAligner(BLOSUM62, -11, -1) with query = row sequence and
template = column sequence comes from the PINNED alignment oracle (oracle/bioshell_oracle.c), and the
k-mer verdict of sequence_identity (:272-292: A = certainly above, B = certainly below, I = inconclusive)
comes from k-mers kept as plain substrings in Python sets -- a computation that shares nothing with the
5-bit codes of the restatement.  The expected clustering is derived from those two tables alone by the
greedy rule of bucket_clustering.rs:209-270: longest first (stable); a candidate joins the FIRST
representative whose verdict is A, or is I with n_identical / shorter length >= level; else it becomes a
representative.  Both tables are stored next to the clusters so that every expected cluster can be
checked by hand.  (At low levels the reference's "lower bound" is no bound at all -- word size 1 puts
everything into one bucket -- and the golden follows the reference, not the alignments.)  The tests then require the k-mer accelerated restatement
(oracle/pybucket.py), the Python product and the C++ driver to reproduce these clusters."""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioshell_b200.scoring import ncbi_text  # noqa: E402
from oracle import c_oracle  # noqa: E402

REF = "/root/reference/bioshell-seq/tests"


def fasta(text):
    seqs, cur = [], None
    for line in text.splitlines():
        if line.startswith(">"):
            cur = []
            seqs.append(cur)
        elif cur is not None:
            cur.append(line.strip())
    return ["".join(s) for s in seqs]


def main():
    rs = open(os.path.join(REF, "test_bucket_clustering.rs")).read()
    fdx = fasta(re.search(r'const FDX_FASTA: &str = "(.*?)";', rs, re.S).group(1))
    fes = fasta(open(os.path.join(REF, "test_files", "4Fe-4S-example.fasta")).read())
    sc, ai = c_oracle.parse_ncbi(ncbi_text("BLOSUM62"))
    out = {"source": "tools/gen_bucket_golden.py; sequences: bioshell-seq/tests/test_bucket_clustering.rs:8-34, "
                     "bioshell-seq/tests/test_files/4Fe-4S-example.fasta", "sets": {}}
    for name, seqs in (("FDX", fdx), ("4Fe-4S", fes)):
        n = len(seqs)
        b = [s.encode() for s in seqs]
        nid = [[0] * n for _ in range(n)]
        for i in range(n):
            for j in range(n):
                nid[i][j] = int(c_oracle.align_pair(b[i], b[j], sc, ai, -11, -1)["n_identical"])
        order = sorted(range(n), key=lambda i: -len(seqs[i]))
        cases = {}
        for level in ((0.5, 0.8, 0.9, 0.95) if name == "FDX" else (0.2, 0.25, 0.3, 0.4, 0.5)):
            # the k-mer verdict of sequence_identity (:272-292), from k-mers kept as plain substrings
            # (none of these sequences holds B or Z, the only letters that share an index)
            k = next((kk for lim, kk in ((0.95, 6), (0.90, 5), (0.85, 5), (0.80, 4), (0.75, 4), (0.70, 3), (0.60, 3), (0.50, 2))
                      if np.float32(level) >= np.float32(lim)), 1)
            ksets = [set(s[i:i + k] for i in range(len(s) - k + 1)) for s in seqs]
            verdict = [["."] * n for _ in range(n)]
            for rep in range(n):
                for c in range(n):
                    different = len(ksets[c] - ksets[rep])
                    shorter = min(len(seqs[rep]), len(seqs[c]))
                    upper = np.float32(shorter - (different // k + 1)) / np.float32(shorter)
                    lower = np.float32(shorter - (different + k - 1)) / np.float32(shorter)
                    assert shorter - (different + k - 1) >= 0      # no usize wrap on these inputs
                    lower, upper = max(lower, np.float32(0)), min(upper, np.float32(1))
                    verdict[rep][c] = "A" if lower >= np.float32(level) else ("B" if upper < np.float32(level) else "I")
            clusters = []
            for c in order:
                for cl in clusters:
                    rep = cl[0]
                    v = verdict[rep][c]
                    if v == "A" or (v == "I" and np.float32(nid[rep][c]) / np.float32(min(len(seqs[rep]), len(seqs[c]))) >= np.float32(level)):
                        cl.append(c)
                        break
                else:
                    clusters.append([c])
            cases[str(level)] = {"word_size": k, "kmer_verdict_row_rep_col_candidate": ["".join(r) for r in verdict],
                                 "clusters": clusters}
        out["sets"][name] = {"sequences": seqs, "lengths": [len(s) for s in seqs],
                             "n_identical_row_query_col_template": nid, "cases": cases}
    with open(os.path.join(ROOT, "tests", "golden", "bucket_kats.json"), "w") as f:
        json.dump(out, f, indent=1)
    for name, d in out["sets"].items():
        print(name, d["lengths"], {k: v["clusters"] for k, v in d["cases"].items()})


if __name__ == "__main__":
    main()
