#!/usr/bin/env python3
"""Regenerate bioshell_b200/data/matrices.json from the NCBI matrix files.

The seven NCBI substitution matrices (BLOSUM45/62/80, PAM30/70/120/250; public
NCBI data, https://ftp.ncbi.nih.gov/blast/matrices/) are the *input data* of the
hot path.  This script reads NCBI-format text files from a
directory (by default the read-only reference checkout, which only exists in the
build container) and stores the full 24x24 integer tables in one JSON container.
The JSON is what ships; NCBI text is re-rendered from it on demand
(`bioshell_b200.scoring.ncbi_text`) so the NCBI *parser* is still exercised.

Usage: python tools/gen_matrices.py [src_dir]
"""
import json
import os
import sys

NAMES = ["BLOSUM45", "BLOSUM62", "BLOSUM80", "PAM30", "PAM70", "PAM120", "PAM250"]


def read_ncbi(path):
    letters, rows, comments = None, [], []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith("#"):
                comments.append(line)
                continue
            if not line.strip():
                continue
            toks = line.split()
            if line.startswith(" "):
                letters = toks
                continue
            rows.append((toks[0], [int(x) for x in toks[1:]]))
    assert letters is not None and len(rows) == len(letters)
    assert [r[0] for r in rows] == letters
    return {"letters": "".join(letters), "rows": [r[1] for r in rows],
            "n_comment_lines": len(comments)}


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else \
        "/root/reference/bioshell-seq/data/substitution_matrices"
    out = {}
    for name in NAMES:
        out[name] = read_ncbi(os.path.join(src, name))
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "bioshell_b200", "data", "matrices.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
        fh.write("\n")
    print("wrote", os.path.normpath(dst), {k: v["letters"] for k, v in out.items()})


if __name__ == "__main__":
    main()
