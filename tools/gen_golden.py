#!/usr/bin/env python3
"""Generate tests/golden/oracle_vectors.json: seeded inputs + the outputs of the CPU oracle
(oracle/bioshell_oracle.c, hclust_oracle.c, local_oracle.c), committed so that (a) the oracle
cannot drift silently and (b) the GPU parity tests also compare against frozen vectors.  The
oracle itself is pinned to the reference's own KATs (tests/golden/ref_kats.json)."""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioshell_b200.scoring import ncbi_text  # noqa: E402
from oracle import c_oracle  # noqa: E402

rng = random.Random(20261017)
AA = "ARNDCQEGHILKMFPSTWYV"
DIRTY = AA + "XBZJ-_ax"
seqs = []
for i in range(36):
    n = rng.choice([1, 2, 7, 20, 33, 48, 64, 65, 90, 130])
    s = "".join(rng.choice(DIRTY if i % 5 == 0 else AA) for _ in range(n))
    if i % 3 == 1 and seqs:
        base = list(seqs[-1])
        for k in range(0, len(base), 4):
            base[k] = rng.choice(AA)
        del base[len(base) // 2: len(base) // 2 + rng.randint(0, 3)]
        s = "".join(base) or "A"
    seqs.append(s)
out = {"_comment": "frozen oracle outputs; regenerate with tools/gen_golden.py", "sequences": seqs, "global": [],
       "local": [], "hclust": []}
S = c_oracle.SeqSet([s.encode() for s in seqs])
for matrix, go, ge in (("BLOSUM62", -10, -1), ("PAM30", -11, -2), ("BLOSUM80", -4, -4)):
    sc, ai = c_oracle.parse_ncbi(ncbi_text(matrix))
    r = c_oracle.align_all_pairs(S, S, sc, ai, go, ge, True)
    paths = [c_oracle.align_pair(seqs[q].encode(), seqs[t].encode(), sc, ai, go, ge, lmax=130)["path"]
             for q, t in zip(r["q"][:120].tolist(), r["t"][:120].tolist())]
    out["global"].append({"matrix": matrix, "gap_open": go, "gap_extend": ge, "scores": r["score"].tolist(),
                          "n_identical": r["n_identical"].tolist(), "paths_first_120": paths})
    loc = [c_oracle.local_align(seqs[q].encode(), seqs[t].encode(), sc, ai, go, ge)
           for q, t in zip(r["q"][:200].tolist(), r["t"][:200].tolist())]
    out["local"].append({"matrix": matrix, "gap_open": go, "gap_extend": ge, "first_200": loc})
nrng = np.random.default_rng(7)
a = np.tril(nrng.integers(1, 30, (40, 40)).astype(np.float32) / 2, -1)
m = a + a.T
for rule in ("single", "complete", "average", "median", "centroid", "ward"):
    h = c_oracle.hclust(m, rule)
    out["hclust"].append({"rule": rule, "mat_i": h["mat_i"].tolist(), "mat_j": h["mat_j"].tolist(),
                          "dist_bits": h["dist"].view(np.uint32).tolist()})
out["hclust_matrix_lower"] = a.tolist()
with open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json"), "w") as fh:
    json.dump(out, fh, separators=(",", ":"))
print("wrote", os.path.getsize(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")), "bytes")
