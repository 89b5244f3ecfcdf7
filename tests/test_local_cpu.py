"""SURVEY.md 8(f) rank 4, CPU side: both restatements of LocalAlignment reproduce the reference's
own test (bioshell-seq/tests/test_aligners.rs:73-149) and agree on random pairs."""
import random

from bioshell_b200.alignment import aligned_strings
from bioshell_b200.scoring import ncbi_text
from oracle import c_oracle, pyoracle

AA = b"ARNDCQEGHILKMFPSTWYV"


def test_local_alignment_kats(kats):
    k = kats["local_alignment"]
    text = ncbi_text(k["matrix"])
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    for c in k["cases"]:
        for impl in (lambda q, t: c_oracle.local_align(q, t, sc, ai, k["gap_open"], k["gap_extend"]),
                     lambda q, t: pyoracle.local_align(q, t, psc, pai, k["gap_open"], k["gap_extend"])):
            r = impl(c["query"].encode(), c["template"].encode())
            assert r["score"] == c["score"]
            assert (r["start_q"], r["start_t"]) == (c["query_start"], c["template_start"])
            assert r["path"] == c["alignment"]
            aq, at = aligned_strings(r["path"], c["query"][r["start_q"]:], c["template"][r["start_t"]:], "-")
            assert (aq, at) == (c["aligned_query"], c["aligned_template"])
            r2 = impl(c["template"].encode(), c["query"].encode())
            assert r2["score"] == c["score"]
            aq, at = aligned_strings(r2["path"], c["template"][r2["start_q"]:], c["query"][r2["start_t"]:], "-")
            assert (aq, at) == (c["aligned_template"], c["aligned_query"])


def test_local_c_vs_python_random():
    text = ncbi_text("BLOSUM62")
    sc, ai = c_oracle.parse_ncbi(text)
    psc, pai = pyoracle.parse_ncbi(text)
    rng = random.Random(4)
    gaps = [(-10, -2), (-10, -1), (-4, -4), (-2, -1), (-3, 0), (-12, -3)]
    for k in range(300):
        n, m = rng.randint(1, 40), rng.randint(1, 40)
        q = bytes(rng.choice(AA) for _ in range(n))
        t = bytearray(rng.choice(AA) for _ in range(m))
        if k % 2:
            t = bytearray(q)
            for x in range(0, len(t), 5):
                t[x] = rng.choice(AA)
            del t[len(t) // 3: len(t) // 3 + rng.randint(0, 3)]
            t = t or bytearray(b"A")
        go, ge = gaps[k % len(gaps)]
        a = c_oracle.local_align(q, bytes(t), sc, ai, go, ge)
        b = pyoracle.local_align(q, bytes(t), psc, pai, go, ge)
        assert a == b, (q, bytes(t), go, ge)
