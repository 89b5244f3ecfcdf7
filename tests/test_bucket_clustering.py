"""bucket clustering.  CPU: the k-mer helpers of the product against the
literal restatement, and the restatement against golden clusterings of the reference's own test
sequences (tests/golden/bucket_kats.json: derived from a table of oracle identities and a table of
k-mer verdicts that can be checked by hand; tools/gen_bucket_golden.py).  GPU: the clustering with
batched aligner calls on the forward score + identity kernels -- Python driver and compiled C++ driver
(bioshell_b200/host/bioshell_bucket.hpp) -- equals the golden clusterings and the one-pair-at-a-time
restatement.  The reference's own tests only run the function (tests/test_bucket_clustering.rs:36-62)."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import bucket_clustering as bc
from bioshell_b200 import synth
from oracle import pybucket

AA = b"ARNDCQEGHILKMFPSTWYV"

FDX = ["MPKLTIVAFDGTRFDLDVDQGSTVMENAVRNSVPGIEAECGGACACATCHVYVDDEWTERVGPPEAMEEDMLDFAFDVRPTSRLSCQIRMKAALDGLTVHVPERQA",
       "MTKLTFIAHDGTQFDVDAENGSTVMENAIRNAVPGIEAECGGACACATCHVYVDEAWTAEVGEPEAMEEDMLDFAYDVQPNSRLSCQIKVRDALDGLVVRVPERQG",
       "MTKLTFIAHDGTQFDVDAENGSTVMENAIRNAVPGIEAECGGACACATCHVYVDEAWTAEVGEPEAMEEDMLDFAYDVQPNSRLSCQIKVRDALDGLVVRVPERQG",
       "MTKLTFIAHDGTHFDMDAENGSTVMENAIRNAVPGIEAECGGACACATCHVYVDEAWTAEVGEPEAMEEDMLDFAYDVQPNSRLSCQIKVRDALDGLVVRVPARQG",
       "MTKITYIAHDGSKFEVEAENGSTVMENAIRNAVPGIEAECGGACACATCHVYVDEAWSAAVGEPEAMEEDMLDFAYDVRPTSRLSCQIRVSDELDGLVVQVPERQA",
       "MPRLKFIAFDGTEFDIQADNGSTLMQNAVRNGVPGIEAECGGACACATCHVYVDEAWAEIVGPPEPMEEDMLDFAYDVRPTSRLSCQVRVREELDGLTVRIPERQG"]


def test_kmer_helpers_match_restatement():
    assert bc.standard_letter_to_index("A") == 0 and bc.standard_letter_to_index("M") == 12   # residue_types.rs:564-567
    assert bc.standard_letter_to_index("B") == 2 and bc.standard_letter_to_index("Z") == 5
    assert bc.standard_letter_to_index("a") == 26 and bc.standard_letter_to_index("-") == 30
    with pytest.raises(ValueError):
        bc.standard_letter_to_index("J")
    rng = random.Random(1)
    for _ in range(50):
        s = bytes(rng.choice(AA + b"XBZ") for _ in range(rng.randint(1, 80)))
        for k in range(0, 8):
            assert bc.generate_kmers(s, k).tolist() == pybucket.generate_kmers(s, k)
    a, b = bc.generate_kmers(FDX[0].encode(), 3), bc.generate_kmers(FDX[1].encode(), 3)
    assert bc.count_intersection_sorted(a, b) == pybucket.count_intersection_sorted(a.tolist(), b.tolist())
    for d, k, n in ((0, 3, 100), (10, 4, 100), (98, 3, 100), (98, 3, 50), (5, 6, 0), (40, 2, 41)):
        got, exp = bc.kmer_identity_bounds(d, k, n), pybucket.kmer_identity_bounds(d, k, n)
        assert (float(got[0]), float(got[1])) == (float(exp[0]), float(exp[1])), (d, k, n)
    for x in (0.99, 0.95, 0.9, 0.86, 0.8, 0.77, 0.7, 0.65, 0.5, 0.3):
        assert bc.suggest_word_length(np.float32(x)) == pybucket.suggest_word_length(x)
    with pytest.raises(AssertionError):
        bc.generate_kmers(b"AC*D", 2)         # '*' has index 33 > 31 (kmers.rs:36)


@pytest.mark.gpu
@pytest.mark.parametrize("id_level", [0.5, 0.8, 0.95])
def test_bucket_clustering_equals_restatement(ctx, oracle_matrices, id_level):
    res, off = synth.generate(150, seed=int(id_level * 100), dist=0, lo=40, hi=160, homolog_fraction=0.7)
    raw = res.tobytes()
    seqs = [raw[int(off[i]):int(off[i + 1])] for i in range(150)] + [s.encode() for s in FDX]
    S = [bs.Sequence("s%d" % i, s) for i, s in enumerate(seqs)]
    M = oracle_matrices["BLOSUM62"]
    for threads in (1, 4):
        ref = pybucket.run(seqs, id_level, M[0], M[1], threads)
        got = bc.bucket_clustering_n(S, id_level, threads, ctx)
        assert [[int(s.description()[1:]) for s in c] for c in got] == ref
    assert sum(len(c) for c in got) == len(seqs)
    if id_level == 0.8:
        assert len(ref) < len(seqs)           # the homolog families do collapse


# ----------------------------------------------------------------------------- golden clusterings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "bucket_kats.json")))
BEXE = os.path.join(ROOT, "bioshell_b200", "host", "test_bucket_host")


def _gold_cases():
    for name, d in GOLD["sets"].items():
        for level, case in d["cases"].items():
            yield name, d["sequences"], float(level), case["clusters"]


def test_restatement_reproduces_golden_clusterings(oracle_matrices):
    """The k-mer accelerated restatement against the clusterings derived from the two stored tables."""
    M = oracle_matrices["BLOSUM62"]
    for name, seqs, level, want in _gold_cases():
        b = [s.encode() for s in seqs]
        assert pybucket.run(b, level, M[0], M[1], 1) == want, (name, level)
    # the stored identity table itself: a sample of its entries against the pinned alignment oracle
    from oracle import c_oracle
    for name, d in GOLD["sets"].items():
        b = [s.encode() for s in d["sequences"]]
        for i, j in ((0, 1), (1, 0), (2, 5), (4, 3)):
            assert c_oracle.align_pair(b[i], b[j], M[0], M[1], -11, -1)["n_identical"] == \
                d["n_identical_row_query_col_template"][i][j]


@pytest.mark.gpu
def test_python_driver_reproduces_golden_clusterings(ctx):
    for name, seqs, level, want in _gold_cases():
        S = [bs.Sequence("s%d" % i, s.encode()) for i, s in enumerate(seqs)]
        got = bc.bucket_clustering(S, level, ctx)
        assert [[int(s.description()[1:]) for s in c] for c in got] == want, (name, level)


def _run_cpp(tmp_path, seqs, level, threads):
    if not os.path.exists(BEXE):
        import __graft_entry__ as g
        g.build()
    from bioshell_b200.scoring import ncbi_text
    m = tmp_path / "BLOSUM62"
    m.write_text(ncbi_text("BLOSUM62"))
    f = tmp_path / "seqs.txt"
    f.write_bytes(b"\n".join(seqs) + b"\n")
    return subprocess.run([BEXE, str(m), str(f), repr(float(level)), str(threads)], capture_output=True, text=True, timeout=600)


@pytest.mark.gpu
def test_cpp_driver_reproduces_golden_and_restatement(tmp_path, oracle_matrices):
    for name, seqs, level, want in _gold_cases():
        r = _run_cpp(tmp_path, [s.encode() for s in seqs], level, 1)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().split("\n")
        assert [[int(x) for x in l.split()] for l in lines[:-1]] == want, (name, level)
        assert lines[-1].startswith("stats ")
    # a larger synthetic set, 1 and 4 "threads", against the restatement
    res, off = synth.generate(200, seed=11, dist=0, lo=40, hi=200, homolog_fraction=0.7)
    raw = res.tobytes()
    seqs = [raw[int(off[i]):int(off[i + 1])] for i in range(200)]
    M = oracle_matrices["BLOSUM62"]
    for threads in (1, 4):
        r = _run_cpp(tmp_path, seqs, 0.8, threads)
        assert r.returncode == 0, r.stderr
        got = [[int(x) for x in l.split()] for l in r.stdout.strip().split("\n")[:-1]]
        assert got == pybucket.run(seqs, 0.8, M[0], M[1], threads)


def test_cpp_bucket_driver_fails_loudly_without_gpu(tmp_path):
    from bioshell_b200 import _lib
    if _lib.lib().bsa_device_count() > 0:
        pytest.skip("a GPU is present")
    r = _run_cpp(tmp_path, [s.encode() for s in FDX], 0.8, 1)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
