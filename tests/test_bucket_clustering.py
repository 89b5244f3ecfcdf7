"""bucket clustering.  CPU: the k-mer helpers of the product against the
literal restatement.  GPU: the clustering with batched aligner calls equals the one-pair-at-a-time
restatement (the reference's own tests only run the function, so this row is pinned by the
restatement alone)."""
import random

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
