"""SURVEY.md 8(f) rank 2 on the CPU side: both clustering restatements (oracle/hclust_oracle.c,
oracle/pyhclust.py) reproduce the reference's own tests, agree with each other on random
matrices with ties, and the product's tree helpers (bioshell_b200/clustering.py) agree with the
oracle's on trees built from the oracle's merge logs."""
import numpy as np
import pytest

from bioshell_b200 import clustering as cl
from oracle import c_oracle, pyhclust

RULES = ["single", "complete", "average", "median", "centroid", "ward"]
f32 = np.float32


def _letters_matrix(data):
    n = len(data)
    return np.array([[abs(ord(data[a]) - ord(data[b])) for b in range(n)] for a in range(n)], f32)


def test_reference_kats_cluster_numbers_and_letters(kats):
    k = kats["hierarchical_clustering"]
    data = k["cluster_numbers"]["data"]
    dm = np.array([[abs(f32(a) - f32(b)) for b in data] for a in data], f32)
    letters = k["cluster_letters"]["data"]
    lm = _letters_matrix(letters)
    for m, items, exp, exp_bal in ((dm, data, k["cluster_numbers"]["order"], k["cluster_numbers"]["order_after_balance"]),
                                   (lm, list(letters), list(k["cluster_letters"]["order"]),
                                    list(k["cluster_letters"]["order_after_balance"]))):
        n = len(items)
        dist = lambda i, j: m[i, j]
        # python restatement
        root, log = pyhclust.hierarchical_clustering(n, dist, "single")
        assert root.cluster_size == n
        assert [items[i] for i in pyhclust.retrieve_data_id(root)] == exp
        pyhclust.balance_clustering_tree(root, dist)
        assert [items[i] for i in pyhclust.retrieve_data_id(root)] == exp_bal
        # C restatement -> product tree helpers
        r = c_oracle.hclust(m, "single")
        assert [(a, b) for a, b, *_ in log] == list(zip(r["mat_i"].tolist(), r["mat_j"].tolist()))
        tree = cl.tree_from_merge_log(n, r["mat_i"], r["mat_j"], r["dist"])
        assert tree.value.cluster_size == n
        assert cl.retrieve_data(tree, items) == exp
        cl.balance_clustering_tree(tree, dist)
        assert cl.retrieve_data(tree, items) == exp_bal


def _random_matrix(rng, n, kind):
    if kind == "ties":
        a = rng.integers(1, 6, (n, n)).astype(f32)          # many equal distances
    elif kind == "all100":
        a = np.full((n, n), 100.0, f32)                       # the reference-compat identity matrix
    else:
        a = (rng.random((n, n)) * 100).astype(f32)
    a = np.tril(a, -1)
    return (a + a.T).astype(f32)


@pytest.mark.parametrize("rule", RULES)
def test_c_vs_python_merge_logs(rule):
    rng = np.random.default_rng(7)
    for kind in ("ties", "random", "all100"):
        for n in (2, 3, 7, 24):
            m = _random_matrix(rng, n, kind)
            r = c_oracle.hclust(m, rule)
            _, log = pyhclust.hierarchical_clustering(n, lambda i, j: m[i, j], rule)
            assert list(zip(r["mat_i"].tolist(), r["mat_j"].tolist())) == [(a, b) for a, b, *_ in log], (kind, n)
            assert list(zip(r["id_i"].tolist(), r["id_j"].tolist())) == [(c, d) for _, _, c, d, _ in log]
            assert np.array_equal(r["dist"], np.array([x[4] for x in log], f32)), (kind, n, rule)


def test_only_the_lower_triangle_is_read():
    rng = np.random.default_rng(3)
    m = _random_matrix(rng, 12, "random")
    junk = m.copy()
    junk[np.triu_indices(12, 0)] = 12345.0
    a, b = c_oracle.hclust(m, "average"), c_oracle.hclust(junk, "average")
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_product_tree_helpers_match_the_oracle():
    rng = np.random.default_rng(11)
    for kind in ("ties", "random"):
        for n in (5, 17, 40):
            m = _random_matrix(rng, n, kind)
            dist = lambda i, j: m[i, j]
            for rule in ("single", "complete", "average"):
                r = c_oracle.hclust(m, rule)
                tree = cl.tree_from_merge_log(n, r["mat_i"], r["mat_j"], r["dist"])
                ref, _ = pyhclust.hierarchical_clustering(n, dist, rule)
                assert cl.retrieve_data_id(tree) == pyhclust.retrieve_data_id(ref)
                for cut in (0.0, 1.0, 2.5, 30.0, 60.0, 1000.0):
                    got = [(c.id, c.value.cluster_size) for c in cl.retrieve_clusters(tree, f32(cut))]
                    exp = [(c.id, c.cluster_size) for c in pyhclust.retrieve_clusters(ref, f32(cut))]
                    assert got == exp
                    for c, e in zip(cl.retrieve_clusters(tree, f32(cut)), pyhclust.retrieve_clusters(ref, f32(cut))):
                        assert cl.medoid_by_min_max(c, dist) == pyhclust.medoid_by_min_max(e, dist)
                cl.balance_clustering_tree(tree, dist)
                pyhclust.balance_clustering_tree(ref, dist)
                assert cl.retrieve_data_id(tree) == pyhclust.retrieve_data_id(ref)
            assert cl.retrieve_outliers(n, dist, f32(20.0)) == pyhclust.retrieve_outliers(n, dist, f32(20.0))
            assert cl.retrieve_outliers(n, dist, f32(2.0)) == pyhclust.retrieve_outliers(n, dist, f32(2.0))


def test_fasta_display_format():
    from bioshell_b200 import Sequence
    s = Sequence("2gb1", "MTYKLILNGKTLKGETTTEAVDAATAEKVFKQYANDNGVDGEWTYDDATKTFTVTE")
    # bioshell-seq/src/sequence/display_sequence.rs:22
    assert cl.format_fasta(s) == "> 2gb1\nMTYKLILNGKTLKGETTTEAVDAATAEKVFKQYANDNGVDGEWTYDDATKTFTVTE\n"
