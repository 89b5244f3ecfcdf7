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


def _product_tree_from_oracle(n, dist, rule):
    """The product's tree classes built from the oracle's merge log (no GPU involved)."""
    from bioshell_b200 import clustering as cl
    root, log = pyhclust.hierarchical_clustering(n, lambda i, j: dist[i, j], rule)
    mi = np.array([e[0] for e in log], np.uint32)
    mj = np.array([e[1] for e in log], np.uint32)
    md = np.array([e[4] for e in log], np.float32)
    return cl.tree_from_merge_log(n, mi, mj, md), root


def _shape(nd, left, right):
    """nested tuple of leaf ids / children, iterative-safe for the small trees used here"""
    if left(nd) is None and right(nd) is None:
        return nd.id
    return (_shape(left(nd), left, right), _shape(right(nd), left, right))


@pytest.mark.parametrize("rule", ["single", "complete", "average"])
def test_linear_time_balance_equals_the_literal_one(rule):
    """balance_clustering_tree with cached outermost leaves and deferred mirrors (product) against
    the literal recursive restatement of hierarchical.rs:86-100,242-287 (oracle), tie-rich input."""
    from bioshell_b200 import clustering as cl
    rng = np.random.default_rng(5)
    for n, levels in ((2, 3), (3, 2), (17, 4), (64, 6), (150, 1000)):
        d = rng.integers(0, levels, (n, n)).astype(np.float32)
        d = np.maximum(d, d.T)
        tree, oroot = _product_tree_from_oracle(n, d, rule)
        assert _shape(tree, lambda x: x._left, lambda x: x._right) == _shape(oroot, lambda x: x.left, lambda x: x.right)
        cl.balance_clustering_tree(tree, lambda i, j: d[i, j])
        pyhclust.balance_clustering_tree(oroot, lambda i, j: d[i, j])
        assert _shape(tree, lambda x: x._left, lambda x: x._right) == _shape(oroot, lambda x: x.left, lambda x: x.right)
        assert cl.retrieve_data_id(tree) == pyhclust.retrieve_data_id(oroot)


def test_tree_functions_survive_a_deep_chain():
    """Single linkage produces chain-like trees; 30,000 levels must neither recurse nor go quadratic."""
    import time
    from bioshell_b200 import clustering as cl
    n = 30000
    mi, mj = np.zeros(n - 1, np.uint32), np.ones(n - 1, np.uint32)
    md = np.arange(n - 1, dtype=np.float32)
    root = cl.tree_from_merge_log(n, mi, mj, md)
    t0 = time.perf_counter()
    cl.balance_clustering_tree(root, lambda i, j: float(abs(i - j)))
    assert time.perf_counter() - t0 < 5.0
    assert sorted(cl.retrieve_data_id(root)) == list(range(n))
    assert len(cl.retrieve_clusters(root, np.float32(n / 2))) > 1


def test_matrix_medoid_equals_the_closure_one():
    from bioshell_b200 import clustering as cl
    rng = np.random.default_rng(11)
    for n, levels in ((1, 2), (2, 2), (9, 3), (40, 4), (80, 1000)):
        d = rng.integers(0, levels, (n, n)).astype(np.float32)          # asymmetric on purpose, many ties
        if n > 5:
            d[3, 4] = np.nan
        sym = np.maximum(d, d.T) if n > 1 else d
        sym = np.nan_to_num(sym, nan=1.0)
        tree, _ = (_product_tree_from_oracle(n, sym, "average") if n > 1 else
                   (cl.tree_from_merge_log(1, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32)), None))
        for c in [tree] + cl.retrieve_clusters(tree, np.float32(levels / 2)):
            assert cl.medoid_by_min_max_matrix(c, d) == cl.medoid_by_min_max(c, lambda i, j: d[i, j])


def test_reference_balance_kats_through_the_product_tree_functions(kats):
    """tests/test_hierarchical.rs:9-41 on the product's tree code (merge log from the oracle, no GPU)."""
    from bioshell_b200 import clustering as cl
    f32 = np.float32
    k = kats["hierarchical_clustering"]
    data = k["cluster_numbers"]["data"]
    d = np.array([[abs(f32(a) - f32(b)) for b in data] for a in data], f32)
    tree, _ = _product_tree_from_oracle(len(data), d, "single")
    assert cl.retrieve_data(tree, data) == k["cluster_numbers"]["order"]
    cl.balance_clustering_tree(tree, lambda i, j: d[i, j])
    assert cl.retrieve_data(tree, data) == k["cluster_numbers"]["order_after_balance"]
    letters = list(k["cluster_letters"]["data"])
    d2 = np.array([[abs(ord(a) - ord(b)) for b in letters] for a in letters], f32)
    tree, _ = _product_tree_from_oracle(len(letters), d2, "single")
    assert "".join(cl.retrieve_data(tree, letters)) == k["cluster_letters"]["order"]
    cl.balance_clustering_tree(tree, lambda i, j: d2[i, j])
    assert "".join(cl.retrieve_data(tree, letters)) == k["cluster_letters"]["order_after_balance"]
