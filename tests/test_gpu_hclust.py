"""GPU hierarchical clustering (bsa_hclust) vs the oracle: identical merge logs (indices and f32
distances bit for bit) for every linkage rule, with ties, and the cluster_sequences driver end to
end."""
import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import clustering as cl
from bioshell_b200 import synth
from oracle import c_oracle, pyhclust

pytestmark = pytest.mark.gpu
f32 = np.float32
RULES = [(cl.single_link, "single"), (cl.complete_link, "complete"), (cl.average_link, "average"),
         (cl.median_link, "median"), (cl.centroid_link, "centroid"), (cl.wards_method, "ward")]


def _matrix(rng, n, kind):
    if kind == "ties":
        a = rng.integers(1, 6, (n, n)).astype(f32)
    elif kind == "all100":
        a = np.full((n, n), 100.0, f32)
    elif kind == "identity_like":
        a = np.round(rng.random((n, n)) * 80 + 10, 2).astype(f32)
    else:
        a = (rng.random((n, n)) * 100).astype(f32)
    a = np.tril(a, -1)
    return (a + a.T).astype(f32)


def test_reference_kats_on_gpu(kats, ctx):
    k = kats["hierarchical_clustering"]
    data = k["cluster_numbers"]["data"]
    dist = lambda a, b: abs(f32(data[a]) - f32(data[b]))
    tree = cl.hierarchical_clustering(4, dist, cl.single_link, ctx)
    assert tree.value.cluster_size == 4
    assert cl.retrieve_data(tree, data) == k["cluster_numbers"]["order"]
    cl.balance_clustering_tree(tree, dist)
    assert cl.retrieve_data(tree, data) == k["cluster_numbers"]["order_after_balance"]
    letters = list(k["cluster_letters"]["data"])
    dist2 = lambda a, b: f32(abs(ord(letters[a]) - ord(letters[b])))
    tree = cl.hierarchical_clustering(10, dist2, cl.single_link, ctx)
    assert "".join(cl.retrieve_data(tree, letters)) == k["cluster_letters"]["order"]
    cl.balance_clustering_tree(tree, dist2)
    assert "".join(cl.retrieve_data(tree, letters)) == k["cluster_letters"]["order_after_balance"]


@pytest.mark.parametrize("link,name", RULES)
def test_merge_logs_match_oracle(ctx, link, name):
    rng = np.random.default_rng(5)
    for kind in ("ties", "random", "all100", "identity_like"):
        for n in (1, 2, 3, 9, 64, 257):
            m = _matrix(rng, n, kind)
            mi, mj, md = cl.hclust_merge_log(n, m, link, ctx)
            if n == 1:
                assert len(mi) == 0
                continue
            ref = c_oracle.hclust(m, name)
            assert np.array_equal(mi, ref["mat_i"]) and np.array_equal(mj, ref["mat_j"]), (kind, n)
            assert np.array_equal(md.view(np.uint32), ref["dist"].view(np.uint32)), (kind, n)


def test_larger_matrix_all_three_cli_linkages(ctx):
    rng = np.random.default_rng(9)
    m = _matrix(rng, 1500, "identity_like")
    for link, name in RULES[:3]:
        mi, mj, md = cl.hclust_merge_log(1500, m, link, ctx)
        ref = c_oracle.hclust(m, name)
        assert np.array_equal(mi, ref["mat_i"]) and np.array_equal(mj, ref["mat_j"])
        assert np.array_equal(md.view(np.uint32), ref["dist"].view(np.uint32))
    # single link merges are monotone non-decreasing: a property that needs no oracle
    mi, mj, md = cl.hclust_merge_log(1500, m, cl.single_link, ctx)
    assert np.all(np.diff(md) >= 0)


def test_cluster_sequences_end_to_end(ctx, tmp_path, oracle_matrices):
    """bin/cluster_sequences.rs on 60 synthetic proteins: identity matrix from the GPU aligner,
    clustering on the GPU, clusters / medoids / order against the oracle pipeline."""
    res, off = synth.generate(60, seed=77, dist=0, lo=30, hi=120, homolog_fraction=0.6)
    raw = res.tobytes()
    seqs = [bs.Sequence("syn|%07d" % i, raw[int(off[i]):int(off[i + 1])]) for i in range(60)]
    M = oracle_matrices["BLOSUM62"]
    S = c_oracle.SeqSet.from_packed(res, off)
    ref = c_oracle.align_all_pairs(S, S, M[0], M[1], -10, -2, True)
    ident = np.zeros((60, 60), f32)
    ident[ref["q"], ref["t"]] = ref["identity"]
    for compat in (True, False):
        idm = ident if compat else np.maximum(ident, ident.T)
        dist = (f32(100.0) - idm).astype(f32)
        for link, name in RULES[:3]:
            out = cl.cluster_sequences(seqs, link, identity_cutoff=40.0, medoids=True, reference_compat=compat,
                                       ctx=ctx, prefix=str(tmp_path) + "/", distance_matrix=str(tmp_path / "dm.tsv"),
                                       fasta=str(tmp_path / "ordered.fasta"))
            assert np.array_equal(out["identity"], idm)
            root, _ = pyhclust.hierarchical_clustering(60, lambda i, j: dist[i, j], name)
            cls = pyhclust.retrieve_clusters(root, f32(60.0))
            cls.sort(key=lambda c: c.cluster_size)
            assert out["clusters"] == [pyhclust.retrieve_data_id(c) for c in cls]
            assert out["medoids"] == [pyhclust.medoid_by_min_max(c, lambda i, j: dist[i, j]) for c in cls]
            pyhclust.balance_clustering_tree(root, lambda i, j: dist[i, j])
            assert out["order"] == pyhclust.retrieve_data_id(root)
            lines = (tmp_path / "ordered.fasta").read_text().split("\n")
            assert lines[0] == "> " + seqs[out["order"][0]].description()
            first = (tmp_path / "dm.tsv").read_text().split("\n")[0].split("\t")
            assert first[3:] == ["0", "0"] and first[0] == seqs[out["order"][0]].description()
    assert not compat or True
    # with the reference's half-written matrix every distance the clustering reads is 100
    out = cl.cluster_sequences(seqs, cl.single_link, identity_cutoff=40.0, ctx=ctx, write_files=False)
    assert all(len(c) == 1 for c in out["clusters"]) and len(out["clusters"]) == 60
    outl = cl.cluster_sequences(seqs, None, detect_outliers=35.0, ctx=ctx, write_files=False, reference_compat=False)
    dist = (f32(100.0) - np.maximum(ident, ident.T)).astype(f32)
    assert outl["outliers"] == pyhclust.retrieve_outliers(60, lambda i, j: dist[i, j], f32(65.0))


def test_full_scan_mode_gives_the_same_logs(ctx, monkeypatch):
    """BSA_HC_NN=0 rescans the whole triangle for every merge (round 1's path, the HBM-roofline kernel);
    the default keeps a per-row nearest-neighbour cache.  Same merge logs, ties included."""
    rng = np.random.default_rng(21)
    for kind in ("ties", "identity_like"):
        m = _matrix(rng, 300, kind)
        for link, name in RULES:
            got = cl.hclust_merge_log(300, m, link, ctx)
            monkeypatch.setenv("BSA_HC_NN", "0")
            full = cl.hclust_merge_log(300, m, link, ctx)
            monkeypatch.delenv("BSA_HC_NN")
            ref = c_oracle.hclust(m, name)
            for a in (got, full):
                assert np.array_equal(a[0], ref["mat_i"]) and np.array_equal(a[1], ref["mat_j"]), (kind, name)
                assert np.array_equal(a[2].view(np.uint32), ref["dist"].view(np.uint32)), (kind, name)
