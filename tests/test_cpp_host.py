"""The C++ host mirror (bioshell_b200/host/bioshell_seq.hpp): the reference's own aligner test
transcribed to C++ runs against the GPU through the C ABI; without a GPU the same program must
fail loudly (no CPU fallback)."""
import os
import subprocess

import pytest

from bioshell_b200 import _lib
from bioshell_b200.scoring import ncbi_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bioshell_b200", "host", "test_host_mirror")


def _run(tmp_path):
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    f = tmp_path / "BLOSUM62"
    f.write_text(ncbi_text("BLOSUM62"))
    return subprocess.run([EXE, str(f)], capture_output=True, text=True, timeout=300)


@pytest.mark.gpu
def test_cpp_host_mirror_reference_kats_on_gpu(tmp_path):
    r = _run(tmp_path)
    assert r.returncode == 0, r.stderr
    assert "host mirror ok" in r.stdout


def test_cpp_host_mirror_fails_loudly_without_gpu(tmp_path):
    if _lib.lib().bsa_device_count() > 0:
        pytest.skip("a GPU is present")
    r = _run(tmp_path)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


# ----------------------------------------------------------------------------- clustering mirror
CEXE = os.path.join(ROOT, "bioshell_b200", "host", "test_clustering_host")
LINKS = {"single": 0, "complete": 1, "average": 2}


def _clustering_case(tmp_path, n, levels, rule, seed):
    import numpy as np
    from oracle import pyhclust
    f32 = np.float32
    rng = np.random.default_rng(seed)
    d = rng.integers(0, levels, (n, n)).astype(f32)
    d = np.maximum(d, d.T)
    np.fill_diagonal(d, 0)
    root, log = pyhclust.hierarchical_clustering(n, lambda i, j: d[i, j], rule)
    cutoff, ocut = f32(levels / 2), f32(levels / 4)
    lines = ["%d %d %r %r" % (n, LINKS[rule], float(cutoff), float(ocut))]
    lines += ["%d %d %r" % (e[0], e[1], e[4]) for e in log]
    lines += [" ".join(repr(float(x)) for x in row) for row in d]
    path = tmp_path / ("case_%s_%d.txt" % (rule, n))
    path.write_text("\n".join(lines) + "\n")
    dist = lambda i, j: d[i, j]
    want = ["order " + " ".join(map(str, pyhclust.retrieve_data_id(root)))]
    cls = pyhclust.retrieve_clusters(root, cutoff)
    cls.sort(key=lambda c: c.cluster_size)
    for c in cls:
        want.append("cluster %d medoid %d : %s" % (c.cluster_size, pyhclust.medoid_by_min_max(c, dist),
                                                   " ".join(map(str, pyhclust.retrieve_data_id(c)))))
    want.append(("outliers " + " ".join(map(str, pyhclust.retrieve_outliers(n, dist, ocut)))).rstrip())
    pyhclust.balance_clustering_tree(root, dist)
    want.append("balanced " + " ".join(map(str, pyhclust.retrieve_data_id(root))))
    want.append("clustering host ok")
    return path, want


def _build_if_missing():
    if not os.path.exists(CEXE):
        import __graft_entry__ as g
        g.build()


@pytest.mark.parametrize("rule", ["single", "complete", "average"])
def test_cpp_clustering_mirror_tree_functions_match_oracle(tmp_path, rule):
    """tree rebuild, clusters, medoids, outliers and the linear-time balance of the C++ mirror against
    the literal restatement, on tie-rich matrices (no GPU: the merge log comes from the oracle)."""
    _build_if_missing()
    for n, levels, seed in ((2, 2, 1), (3, 3, 2), (25, 4, 3), (90, 6, 4), (120, 1000, 5)):
        path, want = _clustering_case(tmp_path, n, levels, rule, seed)
        r = subprocess.run([CEXE, str(path)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        got = [ln.rstrip() for ln in r.stdout.strip().split("\n")]
        assert got == [w.rstrip() for w in want]
